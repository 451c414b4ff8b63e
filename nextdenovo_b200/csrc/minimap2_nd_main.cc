// minimap2_nd_main.cc -- `minimap2-nd`-compatible executable for the read-overlap step
// (`--step 1 ... TARGET QUERY -o OUT`), host side in C++ over the C ABI of libndgpu.so.
//
// Mirrors the parts of the reference's main() the NextDenovo pipeline drives for raw_align
// (ref: nextDenovo:445-464): two-pass option handling (minimap2/main.c:177-363), index parts
// (minimap2/index.c:284-288,353-358), query batches (minimap2/bseq.c:86-123), ordered .ovl output
// (minimap2/map.c:1245-1352).  Input: .2bit (what seq_dump writes, lib/bseq.c:93-139) or FASTA/FASTQ[.gz] with decimal names; other
// modes of the reference binary (--step 3, SAM, chain selection without MM_F_ALL_CHAINS) are rejected by name, never approximated;
// unknown options are an error as in the reference.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <chrono>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <atomic>
#include <zlib.h>
#include "ndgpu.h"
#include "nd_paf.hpp"

struct Reads {
    std::vector<uint32_t> words, len, id, hash, nmask;   // nmask: 1 bit per base at nmask[word_off/2 + (i>>5)] (FASTA/Q input only)
    std::vector<uint64_t> word_off, key;
    bool has_n = false;
    int n() const { return (int)len.size(); }
};

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int64_t parse_num(const char *s)            // ref: minimap2/main.c:102-111
{
    char *p; double x = strtod(s, &p);
    if (*p == 'G' || *p == 'g') x *= 1e9; else if (*p == 'M' || *p == 'm') x *= 1e6; else if (*p == 'K' || *p == 'k') x *= 1e3;
    return (int64_t)(x + .499);
}

static void name_fields(uint32_t id, uint64_t *key, uint32_t *hash)
{
    char buf[16]; int n = snprintf(buf, sizeof buf, "%u", id);
    uint64_t k = 0; uint32_t h = (uint32_t)buf[0];
    for (int i = 0; i < n; ++i) k |= (uint64_t)(buf[i] - '0' + 1) << (60 - 4 * i);      // strcmp-order-preserving
    for (int i = 1; i < n; ++i) h = (h << 5) - h + (uint32_t)buf[i];                     // khash.h:383-388
    *key = k; *hash = h;
}

static bool read_2bit(const char *fn, Reads &R)
{
    FILE *fp = strcmp(fn, "-") ? fopen(fn, "rb") : stdin;
    if (!fp) return false;
    uint8_t hdr[2];
    if (fread(hdr, 1, 2, fp) != 2 || !(hdr[0] == 0 || hdr[0] == 254)) {
        fprintf(stderr, "[ERROR] '%s' is not a .2bit file\n", fn);
        if (fp != stdin) fclose(fp);
        return false;
    }
    R.word_off.push_back(0);
    for (;;) {
        uint32_t h2[2];
        if (fread(h2, 4, 2, fp) != 2) break;
        size_t nw = ((size_t)h2[1] + 15) / 16, o = R.words.size();
        R.words.resize(o + nw);
        if (nw && fread(&R.words[o], 4, nw, fp) != nw) { R.words.resize(o); break; }
        uint64_t k; uint32_t h; name_fields(h2[0], &k, &h);
        R.id.push_back(h2[0]); R.len.push_back(h2[1]); R.key.push_back(k); R.hash.push_back(h);
        R.word_off.push_back(R.words.size());
    }
    if (fp != stdin) fclose(fp);
    return true;
}

// FASTA/FASTQ[.gz] (ref: lib/mseq.h kseq; names must be decimal, ref: minimap2/map.c:1300 strtoul)
static bool read_fastx(const char *fn, Reads &R)
{
    gzFile fp = strcmp(fn, "-") ? gzopen(fn, "rb") : gzdopen(0, "rb");
    if (!fp) return false;
    static unsigned char code[256]; static bool init = false;
    if (!init) { memset(code, 4, 256); code['A'] = code['a'] = 0; code['C'] = code['c'] = 1; code['G'] = code['g'] = 2; code['T'] = code['t'] = code['U'] = code['u'] = 3; init = true; }
    std::vector<char> buf(1 << 16);
    std::string line, seq, name;
    bool have = false, fq = false, ok = true;
    long qual_left = 0;
    auto flush = [&]() {
        if (!have) return;
        char *end; unsigned long v = strtoul(name.c_str(), &end, 10);
        if (name.empty() || *end) { fprintf(stderr, "[ERROR] read name '%s' is not a decimal integer; .ovl records store strtoul(name)\n", name.c_str()); ok = false; }
        size_t L = seq.size(), nw = (L + 15) / 16; if (nw & 1) ++nw;
        size_t o = R.words.size(); R.words.resize(o + nw, 0); R.nmask.resize((o + nw) / 2 + 1, 0);
        for (size_t i = 0; i < L; ++i) {
            unsigned c = code[(unsigned char)seq[i]];
            if (c < 4) R.words[o + (i >> 4)] |= c << (30 - 2 * (i & 15));
            else { R.nmask[o / 2 + (i >> 5)] |= 1u << (i & 31); R.has_n = true; }
        }
        uint64_t k; uint32_t h; name_fields((uint32_t)v, &k, &h);
        R.id.push_back((uint32_t)v); R.len.push_back((uint32_t)L); R.key.push_back(k); R.hash.push_back(h);
        R.word_off.push_back(R.words.size());
        have = false; seq.clear();
    };
    R.word_off.push_back(0);
    while (gzgets(fp, buf.data(), (int)buf.size())) {
        line.assign(buf.data());
        while (!line.empty() && line.back() != '\n' && gzgets(fp, buf.data(), (int)buf.size())) line += buf.data();
        while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
        if (qual_left > 0) { qual_left -= (long)line.size(); continue; }
        if (!line.empty() && (line[0] == '>' || (line[0] == '@' && (!have || fq)))) {
            flush();
            size_t e = 1; while (e < line.size() && !isspace((unsigned char)line[e])) ++e;
            name = line.substr(1, e - 1); have = true; fq = line[0] == '@';
        } else if (!line.empty() && line[0] == '+' && have && fq) qual_left = (long)seq.size();
        else if (have) seq += line;
    }
    flush();
    gzclose(fp);
    return ok;
}

// format sniffing as the reference (ref: lib/bseq.c:99-112): first byte 0 or 254 = .2bit
static bool read_any(const char *fn, Reads &R)
{
    if (strcmp(fn, "-") != 0) {
        FILE *fp = fopen(fn, "rb");
        if (!fp) return false;
        int c = fgetc(fp); fclose(fp);
        if (c == 0 || c == 254) return read_2bit(fn, R);
        return read_fastx(fn, R);
    }
    return read_2bit(fn, R);
}

static ndg_reads_t view(const Reads &R, int r0, int r1)
{
    ndg_reads_t v; memset(&v, 0, sizeof v); v.name_id = R.id.data() + r0;
    if (R.has_n) v.nmask = R.nmask.data();
    v.n_reads = r1 - r0; v.words = R.words.data(); v.word_off = R.word_off.data() + r0; v.len = R.len.data() + r0;
    v.name_key = R.key.data() + r0; v.name_hash = R.hash.data() + r0;
    return v;   // word_off entries are absolute into words[]; the library rebases them
}

int main(int argc, char **argv)
{
    ndg_opt_t o; memset(&o, 0, sizeof o);
    o.k = 15; o.w = 10; o.min_cnt = 3; o.min_chain_score = 40; o.bw = 500; o.max_gap = 5000; o.max_chain_skip = 25; o.max_chain_iter = 5000;
    o.seed = 11; o.mid_occ_frac = 2e-4f;
    int step = 0, minlen = 500, maxhan1 = 5000, maxhan2 = 500, dvt = 0, verbose = 3, device = 0, n_threads = 0, n_gpus = 1;
    uint64_t batch_size = 4000000000ULL, mini_batch = 50000000ULL; int64_t q_batch = 500000000;
    const char *out = nullptr; std::vector<const char *> pos;
    bool cigar = false, outraw = false;                         // outraw: text (PAF) instead of .ovl records (ref: minimap2/main.c:191,216,274, map.c:1293)
    ndg_step2_opt_t s2 = { 2000, 5000, 500, 100, 2, 20, 17, 10, 0, 0.05f };   // minimap2/main.c:194-198, options.c:50-62
    ndg_aln_opt_t ao = { 2, 4, 4, 2, 24, 1, 1, 400, 200, -1, 80, 200, 0, 1.0f };   // minimap2/options.c:37-44
    const double t0 = now();
    // Option grammar of the reference (ref: minimap2/main.c:31-100 long_options, :162 opt_str; ketopt permutes positional arguments).
    // Every option of the reference is known here with its arity; what this build does not implement is refused by name, and an
    // option the reference does not know is an error, as in its '?' branch (main.c:186-188).
    static const char *short_spec = "2aSDw:k:K:t:r:f:Vv:g:G:I:d:XT:s:x:Hcp:M:n:z:A:B:O:E:m:N:Qu:R:hF:LC:yYPo:";
    struct LongOpt { const char *name; int has_arg; const char *canon; };          // has_arg: 0 none, 1 required, 2 optional (=value only)
    static const LongOpt long_spec[] = {
        {"bucket-bits",1,0},{"mb-size",1,"-K"},{"seed",1,0},{"no-kalloc",0,0},{"print-qname",0,0},{"no-self",0,"-D"},{"print-seeds",0,0},
        {"max-chain-skip",1,0},{"min-dp-len",1,0},{"print-aln-seq",0,0},{"splice",0,0},{"cost-non-gt-ag",1,"-C"},{"no-long-join",0,0},{"sr",0,0},
        {"frag",1,0},{"secondary",1,0},{"cs",2,0},{"end-bonus",1,0},{"no-pairing",0,0},{"splice-flank",1,0},{"idx-no-seq",0,0},{"end-seed-pen",1,0},
        {"for-only",0,0},{"rev-only",0,0},{"heap-sort",1,0},{"all-chain",0,"-P"},{"dual",1,0},{"max-clip-ratio",1,0},{"min-occ-floor",1,0},{"MD",0,0},
        {"lj-min-ratio",1,0},{"score-N",1,0},{"eqx",0,0},{"paf-no-hit",0,0},{"split-prefix",1,0},{"no-end-flt",0,0},{"hard-mask-level",0,0},
        {"cap-sw-mem",1,0},{"max-qlen",1,0},{"max-chain-iter",1,0},{"junc-bed",1,0},{"junc-bonus",1,0},{"sam-hit-only",0,0},{"help",0,"-h"},
        {"max-intron-len",1,"-G"},{"version",0,"-V"},{"min-count",1,"-n"},{"min-chain-score",1,"-m"},{"mask-level",1,"-M"},{"min-dp-score",1,"-s"},{"sam",0,"-a"},
        {"step",1,0},{"minlen",1,0},{"minide",1,0},{"maxhan1",1,0},{"maxhan2",1,0},{"outraw",0,0},{"minmatch",1,0},{"mode",1,0},{"kn",1,0},{"wn",1,0},{"cn",1,0},
        {"outctn",0,0},{"df",1,0},{"dvt",0,0},
        {"device",1,0},{"gpus",1,0},                                                 // this build only: which GPU(s) to run on
        {0,0,0} };
    std::vector<std::pair<std::string, std::string>> tok;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-" || a.empty() || a[0] != '-') { pos.push_back(argv[i]); continue; }
        if (a == "--") { for (++i; i < argc; ++i) pos.push_back(argv[i]); break; }
        if (a[1] == '-') {                                      // long option
            const size_t eq = a.find('=');
            const std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            const LongOpt *lo = nullptr;
            for (const LongOpt *x = long_spec; x->name; ++x) if (name == x->name) { lo = x; break; }
            if (!lo) { fprintf(stderr, "[ERROR] unknown option in \"%s\"\n", argv[i]); return 1; }
            std::string val;
            if (eq != std::string::npos) val = a.substr(eq + 1);
            else if (lo->has_arg == 1) { if (i + 1 >= argc) { fprintf(stderr, "[ERROR] missing option argument\n"); return 1; } val = argv[++i]; }
            tok.push_back({ lo->canon ? std::string(lo->canon) : "--" + name, val });
        } else {                                                // short option(s), possibly bundled / with an attached argument
            for (size_t j = 1; j < a.size(); ++j) {
                const char *sp = strchr(short_spec, a[j]);
                if (!sp || a[j] == ':') { fprintf(stderr, "[ERROR] unknown option in \"%s\"\n", argv[i]); return 1; }
                std::string key = std::string("-") + a[j];
                if (sp[1] == ':') {
                    std::string val;
                    if (j + 1 < a.size()) val = a.substr(j + 1);
                    else { if (i + 1 >= argc) { fprintf(stderr, "[ERROR] missing option argument\n"); return 1; } val = argv[++i]; }
                    tok.push_back({ key, val });
                    break;
                }
                tok.push_back({ key, "" });
            }
        }
    }
    bool all_chains = false;
    {
        // options that change results but are not built here: refuse by name instead of mapping something else
        static const char *not_built[] = { "-a", "-S", "--cs", "--MD", "--eqx", "-Y", "-L", "-Q", "-y", "-R", "--splice", "--sr", "--frag", "--no-pairing", "--splice-flank",
            "-u", "-C", "-G", "-F", "--heap-sort", "--for-only", "--rev-only", "--end-seed-pen", "--lj-min-ratio", "--split-prefix", "--junc-bed", "--junc-bonus",
            "--sam-hit-only", "--paf-no-hit", "--secondary", "-T", "-d", "--bucket-bits", "--idx-no-seq", "--cap-sw-mem", "--max-qlen", "--print-seeds", "--print-aln-seq",
            "--print-qname", "--df", "--hard-mask-level", 0 };
        for (auto &kv : tok) for (const char **x = not_built; *x; ++x) if (kv.first == *x) {
            fprintf(stderr, "[ERROR] option '%s' is not built in this GPU build (it implements `--step 1 [-c]` and `--step 2` with .ovl output)\n", kv.first.c_str()); return 1; }
    }
    for (auto &kv : tok) {                                      // pass 1: presets (main.c:177-205)
        if (kv.first == "-x") {
            if (kv.second == "ava-ont") { all_chains = true; o.hpc = 0; o.k = 15; o.w = 5; o.flag_no_diag = o.flag_no_dual = 1; o.min_chain_score = 100; o.max_gap = 10000; o.max_chain_skip = 25; o.bw = 2000; }
            else if (kv.second == "ava-pb") { all_chains = true; o.hpc = 1; o.k = 19; o.w = 5; o.flag_no_diag = o.flag_no_dual = 1; o.min_chain_score = 100; o.max_gap = 10000; o.max_chain_skip = 25; }
            else if (kv.second == "ava-hifi") {                  // options.c:98-109: long k-mers (k = w = 51, HPC), own scoring, mid_occ_frac 1e-4
                all_chains = true; o.hpc = 1; o.k = 51; o.w = 51; o.flag_no_diag = o.flag_no_dual = 1; o.min_chain_score = 100; o.max_gap = 10000; o.max_chain_skip = 25;
                o.mid_occ_frac = 1e-4f; ao.a = 1; ao.b = 4; ao.q = 6; ao.q2 = 26; ao.e = 2; ao.e2 = 1;
            }
            else { fprintf(stderr, "[ERROR] unknown preset '%s'\n", kv.second.c_str()); return 1; }
        } else if (kv.first == "--step") {
            step = atoi(kv.second.c_str());
            outraw = true;                                      // main.c:191; `-o FILE` switches back to records (main.c:274)
            if (step == 1) minlen = 500;
            else if (step == 2) { minlen = 2000; maxhan1 = 5000; maxhan2 = 500; s2.minide = 0.05f; s2.minmatch = 100; s2.kn = 17; s2.wn = 10; s2.cn = 20; }
            else if (step == 3) minlen = 2000;
            else { fprintf(stderr, "[ERROR] unknown preset step '%s'\n", kv.second.c_str()); return 1; }
        }
    }
    for (auto &kv : tok) {                                      // pass 2 (main.c:209-363)
        const std::string &k = kv.first; const char *v = kv.second.c_str();
        if (k == "-w") o.w = atoi(v); else if (k == "-k") o.k = atoi(v); else if (k == "-H") o.hpc = 1;
        else if (k == "-r") o.bw = (int)parse_num(v); else if (k == "-g") o.max_gap = (int)parse_num(v);
        else if (k == "-n") o.min_cnt = atoi(v); else if (k == "-m") o.min_chain_score = atoi(v);
        else if (k == "-I") batch_size = (uint64_t)parse_num(v); else if (k == "-K") q_batch = parse_num(v);
        else if (k == "-D") o.flag_no_diag = 1; else if (k == "-X") { o.flag_no_diag = o.flag_no_dual = 1; all_chains = true; }
        else if (k == "-P") all_chains = true;
        else if (k == "-c") cigar = true;
        else if (k == "--min-dp-len") ao.min_ksw_len = atoi(v); else if (k == "--score-N") ao.sc_ambi = atoi(v);
        else if (k == "-A") ao.a = atoi(v); else if (k == "-B") ao.b = atoi(v);
        else if (k == "-s" || k == "--min-dp-score") ao.min_dp_max = atoi(v);
        else if (k == "--end-bonus") ao.end_bonus = atoi(v); else if (k == "--max-clip-ratio") ao.max_clip_ratio = (float)atof(v);
        else if (k == "--no-end-flt") ao.no_end_flt = 1;
        else if (k == "--minide") s2.minide = (float)atof(v); else if (k == "--minmatch") s2.minmatch = (int)parse_num(v);
        else if (k == "--mode") s2.mode = atoi(v); else if (k == "--kn") s2.kn = atoi(v); else if (k == "--wn") s2.wn = atoi(v);
        else if (k == "--cn") s2.cn = atoi(v); else if (k == "--outctn") s2.outctn = 1;
        else if (k == "-z") { char *e; ao.zdrop = ao.zdrop_inv = (int)strtol(v, &e, 10); if (*e == ',') ao.zdrop_inv = (int)strtol(e + 1, &e, 10); }
        else if (k == "-O") { char *e; ao.q = ao.q2 = (int)strtol(v, &e, 10); if (*e == ',') ao.q2 = (int)strtol(e + 1, &e, 10); }
        else if (k == "-E") { char *e; ao.e = ao.e2 = (int)strtol(v, &e, 10); if (*e == ',') ao.e2 = (int)strtol(e + 1, &e, 10); }
        else if (k == "--minlen") minlen = (int)parse_num(v); else if (k == "--maxhan1") maxhan1 = (int)parse_num(v);
        else if (k == "--maxhan2") maxhan2 = (int)parse_num(v); else if (k == "--dvt") dvt = 1;
        else if (k == "--seed") o.seed = atoi(v); else if (k == "--max-chain-skip") o.max_chain_skip = atoi(v);
        else if (k == "--max-chain-iter") o.max_chain_iter = atoi(v); else if (k == "--min-occ-floor") o.min_mid_occ = atoi(v);
        else if (k == "-v") verbose = atoi(v); else if (k == "--device") device = atoi(v); else if (k == "-t") n_threads = atoi(v);
        else if (k == "--gpus") n_gpus = atoi(v);
        else if (k == "--dual") { if (kv.second == "yes" || kv.second == "y") o.flag_no_dual = 0; else if (kv.second == "no" || kv.second == "n") o.flag_no_dual = 1; }
        else if (k == "-f") { double x = strtod(v, nullptr); if (x < 1.0) { o.mid_occ_frac = (float)x; o.mid_occ = 0; } else o.mid_occ = (int)(x + .499); }
        else if (k == "-o") { if (kv.second != "-") { out = v; if (step) outraw = false; } }
        else if (k == "--outraw") outraw = true;
        else if (k == "-V" || k == "--version") { puts("2.17-r941-ndgpu"); return 0; }
    }
    if (pos.size() < 2) {
        fprintf(stderr, "Usage: minimap2-nd --step 1 [options] <target.2bit> <query.2bit> -o out.ovl\n");
        return 1;
    }
    if (!all_chains) {          // the library keeps every chain (no mm_set_parent / mm_select_sub / best_n / pri_ratio, ref: minimap2/map.c:250-259)
        fprintf(stderr, "[ERROR] this GPU build maps in all-chain mode only (MM_F_ALL_CHAINS): use -x ava-ont|ava-pb, -X or -P\n");
        return 1;
    }
    const bool paf = outraw || !step;                           // map.c:1293: `outraw || !step` prints PAF lines
    if (paf && step != 0 && step != 1) { fprintf(stderr, "[ERROR] PAF output is built for plain mapping and `--step 1 [-c]` only\n"); return 1; }
    if (!paf && ((step != 1 && step != 2) || !out || (step == 2 && cigar))) {
        fprintf(stderr, "[ERROR] this GPU build implements `--step 1 [-c] ... -o FILE` and `--step 2 ... -o FILE` (.ovl output), and PAF text without `-o`/with --outraw\n");
        return 1;
    }
    if (step == 2 && s2.mode != 0 && s2.mode != 2) { fprintf(stderr, "[ERROR] --mode %d is not built (0 and 2 are)\n", s2.mode); return 1; }
    if (cigar) {                                                // mm_check_opt, minimap2/options.c:190-213
        if (ao.q <= 0 || ao.q2 <= 0 || ao.e <= 0 || ao.e2 <= 0) { fprintf(stderr, "[ERROR]\033[1;31m -O and -E must be positive\033[0m\n"); return 1; }
        if ((ao.q != ao.q2 || ao.e != ao.e2) && !(ao.e > ao.e2 && ao.q + ao.e < ao.q2 + ao.e2)) { fprintf(stderr, "[ERROR]\033[1;31m dual gap penalties violating E1>E2 and O1+E1<O2+E2\033[0m\n"); return 1; }
        if ((ao.q + ao.e) + (ao.q2 + ao.e2) > 127) { fprintf(stderr, "[ERROR]\033[1;31m scoring system violating ({-O}+{-E})+({-O2}+{-E2}) <= 127\033[0m\n"); return 1; }
        if (ao.zdrop < ao.zdrop_inv) { fprintf(stderr, "[ERROR]\033[1;31m Z-drop should not be less than inversion-Z-drop\033[0m\n"); return 1; }
    }
    FILE *fo = out ? fopen(out, "wb") : stdout;                 // main.c:259-262 (freopen)
    if (!fo) { fprintf(stderr, "[ERROR]\033[1;31m failed to write the output to file '%s'\033[0m\n", out); return 1; }
    FILE *fbl = nullptr;
    if (step == 2) {                                            // main.c:263-272: FILE.bl next to the records
        std::string bn = std::string(out) + ".bl";
        if (!(fbl = fopen(bn.c_str(), "w"))) { fprintf(stderr, "[ERROR]\033[1;31m failed to write the output to file '%s'\033[0m\n", bn.c_str()); return 1; }
    }

    Reads T;
    if (!read_any(pos[0], T)) { fprintf(stderr, "[ERROR] failed to open file '%s'\n", pos[0]); return 1; }
    // ---- contexts: G GPUs (--gpus, devices device..device+G-1) x C contexts per GPU.  The contexts of a GPU share one copy of the
    // index (ndg_index_share) and map different query batches concurrently, each on its own stream: upload, kernels and download of
    // neighbouring batches overlap (what the reference's 3-stage kt_pipeline does with threads, ref: minimap2/map.c:1389-1390).
    // Step 2 keeps one context: its output stage is order-dependent (filter_ovl, ref: lib/ovl.c:449-563).
    if (n_gpus < 1) n_gpus = 1;
    if (step == 2 || paf) n_gpus = 1;
    const int G = n_gpus, CPG = (step == 2 || cigar || paf) ? 1 : 2;
    std::vector<ndg_ctx_t *> ctx((size_t)G * CPG, nullptr);
    for (int g = 0; g < G; ++g)
        for (int x = 0; x < CPG; ++x) {
            ndg_ctx_t *&h = ctx[(size_t)g * CPG + x];
            if (ndg_init(device + g, &h) != NDG_OK) { fprintf(stderr, "[ERROR] no usable B200 (sm_100) device %d; there is no CPU fallback\n", device + g); return 1; }
            ndg_set_opt(h, &o);
            if (cigar && ndg_set_aln_opt(h, &ao)) { fprintf(stderr, "[ERROR] %s\n", ndg_last_error(h)); return 1; }
        }
    ndg_ctx_t *g = ctx[0];
    if (paf && ndg_set_paf(g, 1)) { fprintf(stderr, "[ERROR] %s\n", ndg_last_error(g)); return 1; }
    if (step == 2) { s2.minlen = minlen; s2.maxhan1 = maxhan1; s2.maxhan2 = maxhan2; if (ndg_step2_begin(g, &s2)) { fprintf(stderr, "[ERROR] %s\n", ndg_last_error(g)); return 1; } }
    n_threads = ndg_set_threads(g, n_threads);                  // -t: host threads (only the -c stage has host-side work worth threading)
    if (cigar && G > 1) for (int x = 0; x < G; ++x) ndg_set_threads(ctx[(size_t)x * CPG], n_threads / G > 0 ? n_threads / G : 1);
#define DIE(h) do { fprintf(stderr, "[ERROR] %s\n", ndg_last_error(h)); exit(1); } while (0)

    uint32_t state[2] = { 0, 0 };
    uint64_t n_rec = 0, sum_qspan = 0;
    const uint64_t mini = mini_batch < batch_size ? mini_batch : batch_size;
    int t_at = 0, n_parts = 0;
    while (t_at < T.n()) {                                     // one index part per iteration (main.c:477)
        int t0i = t_at; uint64_t total = 0;
        while (t_at < T.n() && !(total > batch_size)) {         // index.c:286: stop once sum_len > batch_size
            uint64_t size = 0;
            while (t_at < T.n()) { size += T.len[t_at++]; if (size >= mini) break; }
            total += size;
        }
        ndg_index_stat_t st;
        if (G == 1) {
            ndg_reads_t tv = view(T, t0i, t_at);
            if (ndg_index_begin(g) || ndg_index_add(g, &tv) || ndg_index_finalize(g, &st)) DIE(g);
        } else {
            // N-GPU build of the part (ndg_index_shard_*): GPU r sketches the r-th share of the targets and builds the r-th key range of
            // the table; the collectives are peer copies between the primaries' buffers
            std::vector<ndg_ctx_t *> P((size_t)G);
            std::vector<int> cut((size_t)G + 1);
            for (int r = 0; r <= G; ++r) cut[(size_t)r] = t0i + (int)((int64_t)(t_at - t0i) * r / G);
            for (int r = 0; r < G; ++r) P[(size_t)r] = ctx[(size_t)r * CPG];
            std::vector<std::vector<uint64_t>> cm((size_t)G, std::vector<uint64_t>((size_t)G));
            for (int r = 0; r < G; ++r) {
                ndg_ctx_t *h = P[(size_t)r];
                if (ndg_index_begin(h)) DIE(h);
                for (int j = 0; j < G; ++j) {
                    if (cut[(size_t)j] == cut[(size_t)j + 1]) continue;
                    ndg_reads_t tv = view(T, cut[(size_t)j], cut[(size_t)j + 1]);
                    if (j == r ? ndg_index_add(h, &tv) : ndg_index_add_meta(h, &tv)) DIE(h);
                }
                if (ndg_index_shard_split(h, G, cm[(size_t)r].data())) DIE(h);
            }
            std::vector<uint64_t> npos((size_t)G, 0), nd((size_t)G, 0);
            for (int j = 0; j < G; ++j) {                       // all-to-all: GPU j receives bucket j of every GPU, in GPU (= target) order
                uint64_t tot = 0; for (int r = 0; r < G; ++r) tot += cm[(size_t)r][(size_t)j];
                npos[(size_t)j] = tot;
                void *rk, *rv;
                if (ndg_index_shard_recv_reserve(P[(size_t)j], tot, &rk, &rv)) DIE(P[(size_t)j]);
                uint64_t at = 0;
                for (int r = 0; r < G; ++r) {
                    const void *kp, *vp; uint64_t nn;
                    if (ndg_index_shard_bucket(P[(size_t)r], j, &kp, &vp, &nn)) DIE(P[(size_t)r]);
                    if (ndg_copy_between(P[(size_t)j], (uint64_t *)rk + at, P[(size_t)r], kp, nn * 8) || ndg_copy_between(P[(size_t)j], (uint64_t *)rv + at, P[(size_t)r], vp, nn * 8)) DIE(P[(size_t)j]);
                    at += nn;
                }
            }
            uint64_t nd_max = 0, nd_sum = 0;
            for (int r = 0; r < G; ++r) { if (ndg_index_shard_sort(P[(size_t)r], &nd[(size_t)r])) DIE(P[(size_t)r]); nd_max = nd_max > nd[(size_t)r] ? nd_max : nd[(size_t)r]; nd_sum += nd[(size_t)r]; }
            std::vector<uint32_t> hist(65536, 0), hr(65536);
            for (int r = 0; r < G; ++r) { if (ndg_index_shard_place(P[(size_t)r], r, nd_max, hr.data())) DIE(P[(size_t)r]); for (int b = 0; b < 65536; ++b) hist[(size_t)b] += hr[(size_t)b]; }
            std::vector<std::vector<void *>> seg((size_t)G, std::vector<void *>((size_t)G));
            for (int r = 0; r < G; ++r) if (ndg_index_shard_pos_reserve(P[(size_t)r], r, npos.data(), seg[(size_t)r].data())) DIE(P[(size_t)r]);
            for (int r = 0; r < G; ++r)                         // all-gather of the finished table slices and position segments
                for (int j = 0; j < G; ++j) {
                    if (j == r) continue;
                    void *dp, *sp; uint64_t nb, nb2;
                    if (ndg_index_shard_table_slice(P[(size_t)r], j, &dp, &nb) || ndg_index_shard_table_slice(P[(size_t)j], j, &sp, &nb2)) DIE(P[(size_t)r]);
                    if (ndg_copy_between(P[(size_t)r], dp, P[(size_t)j], sp, nb) || ndg_copy_between(P[(size_t)r], seg[(size_t)r][(size_t)j], P[(size_t)j], seg[(size_t)j][(size_t)j], npos[(size_t)j] * 8)) DIE(P[(size_t)r]);
                }
            for (int r = 0; r < G; ++r) if (ndg_index_shard_finish(P[(size_t)r], hist.data(), nd_sum, &st)) DIE(P[(size_t)r]);
        }
        for (int x = 0; x < G; ++x) for (int y = 1; y < CPG; ++y) if (ndg_index_share(ctx[(size_t)x * CPG + y], ctx[(size_t)x * CPG])) DIE(ctx[(size_t)x * CPG + y]);
        ++n_parts;
        if (verbose >= 3) {
            fprintf(stderr, "[M::main::%.3f] loaded/built the index for %llu target sequence(s)\n", now() - t0, (unsigned long long)st.n_seq);
            fprintf(stderr, "[M::mm_mapopt_update::%.3f] mid_occ = %d\n", now() - t0, st.mid_occ);
            fprintf(stderr, "[M::mm_idx_stat::%.3f] distinct minimizers: %llu (%.2f%% are singletons); average occurrences: %.3lf; average spacing: %.3lf\n", now() - t0,
                    (unsigned long long)st.n_distinct, 100.0 * st.n_singletons / st.n_distinct, (double)st.n_minimizers / st.n_distinct, (double)st.sum_len / st.n_minimizers);
        }
        for (size_t qf = 1; qf < pos.size(); ++qf) {
            Reads Qown; const Reads *Q = &T;
            if (strcmp(pos[qf], pos[0]) != 0) { if (!read_any(pos[qf], Qown)) { fprintf(stderr, "ERROR: failed to open file '%s'\n", pos[qf]); return 1; } Q = &Qown; }
            std::vector<std::pair<int, int>> qb;               // bseq.c:86-123: a batch ends once >= chunk_size bases were read
            for (int q_at = 0; q_at < Q->n();) { int q0 = q_at; int64_t size = 0; while (q_at < Q->n()) { size += Q->len[q_at++]; if (size >= q_batch) break; } qb.push_back({ q0, q_at }); }
            const ndg_emit_opt_t eo = { minlen, dvt, maxhan1, maxhan2 };
            if (paf) {                                          // text output: one line per hit, in query order (map.c:1289-1296)
                const ndpaf::Par pp = { o.min_chain_score, ao.a, 0, cigar ? 1 : 0 };
                std::string text; std::vector<int> mapq;
                for (auto &b : qb) {
                    ndg_reads_t qv = view(*Q, b.first, b.second);
                    ndg_map_result_t R;
                    const ndg_hit_extra_t *he = nullptr; const ndg_query_extra_t *qe = nullptr; const uint32_t *cg = nullptr;
                    if (ndg_map_batch(g, &qv, &R) || ndg_map_extras(g, &he, &qe, &cg)) DIE(g);
                    text.clear();
                    for (int64_t i = 0; i < R.n_hits;) {
                        int64_t j = i; while (j < R.n_hits && R.hits[j].qidx == R.hits[i].qidx) ++j;
                        const int qi = R.hits[i].qidx, qa = b.first + qi;
                        ndpaf::set_mapq(pp, R.hits + i, he + i, (int)(j - i), qe[qi].rep_len, mapq);
                        for (int64_t k = i; k < j; ++k) {
                            const ndg_hit_t &h = R.hits[k];
                            ndpaf::write_hit(text, pp, Q->id[(size_t)qa], (int)Q->len[(size_t)qa], T.id[(size_t)(t0i + h.rid)], (int)T.len[(size_t)(t0i + h.rid)], h, he[k], qe[qi], mapq[(size_t)(k - i)], cg);
                            text.push_back('\n');
                            ++n_rec; sum_qspan += (uint64_t)(h.qe - h.qs);
                        }
                        i = j;
                    }
                    if (!text.empty() && fwrite(text.data(), 1, text.size(), fo) != text.size()) { fprintf(stderr, "[ERROR] failed to write the results\n"); exit(EXIT_FAILURE); }
                    if (verbose >= 3) fprintf(stderr, "[M::worker_pipeline::%.3f] mapped %d sequences\n", now() - t0, b.second - b.first);
                }
                continue;
            }
            if (ctx.size() == 1) {                              // one context: batches in order, the encoder state carried along
                for (auto &b : qb) {
                    ndg_reads_t qv = view(*Q, b.first, b.second);
                    ndg_map_result_t R;
                    const uint8_t *bytes = nullptr; uint64_t nb = 0;
                    if (step == 2 ? ndg_step2_map_batch(g, &qv, &bytes, &nb, &n_rec, &sum_qspan, &R)
                                  : ndg_map_batch_ovl(g, &qv, &eo, state, &bytes, &nb, &n_rec, &sum_qspan, &R)) DIE(g);
                    if (nb && fwrite(bytes, 1, (size_t)nb, fo) != (size_t)nb) { fprintf(stderr, "[ERROR] failed to write the results\n"); exit(EXIT_FAILURE); }
                    if (verbose >= 3) fprintf(stderr, "[M::worker_pipeline::%.3f] mapped %d sequences\n", now() - t0, b.second - b.first);
                }
                continue;
            }
            // several contexts: workers pull batches; every batch is encoded from state (0,0); the writer puts them out in input order and
            // re-bases the first record of each batch onto the names the stream really ended with (ndg_ovl_rebase_first)
            struct Done { std::vector<uint8_t> bytes; uint32_t st[2]; uint64_t rec = 0, sq = 0; bool ready = false; };
            std::vector<Done> done(qb.size());
            std::atomic<size_t> next(0); std::mutex mu; std::condition_variable cv; std::atomic<int> failed(0);
            auto work = [&](ndg_ctx_t *h) {
                for (;;) {
                    const size_t i = next.fetch_add(1);
                    if (i >= qb.size() || failed.load()) break;
                    ndg_reads_t qv = view(*Q, qb[i].first, qb[i].second);
                    ndg_map_result_t R; const uint8_t *bytes = nullptr; uint64_t nb = 0;
                    Done &d = done[i]; d.st[0] = d.st[1] = 0;
                    if (ndg_map_batch_ovl(h, &qv, &eo, d.st, &bytes, &nb, &d.rec, &d.sq, &R)) { fprintf(stderr, "[ERROR] %s\n", ndg_last_error(h)); failed.store(1); }
                    else d.bytes.assign(bytes, bytes + nb);
                    { std::lock_guard<std::mutex> lk(mu); d.ready = true; }
                    cv.notify_all();
                }
            };
            std::vector<std::thread> th;
            for (auto h : ctx) th.emplace_back(work, h);
            for (size_t i = 0; i < qb.size(); ++i) {
                { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return done[i].ready; }); }
                if (failed.load()) break;
                Done &d = done[i];
                if (!d.bytes.empty()) {
                    uint8_t first[48]; uint32_t skip = 0;
                    const int fl = ndg_ovl_rebase_first(d.bytes.data(), d.bytes.size(), state, first, &skip);
                    if (fl < 0) { fprintf(stderr, "[ERROR] malformed batch stream\n"); failed.store(1); break; }
                    if (fwrite(first, 1, (size_t)fl, fo) != (size_t)fl || fwrite(d.bytes.data() + skip, 1, d.bytes.size() - skip, fo) != d.bytes.size() - skip) { fprintf(stderr, "[ERROR] failed to write the results\n"); exit(EXIT_FAILURE); }
                    state[0] = d.st[0]; state[1] = d.st[1];
                }
                n_rec += d.rec; sum_qspan += d.sq;
                std::vector<uint8_t>().swap(d.bytes);
                if (verbose >= 3) fprintf(stderr, "[M::worker_pipeline::%.3f] mapped %d sequences\n", now() - t0, qb[i].second - qb[i].first);
            }
            for (auto &t : th) t.join();
            if (failed.load()) return 1;
        }
    }
    if (fbl) {                                                  // main.c:522-526 (out_bl)
        const char *bl = nullptr; uint64_t nbl = 0;
        if (ndg_step2_finish(g, &bl, &nbl)) { fprintf(stderr, "[ERROR] %s\n", ndg_last_error(g)); return 1; }
        if ((nbl && fwrite(bl, 1, (size_t)nbl, fbl) != (size_t)nbl) || fclose(fbl) != 0) { fprintf(stderr, "[ERROR] failed to write the results\n"); exit(EXIT_FAILURE); }
    }
    if ((fo == stdout ? fflush(fo) : fclose(fo)) != 0) { fprintf(stderr, "[ERROR] failed to write the results\n"); exit(EXIT_FAILURE); }
    if (verbose >= 3 && cigar) {
        ndg_aln_stat_t as; ndg_aln_stats(g, &as);
        fprintf(stderr, "[M::align] chains %llu; DP calls %llu in %llu rounds / %llu launches (%llu retried); %.3g cells; DP kernels %.3f s (backend wall %.3f s); alignment stage %.3f s on %d host threads\n", (unsigned long long)as.n_chains,
                (unsigned long long)as.n_calls, (unsigned long long)as.n_rounds, (unsigned long long)as.n_launches, (unsigned long long)as.n_retries, (double)as.p_bytes, as.dp_ms / 1e3, as.dp_wall_ms / 1e3, as.host_ms / 1e3, n_threads);
    }
    if (verbose >= 3) {
        fprintf(stderr, "[M::main] Version: 2.17-r941-ndgpu (%s)\n[M::main] CMD:", ndg_version());
        for (int i = 0; i < argc; ++i) fprintf(stderr, " %s", argv[i]);
        fprintf(stderr, "\n[M::main] Real time: %.3f sec; records: %llu; index parts: %d\n", now() - t0, (unsigned long long)n_rec, n_parts);
    }
    for (auto h : ctx) ndg_destroy(h);
    return 0;
}
