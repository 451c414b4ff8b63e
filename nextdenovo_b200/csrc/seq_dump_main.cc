// seq_dump -- drop-in for the reference's ingest tool (ref: util/seq_dump.c): FASTA/FASTQ[.gz] files named in a list -> `.2bit` read
// containers + `.idx` side files, the input format of `minimap2-nd --step 1/2`.  Host-only on purpose: the step
// is gzip- and file-bound, there is no arithmetic for a GPU in it.  Output is byte-identical to the reference tool (tests/test_seq_dump.py
// runs both on the same inputs), including what looks like accidents:
//   * read names are one running counter over part AND seed reads (util/seq_dump.c:88-89,117);
//   * a base outside ACGT(U) is written as A and, unless it opens a 16-base word, sets the low bit of the base before it
//     (lib/bseq.c:124 ors the table value 4 into a 2-bit field);
//   * the last character of every list line is dropped, newline or not (util/seq_dump.c:226);
//   * a FASTQ record whose quality length differs ends the reading of that file (kseq_read returns -2, util/seq_dump.c:78).
// Deliberately different: -n (>= 1) and -d are required (the reference reads uninitialised fields without them; the pipeline always
// passes all five options, nextDenovo:542-551).  The branch for seed files beyond 2^64-1 bases (util/seq_dump.c:98-116) is unreachable
// and not reproduced.
//
// usage: seq_dump -f MINLEN -s SEEDLEN -b BLOCK -n N_SEED_FILES -d OUTDIR input.fofn
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cctype>
#include <string>
#include <vector>
#include <unistd.h>
#include <libgen.h>
#include <sys/stat.h>
#include <zlib.h>

namespace {

const uint64_t LEN_LIMIT = 1000000;                        // ref: util/seq_dump.c:13

// byte stream with the end-of-file rules of the reference's reader (lib/mseq.h:56-150): "eof" is known once a refill came back short
struct Stream {
    gzFile f; std::vector<unsigned char> buf; int beg = 0, end = 0; bool eof = false;
    explicit Stream(gzFile f_) : f(f_), buf(16384) {}
    bool refill() {                                         // false: nothing more
        if (eof) return false;
        beg = 0; end = gzread(f, buf.data(), (unsigned)buf.size());
        if (end < (int)buf.size()) eof = true;
        if (end <= 0) { end = 0; return false; }
        return true;
    }
    int getc() { if (beg >= end && !refill()) return -1; return buf[beg++]; }
    // append up to the next delimiter (consumed, reported in *dret); line mode strips one trailing '\r'.  -1 only if the stream was
    // already exhausted on entry (lib/mseq.h:107)
    long until(bool line_mode, std::string &s, int *dret, bool append) {
        if (dret) *dret = 0;
        if (!append) s.clear();
        if (beg >= end && eof) return -1;
        for (;;) {
            if (beg >= end && !refill()) break;
            int i = beg;
            if (line_mode) while (i < end && buf[i] != '\n') ++i;
            else while (i < end && !isspace(buf[i])) ++i;
            s.append(reinterpret_cast<const char *>(buf.data()) + beg, (size_t)(i - beg));
            beg = i + 1;
            if (i < end) { if (dret) *dret = buf[i]; break; }
        }
        if (line_mode && s.size() > 1 && s.back() == '\r') s.pop_back();
        return (long)s.size();
    }
};

struct Record { std::string name, comment, seq, qual; int last_char = 0; };

// ref: lib/mseq.h:194-233 (kseq_read): >= 0 sequence length, -1 end of file, -2 broken quality
long next_record(Stream &ks, Record &r)
{
    int c;
    if (r.last_char == 0) {
        while ((c = ks.getc()) != -1 && c != '>' && c != '@') {}
        if (c == -1) return -1;
        r.last_char = c;
    }
    r.comment.clear(); r.seq.clear(); r.qual.clear();
    if (ks.until(false, r.name, &c, false) < 0) return -1;
    if (c != '\n') ks.until(true, r.comment, nullptr, false);
    while ((c = ks.getc()) != -1 && c != '>' && c != '+' && c != '@') {
        if (c == '\n') continue;
        r.seq.push_back((char)c);
        ks.until(true, r.seq, nullptr, true);
    }
    if (c == '>' || c == '@') r.last_char = c;
    if (c != '+') return (long)r.seq.size();
    while ((c = ks.getc()) != -1 && c != '\n') {}
    if (c == -1) return -2;
    while (ks.until(true, r.qual, nullptr, true) >= 0 && r.qual.size() < r.seq.size()) {}
    r.last_char = 0;
    if (r.seq.size() != r.qual.size()) return -2;
    return (long)r.seq.size();
}

struct Out {
    FILE *fn = nullptr, *fidx = nullptr;
    uint64_t offset = 0, length = 0;
    void open(const std::string &pre, const std::string &idx_pre, int cnt) {
        char num[16]; snprintf(num, sizeof num, "%03d", cnt);
        const std::string a = pre + num + ".2bit", b = idx_pre + num + ".idx";
        fn = fopen(a.c_str(), "w"); fidx = fopen(b.c_str(), "w");
        if (!fn || !fidx) {
            fprintf(stderr, "Failed create split file or index!\n2bit file name: %s\nidx file name: %s\n", a.c_str(), b.c_str());
            exit(1);
        }
        const unsigned char magic[2] = { 0, 254 };         // ref: lib/bseq.c:93-97
        fwrite(magic, 1, 2, fn);
        offset = 2;
    }
    void close() { if (fn) fclose(fn); if (fidx) fclose(fidx); fn = fidx = nullptr; }
    // one read: idx line, then u32 name, u32 length, 16 bases per u32 word, first base in the top bits (ref: lib/bseq.c:114-139)
    void put(const std::string &seq, uint32_t name, std::vector<uint32_t> &words) {
        static unsigned char code[256]; static bool init = false;
        if (!init) { memset(code, 4, sizeof code); code['A'] = code['a'] = 0; code['C'] = code['c'] = 1; code['G'] = code['g'] = 2; code['T'] = code['t'] = code['U'] = code['u'] = 3; init = true; }
        const uint32_t n = (uint32_t)(seq.size() < LEN_LIMIT ? seq.size() : LEN_LIMIT);
        fprintf(fidx, "%u\t%lu\t%u\n", name, (unsigned long)(offset + 8), n);
        words.clear(); words.push_back(name); words.push_back(n);
        uint64_t acc = 0; uint32_t cnt = 0;
        for (uint32_t i = 0; i < n; ++i) {
            acc = acc << 2 | code[(unsigned char)seq[i]];   // 4 spills one bit into the base before: kept, the reference's files have it
            if (++cnt == 16) { words.push_back((uint32_t)acc); acc = 0; cnt = 0; }
        }
        if (cnt) words.push_back((uint32_t)(acc << (32 - (cnt << 1))));
        fwrite(words.data(), 4, words.size(), fn);
        offset += 4 * (uint64_t)words.size();
    }
};

uint64_t parse_num(const char *s)                           // ref: util/seq_dump.c:145-154
{
    char *p; double x = strtod(s, &p);
    if (*p == 'G' || *p == 'g') x *= 1e9; else if (*p == 'M' || *p == 'm') x *= 1e6; else if (*p == 'K' || *p == 'k') x *= 1e3;
    return (uint64_t)(x + .499);
}

int usage()
{
    fprintf(stderr, "Usage: seq_dump [options] input.fofn\nDump fasta/fastq[.gz] files\n\nOptions:\n -f minimum read length\n -s minimum seed length\n"
                    " -b block size (Mb or Gb), 0=inf\n -n number of seed subfiles in total\n -d output directory\n");
    return 1;
}

}  // namespace

int main(int argc, char **argv)
{
    uint32_t min_len = 0, seed_len = 0; uint64_t block = 0; int seed_n = 0; std::string dir; bool have_b = false;
    int c;
    while ((c = getopt(argc, argv, "f:s:b:n:d:")) != -1) {
        switch (c) {
        case 'f': min_len = (uint32_t)parse_num(optarg); break;
        case 's':
            if (parse_num(optarg) <= min_len) { fprintf(stderr, "\033[1;31mError! Seed filter length should be larger than filter length!\033[0m\n"); return usage(); }
            seed_len = (uint32_t)parse_num(optarg); break;
        case 'b': block = parse_num(optarg); if (block == 0) block = UINT64_MAX; have_b = true; break;
        case 'n': seed_n = atoi(optarg); break;
        case 'd':
            if (access(optarg, 0) != 0 && mkdir(optarg, 0755) == -1) { fprintf(stderr, "Failed create output directory!\n"); return 1; }
            dir = optarg; break;
        default: return usage();
        }
    }
    if (optind + 1 > argc) return usage();
    if (seed_n < 1 || dir.empty() || !have_b) { fprintf(stderr, "seq_dump: -b, -n (>= 1) and -d are required\n"); return usage(); }
    const std::string part_pre = dir + "/input.part.", part_idx = dir + "/.input.part.", seed_pre = dir + "/input.seed.", seed_idx = dir + "/.input.seed.";
    char *fofn = strdup(argv[optind]);
    const std::string base = dirname(fofn);
    FILE *list = fopen(argv[optind], "r");
    if (!list) { fprintf(stderr, "Failed open input file list!\n"); return usage(); }

    std::vector<Out> seeds((size_t)seed_n);
    for (int i = 0; i < seed_n; ++i) seeds[(size_t)i].open(seed_pre, seed_idx, i + 1);
    Out part; int part_cnt = 1; part.open(part_pre, part_idx, part_cnt);
    uint32_t name = 0; int seed_cnt = 1;
    std::vector<uint32_t> words;

    char *line = nullptr; size_t cap = 0; ssize_t nread;
    while ((nread = getline(&line, &cap, list)) != -1) {
        if (nread == 1 || line[0] == '#') continue;
        line[strlen(line) - 1] = '\0';
        const std::string path = line[0] == '/' ? std::string(line) : base + "/" + line;
        gzFile fp = gzopen(path.c_str(), "r");
        if (!fp) { fprintf(stderr, "Error! %s does not exist!", path.c_str()); return 1; }
        gzbuffer(fp, 131072);
        Stream ks(fp); Record r;
        while (next_record(ks, r) >= 0) {
            const uint64_t L = r.seq.size();
            if (L >= min_len && L < seed_len) {                             // ref: util/seq_dump.c:79-89
                part.length += L;
                if (part.length > block) { part.close(); part.length = L; part.open(part_pre, part_idx, ++part_cnt); }
                part.put(r.seq, name++, words);
            } else if (L >= seed_len && L < LEN_LIMIT) {                    // seed files take turns (ref: :90-97)
                if (seed_cnt > seed_n) seed_cnt = 1;
                Out &o = seeds[(size_t)seed_cnt - 1];
                o.put(r.seq, name++, words); o.length += L; ++seed_cnt;
            }
        }
        gzclose(fp);
    }
    free(line);
    for (Out &o : seeds) o.close();
    part.close();
    free(fofn); fclose(list);
    return 0;
}
