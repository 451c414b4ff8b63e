// step2_harness.cc -- TEST INFRASTRUCTURE.  Runs the product's host-side `--step 2` logic
// (nextdenovo_b200/csrc/nd_step2.hpp) on a CPU-only box with the oracle (oracle/nd_oracle.c) as the
// mapping backend for the first pass and for every re-mapping.  The product's only mapping backend is the GPU.
//
// usage: step2_harness <preset> <k> <w> <minlen> <maxhan1> target.2bit query.2bit out.ovl   (writes out.ovl and out.ovl.bl)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <string>
#include "../oracle/nd_oracle.h"
#include "../nextdenovo_b200/csrc/nd_step2.hpp"

static void to_regs(const ndo_reg *r, int n, std::vector<nds2::Reg> &out)
{
    out.resize((size_t)n);
    for (int i = 0; i < n; ++i) { nds2::Reg &z = out[(size_t)i]; z.rid = r[i].rid; z.rev = r[i].rev; z.qs = r[i].qs; z.qe = r[i].qe; z.rs = r[i].rs; z.re = r[i].re; z.mlen = r[i].mlen; z.blen = r[i].blen; z.score = r[i].score; z.cnt = r[i].cnt; }
}

// minimal FASTA reader (one line per sequence or wrapped; decimal names) into the oracle's read container
static ndo_reads *read_fasta(const char *fn)
{
    FILE *f = fopen(fn, "r"); if (!f) return 0;
    ndo_reads *R = (ndo_reads *)calloc(1, sizeof(ndo_reads));
    std::vector<std::string> names, seqs; char *line = 0; size_t cap = 0; ssize_t n;
    while ((n = getline(&line, &cap, f)) > 0) {
        while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
        if (line[0] == '>') { std::string nm(line + 1); nm = nm.substr(0, nm.find_first_of(" \t")); names.push_back(nm); seqs.push_back(""); }
        else if (!seqs.empty()) seqs.back() += line;
    }
    free(line); fclose(f);
    R->n = (int)names.size(); R->seq = (char **)calloc(R->n, sizeof(char *)); R->name = (char **)calloc(R->n, sizeof(char *)); R->len = (int *)calloc(R->n, sizeof(int)); R->id = (uint32_t *)calloc(R->n, 4);
    for (int i = 0; i < R->n; ++i) { R->seq[i] = strdup(seqs[i].c_str()); R->name[i] = strdup(names[i].c_str()); R->len[i] = (int)seqs[i].size(); R->id[i] = (uint32_t)strtoul(names[i].c_str(), 0, 10); }
    return R;
}
static ndo_reads *read_any(const char *fn) { size_t L = strlen(fn); return (L > 3 && !strcmp(fn + L - 3, ".fa")) ? read_fasta(fn) : ndo_read_2bit(fn); }

int main(int argc, char **argv)
{
    if (argc != 9) { fprintf(stderr, "usage: step2_harness <preset> <k> <w> <minlen> <maxhan1> target.2bit query.2bit out.ovl\n"); return 2; }
    ndo_opt o; ndo_opt_preset(&o, NULL); ndo_opt_preset(&o, argv[1]);
    if (atoi(argv[2]) > 0) o.k = atoi(argv[2]);
    if (atoi(argv[3]) > 0) o.w = atoi(argv[3]);
    nds2::Opt so; so.minlen = atoi(argv[4]); so.maxhan1 = atoi(argv[5]);
    // optional overrides (test sweeps): S2_MAXHAN2, S2_MINMATCH, S2_CN, S2_OUTCTN, S2_MINIDE (x1000), S2_MODE
    { auto envi = [](const char *n, int d) { const char *v = getenv(n); return v ? atoi(v) : d; };
      so.maxhan2 = envi("S2_MAXHAN2", so.maxhan2); so.minmatch = envi("S2_MINMATCH", so.minmatch); so.cn = envi("S2_CN", so.cn); so.outctn = envi("S2_OUTCTN", so.outctn);
      so.minide = envi("S2_MINIDE", (int)(so.minide * 1000 + .5)) / 1000.f; so.mode = envi("S2_MODE", so.mode); }
    ndo_reads *T = read_any(argv[6]);
    ndo_reads *Q = strcmp(argv[6], argv[7]) ? read_any(argv[7]) : T;
    if (!T || !Q) { fprintf(stderr, "cannot read input\n"); return 1; }
    ndo_index *idx = ndo_index_build(T->n, (const char *const *)T->seq, T->len, (const char *const *)T->name, o.w, o.k, o.hpc);
    const int mid_occ = o.mid_occ > 0 ? o.mid_occ : ndo_index_mid_occ(idx, o.mid_occ_frac);
    std::vector<uint32_t> t_id((size_t)T->n), t_len((size_t)T->n); std::vector<uint64_t> t_key((size_t)T->n);
    for (int i = 0; i < T->n; ++i) { t_id[i] = (uint32_t)strtoul(T->name[i], 0, 10); t_key[i] = t_id[i]; t_len[i] = (uint32_t)T->len[i]; }

    std::vector<nds2::QueryJob> jobs((size_t)Q->n);
    std::vector<std::vector<nds2::Remap> > reqs((size_t)Q->n);
    long n_remap = 0, n_b2 = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+:n_remap, n_b2)
    for (int i = 0; i < Q->n; ++i) {
        nds2::QueryJob &J = jobs[(size_t)i];
        J.qlen = (uint32_t)Q->len[i]; J.qkey = strtoul(Q->name[i], 0, 10);
        int n = 0; ndo_reg *r = ndo_map_read(idx, &o, mid_occ, Q->name[i], Q->seq[i], Q->len[i], &n);
        to_regs(r, n, J.regs); free(r);
        nds2::mark_candidates(so, J, t_key.data(), t_len.data());
        nds2::plan_remaps(so, i, J, reqs[(size_t)i]);
        if (reqs[(size_t)i].empty()) continue;
        if (J.branch1) {
            const char *qs = Q->seq[i]; int ql = Q->len[i]; const char *nm = "q";
            ndo_index *mini = ndo_index_build(1, &qs, &ql, &nm, so.wn, so.kn, o.hpc);
            for (nds2::Remap &m : reqs[(size_t)i]) {
                // the reference re-maps "ACGT"[code] characters: an ambiguous base becomes '\0', which its base table reads as A (map.c:1041-1045)
                std::string ts(T->seq[m.mapped_rid], (size_t)T->len[m.mapped_rid]);
                for (char &ch : ts) if (!strchr("ACGTUacgtu", ch)) ch = 'A';
                int nn = 0; ndo_reg *x = ndo_map_read(mini, &o, mid_occ, NULL, ts.c_str(), T->len[m.mapped_rid], &nn);
                to_regs(x, nn, m.result); free(x); ++n_remap;
            }
            ndo_index_free(mini);
        } else {
            for (nds2::Remap &m : reqs[(size_t)i]) {
                std::vector<const char *> ss, nn2; std::vector<int> ll;
                for (int32_t rid : m.group_rids) { ss.push_back(T->seq[rid]); ll.push_back(T->len[rid]); nn2.push_back("t"); }
                ndo_index *mini = ndo_index_build((int)ss.size(), ss.data(), ll.data(), nn2.data(), so.wn, so.kn, o.hpc);
                int nn = 0; ndo_reg *x = ndo_map_read(mini, &o, mid_occ, NULL, Q->seq[i], Q->len[i], &nn);
                to_regs(x, nn, m.result); free(x); ++n_b2;
                ndo_index_free(mini);
            }
        }
        nds2::apply_remaps(so, J, reqs[(size_t)i], 0, reqs[(size_t)i].size(), t_len.data());
    }
    nds2::Writer W; W.begin();
    uint64_t n_rec = 0, sq = 0;
    for (int i = 0; i < Q->n; ++i)
        W.emit_query(so, jobs[(size_t)i].regs, (uint32_t)strtoul(Q->name[i], 0, 10), jobs[(size_t)i].qkey, (uint32_t)Q->len[i], t_id.data(), t_key.data(), t_len.data(), &n_rec, &sq);
    FILE *fo = fopen(argv[8], "wb"); fwrite(W.out.data(), 1, W.out.size(), fo); fclose(fo);
    std::string bl = W.bl_text();
    std::string bn = std::string(argv[8]) + ".bl";
    fo = fopen(bn.c_str(), "w"); fwrite(bl.data(), 1, bl.size(), fo); fclose(fo);
    fprintf(stderr, "[step2_harness] records=%llu remaps=%ld (query-vs-group: %ld)\n", (unsigned long long)n_rec, n_remap, n_b2);
    return 0;
}
