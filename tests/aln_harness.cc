// aln_harness.cc -- TEST INFRASTRUCTURE.  Runs the product's host-side alignment logic
// (nextdenovo_b200/csrc/nd_align.hpp, regime B / `-c`) on a CPU-only box: chains come from the oracle
// (oracle/nd_oracle.c), the DP backend is a CPU checker -- either the oracle's scalar restatement
// (ndk_extd2) or the unmodified reference function ksw_extd2_sse loaded from
// oracle/_ref/libminimap2_ref.so.  The product itself never does this: its only DP backend is the
// sm_100a kernel.
//
// usage: aln_harness <oracle|ref> <preset> <k|0> <w|0> target.2bit query.2bit out.ovl
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <dlfcn.h>
#include <string>
#include <vector>
#include "../oracle/nd_oracle.h"
#include "../nextdenovo_b200/csrc/nd_align.hpp"

extern "C" {
typedef struct { int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar, reach_end; } ndk_result;
int ndk_extd2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
              int q, int e, int q2, int e2, int w, int zdrop, int end_bonus, int flag, ndk_result *ez, uint32_t *cigar);
}

struct ref_ez { uint32_t max:31, zdropped:1; int max_q, max_t, mqe, mqe_t, mte, mte_q, score, m_cigar, n_cigar, reach_end; uint32_t *cigar; };
typedef void (*ref_extd2_t)(void *, int, const uint8_t *, int, const uint8_t *, int8_t, const int8_t *, int8_t, int8_t, int8_t, int8_t, int, int, int, int, ref_ez *);

struct CpuBackend : ndal::Backend {
    const ndal::Opt *o; int8_t mat[25]; ref_extd2_t ref = 0;
    void run(const std::vector<ndal::DpReq> &reqs, const ndal::SeqStore &T, std::vector<ndal::DpRes> &res, std::vector<uint32_t> &cigars) override
    {
        const size_t n = reqs.size();
        if (getenv("ALN_PSTAT")) {
            static const double edges[] = {1<<16, 1<<18, 1<<20, 1<<22, 1<<24, 1<<26, 1e18};
            double cnt[7] = {0}, vol[7] = {0}, cells[7] = {0};
            for (const ndal::DpReq &c : reqs) {
                int w = c.w < 0 ? std::max(c.qlen, c.tlen) : c.w; int ncol = std::min(c.qlen, c.tlen); ncol = (std::min(ncol, w + 1) + 15) / 16 + 1;
                double pb = ((double)(c.qlen + c.tlen - 1) * ncol + 1) * 16;
                int b = 0; while (pb > edges[b]) ++b;
                cnt[b] += 1; vol[b] += pb; cells[b] += pb;
            }
            fprintf(stderr, "[pstat] round n=%zu:", n);
            for (int b = 0; b < 7; ++b) fprintf(stderr, " <=%.0e: %.0f calls %.2f GB |", edges[b], cnt[b], vol[b] / 1e9);
            fprintf(stderr, "\n");
        }
        res.assign(n, ndal::DpRes());
        std::vector<std::vector<uint32_t> > cg(n);
#pragma omp parallel
        {
            std::vector<uint8_t> qb, tb; std::vector<uint32_t> cb;
#pragma omp for schedule(dynamic, 8)
            for (long long i = 0; i < (long long)n; ++i) {
                const ndal::DpReq &c = reqs[(size_t)i];
                qb.resize((size_t)c.qlen + 64); tb.resize((size_t)c.tlen + 64);
                ndal::materialize(c, T, qb.data(), tb.data());
                ndal::DpRes &z = res[(size_t)i];
                if (ref) {
                    ref_ez ez; memset(&ez, 0, sizeof ez);
                    ref(0, c.qlen, qb.data(), c.tlen, tb.data(), 5, mat, (int8_t)o->q, (int8_t)o->e, (int8_t)o->q2, (int8_t)o->e2, c.w, c.zdrop, c.end_bonus, c.flag, &ez);
                    z.max = (int32_t)ez.max; z.zdropped = ez.zdropped; z.max_q = ez.max_q; z.max_t = ez.max_t; z.mqe = ez.mqe; z.mqe_t = ez.mqe_t;
                    z.mte = ez.mte; z.mte_q = ez.mte_q; z.score = ez.score; z.n_cigar = ez.n_cigar; z.reach_end = ez.reach_end;
                    cg[(size_t)i].assign(ez.cigar, ez.cigar + ez.n_cigar);
                    free(ez.cigar);
                } else {
                    ndk_result ez; cb.resize((size_t)c.qlen + c.tlen + 8);
                    ndk_extd2(c.qlen, qb.data(), c.tlen, tb.data(), 5, mat, o->q, o->e, o->q2, o->e2, c.w, c.zdrop, c.end_bonus, c.flag, &ez, cb.data());
                    z.max = ez.max; z.zdropped = ez.zdropped; z.max_q = ez.max_q; z.max_t = ez.max_t; z.mqe = ez.mqe; z.mqe_t = ez.mqe_t;
                    z.mte = ez.mte; z.mte_q = ez.mte_q; z.score = ez.score; z.n_cigar = ez.n_cigar; z.reach_end = ez.reach_end;
                    cg[(size_t)i].assign(cb.data(), cb.data() + ez.n_cigar);
                }
            }
        }
        cigars.clear();
        for (size_t i = 0; i < n; ++i) { res[i].cig_off = (uint32_t)cigars.size(); cigars.insert(cigars.end(), cg[i].begin(), cg[i].end()); }
    }
};

struct Packed { std::vector<uint32_t> words, len, nmask; std::vector<uint64_t> word_off; };
static void pack(const ndo_reads *R, Packed &P)
{
    P.word_off.assign(1, 0);
    for (int r = 0; r < R->n; ++r) {
        const int L = R->len[r]; const size_t nw = ((size_t)(L + 15) / 16 + 1) & ~(size_t)1;
        const size_t w0 = P.words.size();
        P.words.resize(w0 + nw, 0); P.nmask.resize((w0 + nw) / 2 + 1, 0);
        for (int i = 0; i < L; ++i) {
            int c = R->seq[r][i]; c = (c == 'A' || c == 'a') ? 0 : (c == 'C' || c == 'c') ? 1 : (c == 'G' || c == 'g') ? 2 : (c == 'T' || c == 't' || c == 'U' || c == 'u') ? 3 : 4;
            if (c < 4) P.words[w0 + (i >> 4)] |= (uint32_t)c << (30 - 2 * (i & 15));
            else P.nmask[w0 / 2 + (i >> 5)] |= 1u << (i & 31);
        }
        P.len.push_back((uint32_t)L); P.word_off.push_back(P.words.size());
    }
}

// minimal FASTA reader (decimal names) into the oracle's read container
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
    if (argc != 8) { fprintf(stderr, "usage: aln_harness <oracle|ref> <preset> <k|0> <w|0> target.2bit query.2bit out.ovl\n"); return 2; }
    ndo_opt o; ndo_opt_preset(&o, NULL); ndo_opt_preset(&o, argv[2]);
    if (atoi(argv[3]) > 0) o.k = atoi(argv[3]);
    if (atoi(argv[4]) > 0) o.w = atoi(argv[4]);
    ndo_reads *T = read_any(argv[5]);
    ndo_reads *Q = strcmp(argv[5], argv[6]) ? read_any(argv[6]) : T;
    if (!T || !Q) { fprintf(stderr, "cannot read input\n"); return 1; }
    ndo_index *idx = ndo_index_build(T->n, (const char *const *)T->seq, T->len, (const char *const *)T->name, o.w, o.k, o.hpc);
    const int mid_occ = o.mid_occ > 0 ? o.mid_occ : ndo_index_mid_occ(idx, o.mid_occ_frac);

    // optional overrides (test sweeps): ALN_BW, ALN_MAXGAP, ALN_MINCNT, ALN_MINSC (chaining, as -r -g -n -m), ALN_ZDROP, ALN_ZDROP_INV, ALN_MINDP, ALN_ENDBONUS
    auto envi = [](const char *n, int d) { const char *v = getenv(n); return v ? atoi(v) : d; };
    o.bw = envi("ALN_BW", o.bw); o.max_gap = envi("ALN_MAXGAP", o.max_gap); o.min_cnt = envi("ALN_MINCNT", o.min_cnt); o.min_chain_score = envi("ALN_MINSC", o.min_chain_score);
    ndal::Opt ao; ao.k = o.k; ao.hpc = o.hpc; ao.bw = o.bw; ao.max_gap = o.max_gap; ao.min_cnt = o.min_cnt; ao.min_chain_score = o.min_chain_score;
    ao.zdrop = envi("ALN_ZDROP", ao.zdrop); ao.zdrop_inv = envi("ALN_ZDROP_INV", ao.zdrop_inv); ao.min_dp_max = envi("ALN_MINDP", ao.min_dp_max); ao.end_bonus = envi("ALN_ENDBONUS", ao.end_bonus);
    Packed PT, PQ; pack(T, PT); if (Q != T) pack(Q, PQ);
    ndal::SeqStore ST; ST.words = PT.words.data(); ST.word_off = PT.word_off.data(); ST.len = PT.len.data(); ST.nmask = PT.nmask.data();
    ndal::SeqStore SQ = ST; if (Q != T) { SQ.words = PQ.words.data(); SQ.word_off = PQ.word_off.data(); SQ.len = PQ.len.data(); SQ.nmask = PQ.nmask.data(); }

    CpuBackend be; be.o = &ao; ndal::simple_mat(be.mat, ao);
    if (!strcmp(argv[1], "ref")) {
        void *h = dlopen("oracle/_ref/libminimap2_ref.so", RTLD_NOW);
        if (!h) { fprintf(stderr, "%s\n", dlerror()); return 1; }
        be.ref = (ref_extd2_t)dlsym(h, "ksw_extd2_sse");
        if (!be.ref) { fprintf(stderr, "no ksw_extd2_sse\n"); return 1; }
    }

    std::vector<ndal::ReadJob> jobs((size_t)Q->n);
    std::vector<std::vector<nd_u128> > keep((size_t)Q->n);
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < Q->n; ++i) {
        ndal::ReadJob &J = jobs[(size_t)i];
        const int qlen = Q->len[i];
        J.qid = i; J.qlen = qlen; J.strands.resize((size_t)2 * qlen);
        ndal::encode_strands(SQ, i, qlen, J.strands.data(), J.strands.data() + qlen);
        if (qlen == 0) continue;
        ndo_u128 *mv = (ndo_u128 *)malloc(((size_t)qlen + 1) * sizeof(ndo_u128));
        size_t n_mv = ndo_sketch(Q->seq[i], qlen, o.w, o.k, 0, o.hpc, mv, 0);
        int64_t n_a; int n_u = 0; uint64_t *u = 0;
        ndo_u128 *a = ndo_collect_seeds(idx, &o, mid_occ, Q->name[i], qlen, mv, n_mv, &n_a, 1);
        a = ndo_chain_dp(o.max_gap, o.max_gap, o.bw, o.max_chain_skip, o.max_chain_iter, o.min_cnt, o.min_chain_score, n_a, a, &n_u, &u);
        free(mv);
        if (n_u == 0) { free(a); free(u); continue; }
        // mm_gen_regs (hit.c:52-88) on the oracle's chain array
        const uint32_t hash = ndo_query_hash(Q->name[i], qlen, o.seed);
        std::vector<ndo_u128> z((size_t)n_u);
        size_t k = 0;
        for (int j = 0; j < n_u; ++j) {
            uint32_t h = (uint32_t)ndo_hash64((ndo_hash64(a[k].x, ~0ULL) + ndo_hash64(a[k].y, ~0ULL)) ^ hash, ~0ULL);
            z[j].x = u[j] ^ h; z[j].y = (uint64_t)k << 32 | (uint32_t)u[j];
            k += (uint32_t)u[j];
        }
        ndo_rs_sort_128x(z.data(), z.data() + n_u);
        J.a.resize(k);
        memcpy(J.a.data(), a, k * 16);
        J.regs.resize((size_t)n_u);
        for (int j = 0; j < n_u; ++j) {
            const ndo_u128 &e = z[(size_t)(n_u - 1 - j)];
            ndal::Reg &r = J.regs[(size_t)j];
            r.id = j; r.parent = ndal::PARENT_UNSET; r.score = (int32_t)(e.x >> 32); r.hash = (uint32_t)e.x; r.cnt = (int32_t)e.y; r.as = (int32_t)(e.y >> 32);
            ndal::set_coor(r, qlen, J.a.data());
        }
        free(a); free(u);
    }
    ndal::Stats st;
    ndal::align_batch(ao, ST, jobs, be, &st);
    fprintf(stderr, "[aln_harness] tasks=%llu calls=%llu rounds=%llu cells=%.3g\n", (unsigned long long)st.n_tasks, (unsigned long long)st.n_calls,
            (unsigned long long)st.n_rounds, (double)st.cells);
    fprintf(stderr, "[aln_harness] walk %.2f s, dp %.2f s, hand-back %.2f s, finish %.2f s\n", st.t_walk, st.t_dp, st.t_hand, st.t_finish);

    FILE *fo = fopen(argv[7], "wb");
    uint32_t state[2] = {0, 0}; long n_rec = 0;
    for (int i = 0; i < Q->n; ++i)
        for (const ndal::Reg &r : jobs[(size_t)i].regs) {
            uint8_t buf[48];
            if (strcmp(Q->name[i], T->name[r.rid]) == 0) continue;
            if (r.qe - r.qs < o.minlen) continue;
            int nb = ndo_ovl_encode8(buf, state, r.rev, (uint32_t)strtoul(Q->name[i], 0, 10), r.qs, r.qe, (uint32_t)strtoul(T->name[r.rid], 0, 10), r.rs, r.re, r.mlen);
            fwrite(buf, 1, nb, fo); ++n_rec;
        }
    fclose(fo);
    fprintf(stderr, "[aln_harness] records=%ld\n", n_rec);
    return 0;
}
