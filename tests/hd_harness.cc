// hd_harness.cc -- TEST INFRASTRUCTURE.  Compiles the __host__ __device__ building blocks of the
// CUDA path (nextdenovo_b200/csrc/nd_core.cuh) for the host and checks them, stage by stage,
// against the oracle restatement on a .2bit file.  Lets the exact-semantics logic be validated on a
// box without a GPU; the product never runs this code on the CPU.
//
// usage: hd_harness reads.2bit preset(k,w,hpc as "ava-ont"|"ava-pb"|"k17w10") [max_reads] [chunk]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <string>
#include "../nextdenovo_b200/csrc/nd_core.cuh"
#include "../oracle/nd_oracle.h"

static std::vector<uint32_t> pack2bit(const char *s, int len)
{
    std::vector<uint32_t> w((len + 15) / 16 + 1, 0);
    for (int i = 0; i < len; ++i) {
        int c = s[i] == 'A' ? 0 : s[i] == 'C' ? 1 : s[i] == 'G' ? 2 : 3;
        w[i >> 4] |= (uint32_t)c << (30 - 2 * (i & 15));
    }
    return w;
}

// ambiguous bases: chunked state machine with the 1-bit mask vs the oracle on ASCII with 'N'
static int selftest_n()
{
    unsigned long long st = 88172645463325252ULL;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    long bad = 0, tot = 0;
    const int cfg[][3] = { {5, 15, 0}, {10, 17, 0}, {7, 16, 0}, {5, 19, 1}, {3, 11, 1}, {1, 9, 0},
                           {51, 51, 1}, {51, 51, 0}, {7, 33, 1}, {9, 32, 1}, {20, 64, 1}, {5, 40, 0}, {13, 29, 1}, {6, 100, 1} };   // long k-mers (k > 28)
    for (int it = 0; it < 1500; ++it) {
        int len = 1 + (int)(rnd() % 1200);
        std::string s(len, 'A');
        for (int i = 0; i < len; ++i) s[i] = "ACGT"[rnd() & 3];
        int nN = (int)(rnd() % 6);
        for (int j = 0; j < nN; ++j) { int p = (int)(rnd() % len), run = 1 + (int)(rnd() % (rnd() % 4 == 0 ? 40 : 3)); for (int q = p; q < len && q < p + run; ++q) s[q] = 'N'; }
        if (it % 5 == 0) { int p = (int)(rnd() % len), run = (int)(rnd() % 60); for (int q = p; q < len && q < p + run; ++q) s[q] = 'A'; }
        std::vector<uint32_t> words((len + 15) / 16 + 2, 0), mask((len + 31) / 32 + 2, 0);
        for (int i = 0; i < len; ++i) {
            int c = s[i] == 'A' ? 0 : s[i] == 'C' ? 1 : s[i] == 'G' ? 2 : s[i] == 'T' ? 3 : 4;
            if (c == 4) mask[i >> 5] |= 1u << (i & 31); else words[i >> 4] |= (uint32_t)c << (30 - 2 * (i & 15));
        }
        for (auto &c : cfg) {
            std::vector<ndo_u128> ref(len + 1);
            size_t n_ref = ndo_sketch(s.data(), len, c[0], c[1], 9, c[2], ref.data(), 0);
            nd_seqview sv = { words.data(), mask.data(), len };
            std::vector<nd_u128> got;
            int cl = 16 + (int)(rnd() % 300);
            for (int cs = 0; cs < len; cs += cl) {
                int ce = cs + cl < len ? cs + cl : len;
                std::vector<nd_u128> tmp(cl + c[0] + 8);
                int n = (c[1] > 28 ? nd_sketch_chunk<64, true>(sv, c[0], c[1], c[2], 9, cs, ce, tmp.data(), (int)tmp.size()) : nd_sketch_chunk<16, false>(sv, c[0], c[1], c[2], 9, cs, ce, tmp.data(), (int)tmp.size()));
                got.insert(got.end(), tmp.begin(), tmp.begin() + std::min<int>(n, (int)tmp.size()));
            }
            bool ok = got.size() == n_ref;
            for (size_t i = 0; ok && i < n_ref; ++i) ok = got[i].x == ref[i].x && got[i].y == ref[i].y;
            if (!ok) { if (bad < 5) fprintf(stderr, "N selftest mismatch it=%d len=%d w=%d k=%d hpc=%d: %zu vs %zu\n", it, len, c[0], c[1], c[2], got.size(), n_ref); ++bad; }
            ++tot;
        }
    }
    printf("selftest_n cases=%ld bad=%ld\n", tot, bad);
    return bad != 0;
}

// ordered index table: the placement rule of k_index_place (slot_i = i + running max of (ideal_i - i)) and the forward scan of
// nd_idx_lookup, replayed on the host with the same nd_idx_slice / nd_idx_ideal: every key is found, absent keys are not, slices
// partition the key range monotonically, nothing leaves the pad.
static int selftest_idx()
{
    unsigned long long st = 0x9E3779B97F4A7C15ULL;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    long bad = 0;
    const int cfgs[][3] = { {30, 1, 200000}, {34, 1, 300000}, {34, 4, 300000}, {38, 8, 100000}, {56, 2, 150000}, {22, 3, 50000}, {34, 1, 3}, {34, 5, 7} };
    for (auto &cf : cfgs) {
        const int bits = cf[0]; const uint32_t ns = (uint32_t)cf[1]; const size_t n = (size_t)cf[2];
        const nd_u64 range = 1ULL << bits, key_mul = ~0ULL / range;
        std::vector<nd_u64> keys(n);
        for (auto &k : keys) { k = rnd() % range; if (cf[0] != 56) { nd_u64 m = rnd() % range; k = k < m ? k : m; m = rnd() % range; k = k < m ? k : m; } k = nd_idx_mix(k, bits); }   // skewed like minimizers (minimum of three), then mixed
        std::sort(keys.begin(), keys.end()); keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        std::vector<std::vector<nd_u64>> per(ns);
        uint32_t prev = 0;
        for (nd_u64 k : keys) { uint32_t sl = nd_idx_slice(k, key_mul, ns); if (sl < prev || sl >= ns) ++bad; prev = sl; per[sl].push_back(k); }
        size_t mx = 0; for (auto &v : per) mx = std::max(mx, v.size());
        const nd_u64 cap = ((nd_u64)(mx / 0.70) + 64 + 7) & ~7ULL, pad = 4096, slots = cap + pad;
        std::vector<nd_u64> tab((size_t)ns * slots, ND_U64_MAX);
        for (uint32_t sl = 0; sl < ns; ++sl) {
            nd_i64 run = -(1LL << 62);
            for (size_t i = 0; i < per[sl].size(); ++i) {
                nd_i64 d = (nd_i64)nd_idx_ideal(per[sl][i], key_mul, ns, cap) - (nd_i64)i;
                if (i && nd_idx_ideal(per[sl][i], key_mul, ns, cap) < nd_idx_ideal(per[sl][i - 1], key_mul, ns, cap)) ++bad;   // monotone inside a slice
                run = d > run ? d : run;
                nd_u64 slot = (nd_u64)((nd_i64)i + run);
                if (slot >= slots - 1) { ++bad; continue; }
                tab[(size_t)sl * slots + slot] = per[sl][i];
            }
        }
        auto find = [&](nd_u64 k) -> bool {
            const uint32_t sl = nd_idx_slice(k, key_mul, ns);
            size_t s = (size_t)sl * slots + nd_idx_ideal(k, key_mul, ns, cap);
            for (;; ++s) { if (tab[s] >= k) return tab[s] == k; }
        };
        for (nd_u64 k : keys) if (!find(k)) ++bad;
        for (int q = 0; q < 200000; ++q) { nd_u64 k = rnd() % range; bool in = std::binary_search(keys.begin(), keys.end(), k); if (find(k) != in) ++bad; }
    }
    printf("idx bad=%ld\n", bad);
    return bad != 0;
}

int main(int argc, char **argv)
{
    if (argc == 2 && !strcmp(argv[1], "selftest_idx")) return selftest_idx();
    if (argc == 2 && !strcmp(argv[1], "selftest_n")) return selftest_n();
    if (argc < 3) { fprintf(stderr, "usage: hd_harness reads.2bit preset [max_reads] [chunk]\n"); return 2; }
    ndo_reads *R = ndo_read_2bit(argv[1]);
    if (!R) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    ndo_opt o; ndo_opt_preset(&o, NULL);
    if (!strcmp(argv[2], "k17w10")) { ndo_opt_preset(&o, "ava-ont"); o.k = 17; o.w = 10; }
    else if (!strcmp(argv[2], "k16w7")) { ndo_opt_preset(&o, "ava-ont"); o.k = 16; o.w = 7; }
    else if (!strcmp(argv[2], "k11w1")) { ndo_opt_preset(&o, "ava-ont"); o.k = 11; o.w = 1; }
    else ndo_opt_preset(&o, argv[2]);
    int max_reads = argc > 3 ? atoi(argv[3]) : R->n;
    int chunk = argc > 4 ? atoi(argv[4]) : 256;
    if (max_reads > R->n) max_reads = R->n;
    long bad_sketch = 0, bad_sort = 0, bad_hits = 0, n_hits_tot = 0, n_mz_tot = 0, n_a_tot = 0;

    ndo_index *idx = ndo_index_build(R->n, (const char *const *)R->seq, R->len, (const char *const *)R->name, o.w, o.k, o.hpc);
    int mid_occ = ndo_index_mid_occ(idx, o.mid_occ_frac);
    fprintf(stderr, "[hd] n=%d mid_occ=%d keys=%zu\n", R->n, mid_occ, ndo_index_n_keys(idx));
    nd_chain_par par = { o.max_gap, o.max_gap, o.bw, o.max_chain_skip, o.max_chain_iter, o.min_cnt, o.min_chain_score };

    for (int r = 0; r < max_reads; ++r) {
        int len = R->len[r];
        // --- sketch, chunked ---
        std::vector<ndo_u128> ref(len + 1);
        size_t n_ref = ndo_sketch(R->seq[r], len, o.w, o.k, 0, o.hpc, ref.data(), 0);
        std::vector<uint32_t> words = pack2bit(R->seq[r], len);
        nd_seqview sv = { words.data(), nullptr, len };
        std::vector<nd_u128> got;
        int cl = chunk + (r % 7) * 13;                  // vary the chunking between reads
        for (int cs = 0; cs < len; cs += cl) {
            int ce = cs + cl < len ? cs + cl : len;
            std::vector<nd_u128> tmp(cl + o.w + 8);
            int n = (o.k > 28 ? nd_sketch_chunk<64, true>(sv, o.w, o.k, o.hpc, 0, cs, ce, tmp.data(), (int)tmp.size()) : nd_sketch_chunk<16, false>(sv, o.w, o.k, o.hpc, 0, cs, ce, tmp.data(), (int)tmp.size()));
            if (n > (int)tmp.size()) { fprintf(stderr, "overflow read %d\n", r); return 1; }
            got.insert(got.end(), tmp.begin(), tmp.begin() + n);
        }
        bool ok = got.size() == n_ref;
        for (size_t i = 0; ok && i < n_ref; ++i) ok = got[i].x == ref[i].x && got[i].y == ref[i].y;
        if (!ok) { if (bad_sketch < 5) fprintf(stderr, "sketch mismatch read %d: %zu vs %zu\n", r, got.size(), n_ref); ++bad_sketch; }
        n_mz_tot += n_ref;
        if (!o.hpc && len >= o.k + o.w - 1) {           // closed form (fast path of the GPU sketch)
            std::vector<nd_u64> xs(len); std::vector<int> zs(len); bool any_sym = false;
            for (int p = 0; p < len; ++p) { bool sy; xs[p] = nd_kmer_x(words.data(), p, o.k, &zs[p], &sy); any_sym |= sy; }
            if (!any_sym) {
                std::vector<nd_u128> cf;
                for (int t = o.k - 1; t < len; ++t)
                    if (nd_first_window_fix(xs.data(), 0, t, o.w, o.k) >= 0 ? nd_first_window_fix(xs.data(), 0, t, o.w, o.k) == 1 : nd_window_emit(xs.data(), 0, t, o.w, o.k, len)) { nd_u128 m; m.x = xs[t]; m.y = (nd_u64)(uint32_t)t << 1 | (nd_u64)zs[t]; cf.push_back(m); }
                bool okc = cf.size() == n_ref;
                for (size_t i = 0; okc && i < n_ref; ++i) okc = cf[i].x == ref[i].x && cf[i].y == ref[i].y;
                if (!okc) { if (bad_sketch < 5) fprintf(stderr, "closed-form sketch mismatch read %d: %zu vs %zu\n", r, cf.size(), n_ref); ++bad_sketch; }
            }
        }

        // --- seeds (oracle, unsorted) + exact sort ---
        int64_t n_a;
        ndo_u128 *a0 = ndo_collect_seeds(idx, &o, mid_occ, R->name[r], len, ref.data(), n_ref, &n_a, 0);
        std::vector<nd_u128> mine(n_a);
        memcpy(mine.data(), a0, n_a * 16);
        ndo_rs_sort_128x(a0, a0 + n_a);
        {
            std::vector<nd_rs_job> jobs(2 * (n_a / 65 + 1) + 2);
            uint32_t head[256], tail[256];
            nd_rs_sort_128x_seq(mine.data(), (uint32_t)n_a, jobs.data(), head, tail);
        }
        if (n_a && memcmp(mine.data(), a0, n_a * 16) != 0) { if (bad_sort < 5) fprintf(stderr, "sort mismatch read %d (n_a=%ld)\n", r, (long)n_a); ++bad_sort; }
        n_a_tot += n_a;

        // --- chain per segment ---
        std::vector<nd_hit> hits;
        {
            uint64_t sum_span = 0;
            for (int64_t i = 0; i < n_a; ++i) sum_span += mine[i].y >> 32 & 0xff;
            float avg = n_a ? (float)sum_span / n_a : 0.f;
            std::vector<int32_t> f(n_a), p(n_a), v(n_a), t(n_a), e(n_a);
            uint32_t qh = ndo_query_hash(R->name[r], len, o.seed);
            int64_t s0 = 0;
            for (int64_t i = 1; i <= n_a; ++i) {
                if (i == n_a || mine[i].x > mine[i - 1].x + (nd_u64)par.max_dist_x) {
                    int n = (int)(i - s0);
                    if (n >= par.min_cnt) {
                        nd_chain_fill_seq(mine.data() + s0, n, par, avg, f.data() + s0, p.data() + s0, v.data() + s0, t.data() + s0);
                        nd_chain_backtrack_seq(mine.data() + s0, n, par, f.data() + s0, p.data() + s0, v.data() + s0, t.data() + s0,
                                               e.data() + s0, qh, len, r, (int)s0, [&](const nd_hit &h) { hits.push_back(h); });
                    }
                    s0 = i;
                }
            }
            std::sort(hits.begin(), hits.end(), [](const nd_hit &a, const nd_hit &b) { return a.zkey > b.zkey; });
        }
        int n_regs;
        ndo_reg *regs = ndo_map_read(idx, &o, mid_occ, R->name[r], R->seq[r], len, &n_regs);
        bool hok = (int)hits.size() == n_regs;
        for (int i = 0; hok && i < n_regs; ++i) {
            const nd_hit &h = hits[i]; const ndo_reg &g = regs[i];
            hok = h.rid == g.rid && h.rev == g.rev && h.qs == g.qs && h.qe == g.qe && h.rs == g.rs && h.re == g.re &&
                  h.mlen == g.mlen && h.blen == g.blen && h.score == g.score && h.cnt == g.cnt && (uint32_t)h.zkey == g.hash;
        }
        if (!hok) { if (bad_hits < 5) fprintf(stderr, "hits mismatch read %d: %zu vs %d\n", r, hits.size(), n_regs); ++bad_hits; }
        n_hits_tot += n_regs;
        free(regs); free(a0);
    }
    printf("reads=%d minimizers=%ld anchors=%ld hits=%ld bad_sketch=%ld bad_sort=%ld bad_hits=%ld\n",
           max_reads, n_mz_tot, n_a_tot, n_hits_tot, bad_sketch, bad_sort, bad_hits);
    ndo_index_free(idx);
    ndo_reads_free(R);
    return (bad_sketch || bad_sort || bad_hits) ? 1 : 0;
}
