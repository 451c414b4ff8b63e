// nd_align.hpp -- host side of regime B (`-c`): which dynamic-programming calls to make for a chain,
// and how their results are stitched into an alignment.  The DP itself is not here: every call goes
// through a batch backend (the sm_100a kernel k_ksw_extd2_px in the product; tests may plug a CPU
// checker in).  Header-only, no CUDA.
//
// What it replaces in the reference (all "ref:" relative to /root/reference/):
//   mm_align_skeleton / mm_align1 / mm_align1_inv     minimap2/align.c:565-913
//   mm_fix_bad_ends, mm_filter_bad_seeds(_alt)        minimap2/align.c:359-492
//   mm_test_zdrop, mm_fix_cigar, mm_update_extra      minimap2/align.c:33-286
//   mm_split_reg, mm_filter_regs, mm_hit_sort         minimap2/hit.c:90-107,169-201,257-276
//   ksw_ll_qinit / ksw_ll_i16 (inversion test only)   minimap2/ksw2_ll_sse.c:31-147
//
// Execution model.  The reference walks one chain at a time and blocks on every DP call.  Here every
// chain of every read of a batch is a task; a task is advanced by re-running its (side-effect free)
// walk against a memo of finished DP calls.  A call whose result is not in the memo is queued and the
// walk continues under the assumption "no z-drop", so that one round collects the left extension, all
// gap fills and the right extension of a chain; the round's calls of all tasks go to the GPU as one
// batch.  The walk that finally finds every result it asks for is exactly the reference's sequential
// control flow, so results are identical by construction; speculation only costs discarded calls
// after a z-drop.
#pragma once
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <stdio.h>
#include <vector>
#include <algorithm>
#include "nd_core.cuh"
#ifdef _OPENMP
#include <omp.h>
#endif

namespace ndal {

#ifdef _OPENMP
inline double wall() { return omp_get_wtime(); }
#else
inline double wall() { return 0; }
#endif

static const uint64_t SEED_LONG_JOIN = 1ULL << 40, SEED_IGNORE = 1ULL << 41;     // ref: minimap2/mmpriv.h:18-21
static const uint64_t SEED_TANDEM = 1ULL << 42, SEED_SELF = 1ULL << 43;
enum { EZ_RIGHT = 0x02, EZ_APPROX_MAX = 0x08, EZ_EXTZ_ONLY = 0x40, EZ_REV_CIGAR = 0x80 };   // ref: minimap2/ksw2.h:8-14
enum { PARENT_UNSET = -1, PARENT_TMP_PRI = -2 };                                  // ref: minimap2/mmpriv.h:10-11
static const int32_t EZ_NEG_INF = -0x40000000;                                    // ref: minimap2/ksw2.h:6

struct Opt {                                       // ref: minimap2/options.c:28-62,90-97 (effective values)
    int k = 15, hpc = 0;
    int a = 2, b = 4, q = 4, e = 2, q2 = 24, e2 = 1, sc_ambi = 1;
    int zdrop = 400, zdrop_inv = 200, end_bonus = -1, min_dp_max = 80, min_ksw_len = 200;
    int bw = 2000, max_gap = 10000, min_cnt = 3, min_chain_score = 100;
    float max_clip_ratio = 1.0f;
    int no_end_flt = 0;
    long long max_sw_mat = 0;
};

// ---- sequences -------------------------------------------------------------------------------------

struct SeqStore {                                  // 2-bit words as in the .2bit container (lib/bseq.c:114-139) + N mask
    const uint32_t *words = 0; const uint64_t *word_off = 0; const uint32_t *len = 0; const uint32_t *nmask = 0;
    inline uint8_t base(int rid, uint32_t i) const
    {
        const uint64_t wo = word_off[rid];
        if (nmask && (nmask[wo / 2 + (i >> 5)] >> (i & 31) & 1)) return 4;
        return (uint8_t)(words[wo + (i >> 4)] >> (30 - 2 * (i & 15)) & 3);
    }
    void get(int rid, int st, int en, uint8_t *dst) const            // mm_idx_getseq, ref: minimap2/index.c:152-162
    {
        const uint32_t *w = words + word_off[rid];
        int i = st;
        for (; i < en && (i & 15); ++i) dst[i - st] = (uint8_t)(w[i >> 4] >> (30 - 2 * (i & 15)) & 3);
        for (; i + 16 <= en; i += 16) {
            const uint32_t x = w[i >> 4]; uint8_t *d = dst + (i - st);
            for (int k = 0; k < 16; ++k) d[k] = (uint8_t)(x >> (30 - 2 * k) & 3);
        }
        for (; i < en; ++i) dst[i - st] = (uint8_t)(w[i >> 4] >> (30 - 2 * (i & 15)) & 3);
        if (nmask) {
            const uint32_t *m = nmask + word_off[rid] / 2;
            for (i = st; i < en; ++i) if (m[i >> 5] >> (i & 31) & 1) dst[i - st] = 4;
        }
    }
    void get_reversed(int rid, int st, int en, uint8_t *dst) const   // the same followed by mm_seq_rev (align.c:25-31)
    {
        const uint32_t *w = words + word_off[rid];
        int i = st;
        for (; i < en && (i & 15); ++i) dst[en - 1 - i] = (uint8_t)(w[i >> 4] >> (30 - 2 * (i & 15)) & 3);
        for (; i + 16 <= en; i += 16) {
            const uint32_t x = w[i >> 4]; uint8_t *d = dst + (en - 1 - i);
            for (int k = 0; k < 16; ++k) d[-k] = (uint8_t)(x >> (30 - 2 * k) & 3);
        }
        for (; i < en; ++i) dst[en - 1 - i] = (uint8_t)(w[i >> 4] >> (30 - 2 * (i & 15)) & 3);
        if (nmask) {
            const uint32_t *m = nmask + word_off[rid] / 2;
            for (i = st; i < en; ++i) if (m[i >> 5] >> (i & 31) & 1) dst[en - 1 - i] = 4;
        }
    }
};

inline void encode_strands(const SeqStore &S, int rid, int qlen, uint8_t *fwd, uint8_t *rc)   // ref: align.c:866-873
{
    for (int i = 0; i < qlen; ++i) {
        uint8_t c = S.base(rid, (uint32_t)i);
        fwd[i] = c; rc[qlen - 1 - i] = c < 4 ? 3 - c : 4;
    }
}

inline void simple_mat(int8_t mat[25], const Opt &o)                // ksw_gen_simple_mat, ref: align.c:9-23
{
    int8_t a = (int8_t)(o.a < 0 ? -o.a : o.a), b = (int8_t)(o.b > 0 ? -o.b : o.b), n = (int8_t)(o.sc_ambi > 0 ? -o.sc_ambi : o.sc_ambi);
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j)
            mat[i * 5 + j] = (i == 4 || j == 4) ? n : (i == j ? a : b);
}

// ---- hits --------------------------------------------------------------------------------------------

struct Reg {                                       // the mm_reg1_t / mm_extra_t fields this path touches (minimap.h:76-99)
    int32_t id = 0, cnt = 0, rid = 0, score = 0, qs = 0, qe = 0, rs = 0, re = 0, parent = PARENT_UNSET, as = 0, mlen = 0, blen = 0;
    uint32_t hash = 0;
    uint8_t rev = 0, inv = 0, split = 0, split_inv = 0, has_p = 0;
    int32_t dp_score = 0, dp_max = 0, n_ambi = 0;
    int32_t score0 = 0;                            // chain score before any split (mm_reg1_t::score0; PAF/mapq only)
    std::vector<uint32_t> cigar;
};

inline void set_coor(Reg &r, int qlen, const nd_u128 *a)            // mm_reg_set_coor + mm_cal_fuzzy_len, ref: hit.c:8-38
{
    const nd_u128 &f = a[r.as], &l = a[r.as + r.cnt - 1];
    const int32_t span = (int32_t)(f.y >> 32 & 0xff);
    r.rev = (uint8_t)(f.x >> 63); r.rid = (int32_t)(f.x << 1 >> 33);
    r.rs = (int32_t)f.x + 1 > span ? (int32_t)f.x + 1 - span : 0;
    r.re = (int32_t)l.x + 1;
    if (!r.rev) { r.qs = (int32_t)f.y + 1 - span; r.qe = (int32_t)l.y + 1; }
    else { r.qs = qlen - ((int32_t)l.y + 1); r.qe = qlen - ((int32_t)f.y + 1 - span); }
    r.mlen = r.blen = 0;
    if (r.cnt <= 0) return;
    r.mlen = r.blen = span;
    for (int i = r.as + 1; i < r.as + r.cnt; ++i) {
        int sp = (int)(a[i].y >> 32 & 0xff), tl = (int32_t)a[i].x - (int32_t)a[i - 1].x, ql = (int32_t)a[i].y - (int32_t)a[i - 1].y;
        r.blen += tl > ql ? tl : ql;
        r.mlen += tl > sp && ql > sp ? sp : tl < ql ? tl : ql;
    }
}

inline void split_reg(Reg &r, Reg &r2, int n, int qlen, const nd_u128 *a)     // ref: hit.c:90-107
{
    if (n <= 0 || n >= r.cnt) return;
    r2 = r;
    r2.id = -1; r2.has_p = 0; r2.cigar.clear(); r2.dp_score = r2.dp_max = r2.n_ambi = 0; r2.split_inv = 0;
    r2.cnt = r.cnt - n;
    r2.score = (int32_t)(r.score * ((float)r2.cnt / r.cnt) + .499);
    r2.as = r.as + n;
    if (r.parent == r.id) r2.parent = PARENT_TMP_PRI;
    set_coor(r2, qlen, a);
    r.cnt -= r2.cnt; r.score -= r2.score;
    set_coor(r, qlen, a);
    r.split |= 1; r2.split |= 2;
}

// ---- DP calls ----------------------------------------------------------------------------------------

struct DpReq {                                     // one ksw_extd2 call, sequences given as recipes (align.c:300-313)
    int32_t task; uint32_t key;
    int32_t qlen, tlen, w, zdrop, end_bonus, flag;
    const uint8_t *qsrc; int32_t q_reversed;       // query bytes: qsrc[0..qlen), emitted back to front if reversed
    int32_t rid, t_st, t_reversed;                 // target bytes: read rid, [t_st, t_st+tlen)
};

struct DpRes {                                     // ksw_extz_t (ksw2.h:23-32); cigar lives in the task's pool
    int32_t max = 0, zdropped = 0, max_q = -1, max_t = -1, mqe = EZ_NEG_INF, mqe_t = -1, mte = EZ_NEG_INF, mte_q = -1, score = EZ_NEG_INF, n_cigar = 0, reach_end = 0;
    uint32_t cig_off = 0; int32_t zcode = -1;
};

inline void materialize(const DpReq &c, const SeqStore &T, uint8_t *qdst, uint8_t *tdst)
{
    if (!c.q_reversed) memcpy(qdst, c.qsrc, (size_t)c.qlen);
    else for (int i = 0; i < c.qlen; ++i) qdst[i] = c.qsrc[c.qlen - 1 - i];
    if (!c.t_reversed) T.get(c.rid, c.t_st, c.t_st + c.tlen, tdst);
    else T.get_reversed(c.rid, c.t_st, c.t_st + c.tlen, tdst);
}

// ---- striped local alignment score, 8 x int16 lanes (inversion test) --------------------------------
// Lane-by-lane statement of the striped kernel (ksw2_ll_sse.c:80-147) with its query profile
// (:31-78, size 2): the end positions it reports depend on the stripe order, so the stripes are kept.

struct LocalSW {
    int slen = 0, qlen = 0;
    std::vector<int16_t> prof, H0, H1, E, Hmax;
    void init(int ql, const uint8_t *query, const int8_t *mat)
    {
        qlen = ql; slen = (ql + 7) / 8;
        prof.assign((size_t)5 * slen * 8, 0);
        for (int c = 0; c < 5; ++c)
            for (int i = 0; i < slen; ++i)
                for (int l = 0; l < 8; ++l) {
                    int k = i + l * slen;
                    prof[((size_t)c * slen + i) * 8 + l] = k >= ql ? 0 : mat[c * 5 + query[k]];
                }
    }
    // All running values are in [0, 32767] (E and F never go below 0, so H = max(H_diag + s, E, F) >= 0): the unsigned saturating
    // subtractions of the striped kernel are max(x - y, 0) and its signed saturating add is min(x + y, 32767).  Lane loops are
    // written for the vectoriser.
    int run(int tlen, const uint8_t *target, int gapo, int gape, int *qe, int *te)
    {
        int gmax = 0;
        const int goe = gapo + gape, ge = gape;
        *qe = *te = -1;
        H0.assign((size_t)slen * 8, 0); H1.assign((size_t)slen * 8, 0); E.assign((size_t)slen * 8, 0); Hmax.assign((size_t)slen * 8, 0);
        int16_t *h0 = H0.data(), *h1 = H1.data(), *Ev = E.data();
        for (int i = 0; i < tlen; ++i) {
            int16_t h[8], f[8], mx[8];
            const int16_t *S = &prof[(size_t)target[i] * slen * 8];
            for (int l = 0; l < 8; ++l) { f[l] = 0; mx[l] = 0; }
            h[0] = 0; for (int l = 1; l < 8; ++l) h[l] = h0[(size_t)(slen - 1) * 8 + l - 1];
            for (int j = 0; j < slen; ++j) {
                const int16_t *Sj = S + (size_t)j * 8, *H0j = h0 + (size_t)j * 8; int16_t *Ej = Ev + (size_t)j * 8, *H1j = h1 + (size_t)j * 8;
#pragma omp simd
                for (int l = 0; l < 8; ++l) {
                    int hh = (int)h[l] + Sj[l]; hh = hh > 32767 ? 32767 : hh;
                    const int e = Ej[l];
                    hh = hh > e ? hh : e; hh = hh > f[l] ? hh : f[l];
                    mx[l] = (int16_t)(mx[l] > hh ? mx[l] : hh);
                    H1j[l] = (int16_t)hh;
                    const int h2 = hh - goe > 0 ? hh - goe : 0;
                    int e2 = e - ge > 0 ? e - ge : 0; e2 = e2 > h2 ? e2 : h2;
                    Ej[l] = (int16_t)e2;
                    int f2 = f[l] - ge > 0 ? f[l] - ge : 0; f2 = f2 > h2 ? f2 : h2;
                    f[l] = (int16_t)f2;
                    h[l] = H0j[l];
                }
            }
            bool stop = false;
            for (int k = 0; k < 8 && !stop; ++k) {
                for (int l = 7; l > 0; --l) f[l] = f[l - 1];
                f[0] = 0;
                for (int j = 0; j < slen && !stop; ++j) {
                    int16_t *H1j = h1 + (size_t)j * 8; int any = 0;
#pragma omp simd reduction(|:any)
                    for (int l = 0; l < 8; ++l) {
                        int hh = H1j[l]; hh = hh > f[l] ? hh : f[l];
                        H1j[l] = (int16_t)hh;
                        const int h2 = hh - goe > 0 ? hh - goe : 0;
                        const int f2 = f[l] - ge > 0 ? f[l] - ge : 0;
                        f[l] = (int16_t)f2;
                        any |= f2 > h2;
                    }
                    if (!any) stop = true;
                }
            }
            int imax = 0;
            for (int l = 0; l < 8; ++l) if (mx[l] > imax) imax = mx[l];
            if (imax >= gmax) { gmax = imax; *te = i; memcpy(Hmax.data(), h1, (size_t)slen * 16); }
            int16_t *t = h1; h1 = h0; h0 = t;
        }
        for (int i = 0; i < slen * 8; ++i)
            if ((int)(uint16_t)Hmax[i] == gmax) *qe = i / 8 + i % 8 * slen;
        return gmax;
    }
};

// ---- CIGAR bookkeeping -------------------------------------------------------------------------------

inline void append_cigar(Reg &r, int n, const uint32_t *cig)         // mm_append_cigar, ref: align.c:259-283
{
    if (n == 0) return;
    r.has_p = 1;
    if (!r.cigar.empty() && (r.cigar.back() & 0xf) == (cig[0] & 0xf)) {
        r.cigar.back() += cig[0] >> 4 << 4;
        r.cigar.insert(r.cigar.end(), cig + 1, cig + n);
    } else r.cigar.insert(r.cigar.end(), cig, cig + n);
}

// score along an alignment path, largest score drop against the diagonal; 0 none, 1 z-drop, 2 candidate
// inversion (ref: align.c:33-89)
inline int test_zdrop(const Opt &o, const uint8_t *qseq, const uint8_t *tseq, int n_cigar, const uint32_t *cigar, const int8_t *mat)
{
    int32_t score = 0, best = INT32_MIN, bi = -1, bj = -1, i = 0, j = 0, worst = 0;
    int box[2][2] = {{-1, -1}, {-1, -1}};
    auto visit = [&](int ii, int jj) {
        if (score < best) {
            int li = ii - bi, lj = jj - bj, diff = li > lj ? li - lj : lj - li, z = best - score - diff * o.e;
            if (z > worst) { worst = z; box[0][0] = bi; box[0][1] = ii + 1; box[1][0] = bj; box[1][1] = jj + 1; }
        } else { best = score; bi = ii; bj = jj; }
    };
    for (int k = 0; k < n_cigar; ++k) {
        const uint32_t op = cigar[k] & 0xf; const int len = (int)(cigar[k] >> 4);
        if (op == 0) {
            for (int l = 0; l < len; ++l) { score += mat[tseq[i + l] * 5 + qseq[j + l]]; visit(i + l, j + l); }
            i += len; j += len;
        } else if (op == 1 || op == 2 || op == 3) {
            score -= o.q + o.e * len;
            if (op == 1) j += len; else i += len;
            visit(i, j);
        }
    }
    const int q_len = box[1][1] - box[1][0], t_len = box[0][1] - box[0][0];
    if (worst > o.zdrop_inv && q_len < o.max_gap && t_len < o.max_gap) {
        std::vector<uint8_t> rcq((size_t)(q_len > 0 ? q_len : 0));
        for (int l = 0; l < q_len; ++l) { int c = qseq[box[1][1] - l - 1]; rcq[l] = (uint8_t)(c >= 4 ? 4 : 3 - c); }
        LocalSW sw; int qo, to;
        sw.init(q_len, rcq.data(), mat);
        int sc = sw.run(t_len, tseq + box[0][0], o.q, o.e, &qo, &to);
        if (sc >= o.min_chain_score * o.a && sc >= o.min_dp_max) return 2;
    }
    return worst > o.zdrop ? 1 : 0;
}

// left-align indels, merge I/D runs, drop a leading gap (ref: align.c:91-167)
inline void fix_cigar(Reg &r, const uint8_t *qseq, const uint8_t *tseq, int *qshift, int *tshift)
{
    std::vector<uint32_t> &c = r.cigar;
    int32_t toff = 0, qoff = 0; bool shrink = false;
    *qshift = *tshift = 0;
    if (c.size() <= 1) return;
    const uint32_t n = (uint32_t)c.size();
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t op = c[k] & 0xf; const int len = (int)(c[k] >> 4);
        if (len == 0) shrink = true;
        if (op == 0) { toff += len; qoff += len; }
        else if (op == 1 || op == 2) {
            if (k > 0 && k < n - 1 && (c[k - 1] & 0xf) == 0 && (c[k + 1] & 0xf) == 0) {
                const int prev = (int)(c[k - 1] >> 4); int l;
                const uint8_t *s = op == 1 ? qseq : tseq; const int off = op == 1 ? qoff : toff;
                for (l = 0; l < prev; ++l) if (s[off - 1 - l] != s[off + len - 1 - l]) break;
                if (l > 0) { c[k - 1] -= (uint32_t)l << 4; c[k + 1] += (uint32_t)l << 4; qoff -= l; toff -= l; }
                if (l == prev) shrink = true;
            }
            if (op == 1) qoff += len; else toff += len;
        } else if (op == 3) toff += len;
    }
    for (uint32_t k = 0; k + 2 < n; ++k) {                           // runs like 5I6D7I -> one I and one D
        if ((c[k] & 0xf) > 0 && (c[k] & 0xf) + (c[k + 1] & 0xf) == 3) {
            uint32_t l, s[3] = {0, 0, 0};
            for (l = k; l < n; ++l) {
                const uint32_t op = c[l] & 0xf;
                if (op == 1 || op == 2 || c[l] >> 4 == 0) s[op] += c[l] >> 4; else break;
            }
            if (s[1] > 0 && s[2] > 0 && l - k > 2) {
                c[k] = s[1] << 4 | 1; c[k + 1] = s[2] << 4 | 2;
                for (k += 2; k < l; ++k) c[k] &= 0xf;
                shrink = true;
            }
            k = l;
        }
    }
    if (shrink) {
        size_t l = 0;
        for (size_t k = 0; k < c.size(); ++k) if (c[k] >> 4 != 0) c[l++] = c[k];
        c.resize(l);
        l = 0;
        for (size_t k = 0; k < c.size(); ++k)
            if (k == c.size() - 1 || (c[k] & 0xf) != (c[k + 1] & 0xf)) c[l++] = c[k];
            else c[k + 1] += c[k] >> 4 << 4;
        c.resize(l);
    }
    if ((c[0] & 0xf) == 1 || (c[0] & 0xf) == 2) {
        const int l = (int)(c[0] >> 4);
        if ((c[0] & 0xf) == 1) { if (r.rev) r.qe -= l; else r.qs += l; *qshift = l; }
        else { r.rs += l; *tshift = l; }
        c.erase(c.begin());
    }
}

// recompute mlen / blen / dp_max from the final CIGAR (ref: align.c:240-286)
inline void update_extra(Reg &r, const uint8_t *qseq, const uint8_t *tseq, const int8_t *mat, int q, int e)
{
    if (!r.has_p) return;
    int qshift, tshift; int32_t s = 0, best = 0, toff = 0, qoff = 0;
    fix_cigar(r, qseq, tseq, &qshift, &tshift);
    qseq += qshift; tseq += tshift;
    r.blen = r.mlen = 0;
    for (size_t k = 0; k < r.cigar.size(); ++k) {
        const uint32_t op = r.cigar[k] & 0xf; const int len = (int)(r.cigar[k] >> 4);
        if (op == 0) {
            int n_ambi = 0, n_diff = 0;
            for (int l = 0; l < len; ++l) {
                const int cq = qseq[qoff + l], ct = tseq[toff + l];
                if (ct > 3 || cq > 3) ++n_ambi; else if (ct != cq) ++n_diff;
                s += mat[ct * 5 + cq];
                if (s < 0) s = 0; else best = best > s ? best : s;
            }
            r.blen += len - n_ambi; r.mlen += len - (n_ambi + n_diff); r.n_ambi += n_ambi;
            toff += len; qoff += len;
        } else if (op == 1 || op == 2) {
            const uint8_t *sq = op == 1 ? qseq + qoff : tseq + toff; int n_ambi = 0;
            for (int l = 0; l < len; ++l) if (sq[l] > 3) ++n_ambi;
            r.blen += len - n_ambi; r.n_ambi += n_ambi;
            s -= q + e * len; if (s < 0) s = 0;
            if (op == 1) qoff += len; else toff += len;
        } else if (op == 3) toff += len;
    }
    r.dp_max = best;
}

// ---- seed clean-up before alignment -------------------------------------------------------------------

inline int32_t wadd(int32_t x, int32_t y) { return (int32_t)((uint32_t)x + (uint32_t)y); }
inline int32_t ax(const nd_u128 &p) { return (int32_t)p.x; }
inline int32_t ay(const nd_u128 &p) { return (int32_t)p.y; }
inline int32_t aspan(const nd_u128 &p) { return (int32_t)(p.y >> 32 & 0xff); }

// trim chain ends whose seeds drift off the diagonal (ref: align.c:455-492)
inline void fix_bad_ends(const Reg &r, const nd_u128 *a, int bw, int min_match, int32_t *as, int32_t *cnt)
{
    *as = r.as; *cnt = r.cnt;
    if (r.cnt < 3) return;
    int32_t l, m; m = l = aspan(a[r.as]);
    for (int32_t i = r.as + 1; i < r.as + r.cnt - 1; ++i) {
        if (a[i].y & SEED_LONG_JOIN) break;
        const int32_t lr = ax(a[i]) - ax(a[i - 1]), lq = ay(a[i]) - ay(a[i - 1]);
        const int32_t mn = lr < lq ? lr : lq, mx = lr > lq ? lr : lq, sp = aspan(a[i]);
        if (mx - mn > l >> 1) *as = i;
        l += mn; m += mn < sp ? mn : sp;
        if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r.mlen >> 1) break;
    }
    *cnt = r.as + r.cnt - *as;
    m = l = aspan(a[r.as + r.cnt - 1]);
    for (int32_t i = r.as + r.cnt - 2; i > *as; --i) {
        if (a[i + 1].y & SEED_LONG_JOIN) break;
        const int32_t lr = ax(a[i + 1]) - ax(a[i]), lq = ay(a[i + 1]) - ay(a[i]);
        const int32_t mn = lr < lq ? lr : lq, mx = lr > lq ? lr : lq, sp = aspan(a[i + 1]);
        if (mx - mn > l >> 1) *cnt = i + 1 - *as;
        l += mn; m += mn < sp ? mn : sp;
        if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r.mlen >> 1) break;
    }
}

inline int seed_gap(const nd_u128 *a, int i)       // query step minus target step between seeds i-1 and i
{
    return (ay(a[i]) - ay(a[i - 1])) - (ax(a[i]) - ax(a[i - 1]));
}

inline void long_gaps(int as1, int cnt1, const nd_u128 *a, int min_gap, std::vector<int> &K)   // ref: align.c:359-377
{
    K.clear();
    for (int i = 1; i < cnt1; ++i) { int g = seed_gap(a, as1 + i); if (g < -min_gap || g > min_gap) K.push_back(i); }
    if (K.size() <= 1) K.clear();
}

// mark seeds between compensating indels as ignorable (ref: align.c:379-416)
inline void filter_bad_seeds(int as1, int cnt1, nd_u128 *a, int min_gap, int diff_thres, int max_ext_len, int max_ext_cnt, std::vector<int> &K)
{
    long_gaps(as1, cnt1, a, min_gap, K);
    const int n = (int)K.size();
    if (n == 0) return;
    int best = 0, best_st = -1, best_en = -1;
    for (int k = 0;; ++k) {
        if (k == n || k >= best_en) {
            if (best_en > 0) for (int i = K[best_st]; i < K[best_en]; ++i) a[as1 + i].y |= SEED_IGNORE;
            best = 0; best_st = best_en = -1;
            if (k == n) break;
        }
        int i = K[k], n_ins = 0, n_del = 0, max_diff = 0, max_diff_l = -1;
        int gap = seed_gap(a, as1 + i);
        if (gap > 0) n_ins += gap; else n_del += -gap;
        const int qs = ay(a[as1 + i - 1]), rs = ax(a[as1 + i - 1]);
        for (int l = k + 1; l < n && l <= k + max_ext_cnt; ++l) {
            const int j = K[l];
            if (ay(a[as1 + j]) - qs > max_ext_len || ax(a[as1 + j]) - rs > max_ext_len) break;
            gap = seed_gap(a, as1 + j);
            if (gap > 0) n_ins += gap; else n_del += -gap;
            const int diff = n_ins + n_del - abs(n_ins - n_del);
            if (max_diff < diff) { max_diff = diff; max_diff_l = l; }
        }
        if (max_diff > diff_thres && max_diff > best) { best = max_diff; best_st = k; best_en = max_diff_l; }
    }
}

// join seeds over clusters of long gaps (ref: align.c:418-453)
inline void filter_bad_seeds_alt(int as1, int cnt1, nd_u128 *a, int min_gap, int max_ext, std::vector<int> &K)
{
    long_gaps(as1, cnt1, a, min_gap, K);
    const int n = (int)K.size();
    for (int k = 0; k < n;) {
        const int i = K[k]; int l;
        int gap1 = seed_gap(a, as1 + i); gap1 = gap1 > 0 ? gap1 : -gap1;
        int re1 = ax(a[as1 + i]), qe1 = ay(a[as1 + i]);
        for (l = k + 1; l < n; ++l) {
            const int j = K[l];
            if (ay(a[as1 + j]) - qe1 > max_ext || ax(a[as1 + j]) - re1 > max_ext) break;
            int gap2 = seed_gap(a, as1 + j);
            const int sp = aspan(a[as1 + j - 1]), rs2 = ax(a[as1 + j - 1]) + sp, qs2 = ay(a[as1 + j - 1]) + sp;
            const int m = rs2 - re1 < qs2 - qe1 ? rs2 - re1 : qs2 - qe1;
            gap2 = gap2 > 0 ? gap2 : -gap2;
            if (m > gap1 + gap2) break;
            re1 = ax(a[as1 + j]); qe1 = ay(a[as1 + j]); gap1 = gap2;
        }
        if (l > k + 1) {
            const int end = K[l - 1];
            for (int j = K[k]; j < end; ++j) a[as1 + j].y |= SEED_IGNORE;
            a[as1 + end].y |= SEED_LONG_JOIN;
        }
        k = l;
    }
}

// ---- one read's worth of input / output --------------------------------------------------------------

struct ReadJob {
    int qid = 0, qlen = 0;                         // query read id in the query SeqStore
    std::vector<uint8_t> strands;                  // 2*qlen: forward codes, then reverse complement
    std::vector<nd_u128> a;                        // chain anchors in chain-array order (chain.c:145-160)
    std::vector<Reg> regs;                         // in mm_gen_regs order
    const uint8_t *strand(int rev) const { return strands.data() + (rev ? qlen : 0); }
};

// where a chain's DP windows start and end (ref: align.c:565-680, non-splice, non-sr branch)
struct Windows { int32_t rid, rev, as1, cnt1, rs, qs, re, qe, rs0, qs0, re0, qe0, bw; };

inline void seed_anchor_point(const Opt &o, const SeqStore &T, const uint8_t *qstrand, const nd_u128 &p, int32_t *r, int32_t *q)   // mm_adjust_minier, ref: align.c:341-357
{
    if (o.hpc) {
        int i, c;
        *q = ay(p);
        for (i = *q - 1, c = qstrand[*q]; i > 0; --i) if (qstrand[i] != c) break;
        *q = i + 1;
        const int rid = (int)(p.x << 1 >> 33); const int32_t x = ax(p);
        const int cb = T.base(rid, (uint32_t)x); int32_t t;
        for (t = x - 1; t >= 0; --t) if (T.base(rid, (uint32_t)t) != cb) break;
        *r = x + 1 - (x - t);
    } else { *r = ax(p) - (o.k >> 1); *q = ay(p) - (o.k >> 1); }
}

inline void plan_windows(const Opt &o, const SeqStore &T, ReadJob &J, const Reg &r, Windows &W, std::vector<int> &scratch)
{
    nd_u128 *a = J.a.data(); const int n_a = (int)J.a.size(), qlen = J.qlen;
    W.rid = (int32_t)(a[r.as].x << 1 >> 33); W.rev = (int32_t)(a[r.as].x >> 63);
    W.bw = (int)(o.bw * 1.5 + 1.);
    const int32_t tlen = (int32_t)T.len[W.rid];
    const uint8_t *qstrand = J.strand(W.rev);
    if (!o.no_end_flt) fix_bad_ends(r, a, o.bw, o.min_chain_score * 2, &W.as1, &W.cnt1);
    else { W.as1 = r.as; W.cnt1 = r.cnt; }
    filter_bad_seeds(W.as1, W.cnt1, a, 10, 40, o.max_gap >> 1, 10, scratch);
    filter_bad_seeds_alt(W.as1, W.cnt1, a, 30, o.max_gap >> 1, scratch);
    seed_anchor_point(o, T, qstrand, a[W.as1], &W.rs, &W.qs);
    seed_anchor_point(o, T, qstrand, a[W.as1 + W.cnt1 - 1], &W.re, &W.qe);
    const int32_t rs = W.rs, qs = W.qs, re = W.re, qe = W.qe;
    int32_t rs0, qs0, re0, qe0, rs1, qs1, re1, qe1, l, i;
    // left
    rs0 = ax(a[r.as]) + 1 - aspan(a[r.as]); qs0 = ay(a[r.as]) + 1 - aspan(a[r.as]);
    if (rs0 < 0) rs0 = 0;
    rs1 = qs1 = 0;
    for (i = r.as - 1, l = 0; i >= 0 && a[i].x >> 32 == a[r.as].x >> 32; --i) {
        const int32_t x = ax(a[i]) + 1 - aspan(a[i]), y = ay(a[i]) + 1 - aspan(a[i]);
        if (x < rs0 && y < qs0) {
            if (++l > o.min_cnt) {
                l = rs0 - x > qs0 - y ? rs0 - x : qs0 - y;
                rs1 = rs0 - l; qs1 = qs0 - l;
                if (rs1 < 0) rs1 = 0;
                break;
            }
        }
    }
    if (qs > 0 && rs > 0) {
        l = qs < o.max_gap ? qs : o.max_gap;
        qs1 = qs1 > qs - l ? qs1 : qs - l;
        qs0 = qs0 < qs1 ? qs0 : qs1;
        l += l * o.a > o.q ? (l * o.a - o.q) / o.e : 0;
        l = l < o.max_gap ? l : o.max_gap;
        l = l < rs ? l : rs;
        rs1 = rs1 > rs - l ? rs1 : rs - l;
        rs0 = rs0 < rs1 ? rs0 : rs1;
        rs0 = rs0 < rs ? rs0 : rs;
    } else { rs0 = rs; qs0 = qs; }
    // right
    re0 = ax(a[r.as + r.cnt - 1]) + 1; qe0 = ay(a[r.as + r.cnt - 1]) + 1;
    re1 = tlen; qe1 = qlen;
    for (i = r.as + r.cnt, l = 0; i < n_a && a[i].x >> 32 == a[r.as].x >> 32; ++i) {
        const int32_t x = ax(a[i]) + 1, y = ay(a[i]) + 1;
        if (x > re0 && y > qe0) {
            if (++l > o.min_cnt) {
                l = x - re0 > y - qe0 ? x - re0 : y - qe0;
                re1 = re0 + l; qe1 = qe0 + l;
                break;
            }
        }
    }
    if (qe < qlen && re < tlen) {
        l = qlen - qe < o.max_gap ? qlen - qe : o.max_gap;
        qe1 = qe1 < qe + l ? qe1 : qe + l;
        qe0 = qe0 > qe1 ? qe0 : qe1;
        l += l * o.a > o.q ? (l * o.a - o.q) / o.e : 0;
        l = l < o.max_gap ? l : o.max_gap;
        l = l < tlen - re ? l : tlen - re;
        re1 = re1 < re + l ? re1 : re + l;
        re0 = re0 > re1 ? re0 : re1;
    } else { re0 = re; qe0 = qe; }
    if (a[r.as].y & SEED_SELF) {
        int max_ext = r.qs > r.rs ? r.qs - r.rs : r.rs - r.qs;
        if (r.rs - rs0 > max_ext) rs0 = r.rs - max_ext;
        if (r.qs - qs0 > max_ext) qs0 = r.qs - max_ext;
        max_ext = r.qe > r.re ? r.qe - r.re : r.re - r.qe;
        if (re0 - r.re > max_ext) re0 = r.re + max_ext;
        if (qe0 - r.qe > max_ext) qe0 = r.qe + max_ext;
    }
    W.rs0 = rs0; W.qs0 = qs0; W.re0 = re0; W.qe0 = qe0;
}

// ---- tasks --------------------------------------------------------------------------------------------

enum { K_LEFT = 1, K_FILL1 = 2, K_FILL2 = 3, K_RIGHT = 4, K_INV = 5 };

struct Task {
    int read = 0;
    std::vector<Reg> out; int cur = 0; int phase = 0; bool done = false;
    bool planned = false; Windows W;
    std::vector<std::pair<uint32_t, DpRes> > memo; std::vector<uint32_t> cigpool;
    DpRes *find(uint32_t key) { for (auto &m : memo) if (m.first == key) return &m.second; return 0; }
    void reset_call_state() { planned = false; memo.clear(); cigpool.clear(); }
};

struct Ctx {
    const Opt *o; const SeqStore *T; int8_t mat[25];
};

struct WalkProf { double t_plan = 0, t_zdrop = 0, t_extra = 0, t_task_max = 0; long n_zdrop = 0, n_extra = 0, n_inv = 0; };
inline WalkProf *&walk_prof() { static thread_local WalkProf *p = 0; return p; }

struct Walker {                                    // one attempt at mm_align1 / mm_align1_inv for the current hit of a task
    const Ctx &C; ReadJob &J; Task &K; int task_id; std::vector<DpReq> &reqs; std::vector<uint8_t> &tbuf; std::vector<int> &scratch;
    bool missing = false;

    DpRes *call(int kind, int idx, int qlen, const uint8_t *qsrc, int q_rev, int tlen, int rid, int t_st, int t_rev, int w, int end_bonus, int zdrop, int flag)
    {
        const uint32_t key = (uint32_t)kind << 28 | (uint32_t)idx;
        DpRes *m = K.find(key);
        if (m) return m;
        const Opt &o = *C.o;
        if (qlen <= 0 || tlen <= 0 || (o.max_sw_mat > 0 && (long long)tlen * qlen > o.max_sw_mat)) {   // ref: ksw2_extd2_sse.c:44, align.c:300-303
            DpRes z; if (qlen > 0 && tlen > 0) z.zdropped = 1;
            K.memo.push_back(std::make_pair(key, z));
            return &K.memo.back().second;
        }
        DpReq q; q.task = task_id; q.key = key; q.qlen = qlen; q.tlen = tlen; q.w = w; q.zdrop = zdrop; q.end_bonus = end_bonus; q.flag = flag;
        q.qsrc = qsrc; q.q_reversed = q_rev; q.rid = rid; q.t_st = t_st; q.t_reversed = t_rev;
        reqs.push_back(q);
        missing = true;
        return 0;
    }

    // returns true when the hit is finished (r and r2 final), false when DP results are outstanding
    bool align1(Reg &r, Reg &r2)
    {
        const Opt &o = *C.o; const SeqStore &T = *C.T;
        r2.cnt = 0;
        if (r.cnt == 0) return true;
        if (!K.planned) { const double tp = walk_prof() ? wall() : 0; plan_windows(o, T, J, r, K.W, scratch); K.planned = true; if (walk_prof()) walk_prof()->t_plan += wall() - tp; }
        const Windows &W = K.W;
        const nd_u128 *a = J.a.data();
        const int qlen = J.qlen, rid = W.rid, rev = W.rev, as1 = W.as1, cnt1 = W.cnt1, bw = W.bw;
        const uint8_t *qstrand = J.strand(rev);
        int32_t rs = W.rs, qs = W.qs, re = W.re, qe = W.qe, rs1, qs1, re1, qe1;
        bool dropped = false;
        Reg w = r;                                 // committed only if nothing is missing
        Reg w2; w2.cnt = 0;
        const size_t n_req0 = reqs.size();

        if (qs > 0 && rs > 0) {                    // left extension, both sequences reversed (ref: align.c:690-704)
            DpRes *z = call(K_LEFT, 0, qs - W.qs0, qstrand + W.qs0, 1, rs - W.rs0, rid, W.rs0, 1, bw, o.end_bonus, r.split_inv ? o.zdrop_inv : o.zdrop, EZ_EXTZ_ONLY | EZ_RIGHT | EZ_REV_CIGAR);
            if (z) {
                if (z->n_cigar > 0) { append_cigar(w, z->n_cigar, K.cigpool.data() + z->cig_off); w.dp_score = wadd(w.dp_score, z->max); }
                rs1 = rs - (z->reach_end ? z->mqe_t + 1 : z->max_t + 1);
                qs1 = qs - (z->reach_end ? qs - W.qs0 : z->max_q + 1);
            } else { rs1 = rs; qs1 = qs; }
        } else { rs1 = rs; qs1 = qs; }
        re1 = rs; qe1 = qs;

        for (int i = 1; i < cnt1; ++i) {           // gap filling between selected seeds (ref: align.c:709-760)
            const nd_u128 &s = a[as1 + i];
            if ((s.y & (SEED_IGNORE | SEED_TANDEM)) && i != cnt1 - 1) continue;
            seed_anchor_point(o, T, qstrand, s, &re, &qe);
            re1 = re; qe1 = qe;
            if (i == cnt1 - 1 || (s.y & SEED_LONG_JOIN) || (qe - qs >= o.min_ksw_len && re - rs >= o.min_ksw_len)) {
                int bw1 = bw;
                if (s.y & SEED_LONG_JOIN) bw1 = qe - qs > re - rs ? qe - qs : re - rs;
                DpRes *z = call(K_FILL1, i, qe - qs, qstrand + qs, 0, re - rs, rid, rs, 0, bw1, -1, o.zdrop, EZ_APPROX_MAX);
                if (!z) { rs = re; qs = qe; continue; }
                if (z->zcode < 0) {
                    if (re > rs) {
                        if (tbuf.size() < (size_t)(re - rs) + 1) tbuf.resize((size_t)(re - rs) + 1);
                        T.get(rid, rs, re, tbuf.data());
                    }
                    const double tz = walk_prof() ? wall() : 0;
                    z->zcode = test_zdrop(o, qstrand + qs, tbuf.data(), z->n_cigar, K.cigpool.data() + z->cig_off, C.mat);
                    if (walk_prof()) { walk_prof()->t_zdrop += wall() - tz; ++walk_prof()->n_zdrop; if (z->zcode == 2) ++walk_prof()->n_inv; }
                }
                const int zcode = z->zcode;
                if (zcode != 0) {
                    z = call(K_FILL2, i, qe - qs, qstrand + qs, 0, re - rs, rid, rs, 0, bw1, -1, zcode == 2 ? o.zdrop_inv : o.zdrop, 0);
                    if (!z) { rs = re; qs = qe; continue; }
                }
                if (!missing && z->n_cigar > 0) append_cigar(w, z->n_cigar, K.cigpool.data() + z->cig_off);
                if (z->zdropped) {
                    int j;
                    for (j = i - 1; j >= 0; --j) if (ax(a[as1 + j]) <= rs + z->max_t) break;
                    dropped = true;
                    if (j < 0) j = 0;
                    w.dp_score = wadd(w.dp_score, z->max);
                    re1 = rs + (z->max_t + 1); qe1 = qs + (z->max_q + 1);
                    if (!missing && cnt1 - (j + 1) >= o.min_cnt) {
                        split_reg(w, w2, as1 + j + 1 - w.as, qlen, a);
                        if (w2.cnt > 0 && zcode == 2) w2.split_inv = 1;
                    }
                    break;
                } else w.dp_score = wadd(w.dp_score, z->score);
                rs = re; qs = qe;
            }
        }

        if (!dropped && qe < W.qe0 && re < W.re0) {  // right extension (ref: align.c:762-773)
            DpRes *z = call(K_RIGHT, 0, W.qe0 - qe, qstrand + qe, 0, W.re0 - re, rid, re, 0, bw, o.end_bonus, o.zdrop, EZ_EXTZ_ONLY);
            if (z) {
                if (!missing && z->n_cigar > 0) { append_cigar(w, z->n_cigar, K.cigpool.data() + z->cig_off); w.dp_score = wadd(w.dp_score, z->max); }
                re1 = re + (z->reach_end ? z->mqe_t + 1 : z->max_t + 1);
                qe1 = qe + (z->reach_end ? W.qe0 - qe : z->max_q + 1);
            }
        }
        if (missing || reqs.size() != n_req0) return false;

        w.rs = rs1; w.re = re1;
        if (rev) { w.qs = qlen - qe1; w.qe = qlen - qs1; } else { w.qs = qs1; w.qe = qe1; }
        if (w.has_p) {
            const double tx = walk_prof() ? wall() : 0;
            if (tbuf.size() < (size_t)(re1 - rs1) + 1) tbuf.resize((size_t)(re1 - rs1) + 1);
            T.get(rid, rs1, re1, tbuf.data());
            update_extra(w, J.strand(w.rev) + qs1, tbuf.data(), C.mat, o.q, o.e);
            if (walk_prof()) { walk_prof()->t_extra += wall() - tx; ++walk_prof()->n_extra; }
        }
        r = std::move(w); r2 = std::move(w2);
        return true;
    }

    // inversion between two halves of a split hit (ref: align.c:790-855).  ret: 0 none, 1 r_inv valid, -1 pending
    int align1_inv(const Reg &r1, const Reg &r2, Reg &ri)
    {
        const Opt &o = *C.o; const SeqStore &T = *C.T; const int qlen = J.qlen;
        if (!(r1.split & 1) || !(r2.split & 2)) return 0;
        if (r1.id != r1.parent && r1.parent != PARENT_TMP_PRI) return 0;
        if (r2.id != r2.parent && r2.parent != PARENT_TMP_PRI) return 0;
        if (r1.rid != r2.rid || r1.rev != r2.rev) return 0;
        const int ql = r1.rev ? r1.qs - r2.qe : r2.qs - r1.qe, tl = r2.rs - r1.re;
        if (ql < o.min_chain_score || ql > o.max_gap) return 0;
        if (tl < o.min_chain_score || tl > o.max_gap) return 0;
        const uint8_t *qseq = r1.rev ? J.strand(0) + r2.qe : J.strand(1) + (qlen - r2.qs);
        int q_off, t_off;
        DpRes *z = K.find((uint32_t)K_INV << 28);
        std::vector<uint8_t> tq((size_t)ql), tt((size_t)tl);
        if (!z) {
            for (int i = 0; i < ql; ++i) tq[i] = qseq[ql - 1 - i];
            T.get_reversed(r1.rid, r1.re, r2.rs, tt.data());
            LocalSW sw; sw.init(ql, tq.data(), C.mat);
            const int score = sw.run(tl, tt.data(), o.q, o.e, &q_off, &t_off);
            if (score < o.min_dp_max) return 0;
            q_off = ql - (q_off + 1); t_off = tl - (t_off + 1);
            K.W.rs = q_off; K.W.qs = t_off;        // remembered for the pass that finds the result
            z = call(K_INV, 0, ql - q_off, qseq + q_off, 0, tl - t_off, r1.rid, r1.re + t_off, 0, (int)(o.bw * 1.5), -1, o.zdrop, EZ_EXTZ_ONLY);
            if (!z) return -1;
        } else { q_off = K.W.rs; t_off = K.W.qs; }
        if (z->n_cigar == 0) return 0;
        ri = Reg();
        append_cigar(ri, z->n_cigar, K.cigpool.data() + z->cig_off);
        ri.dp_score = z->max; ri.id = -1; ri.parent = PARENT_UNSET; ri.inv = 1; ri.rev = !r1.rev; ri.rid = r1.rid;
        if (ri.rev == 0) { ri.qs = r2.qe + q_off; ri.qe = ri.qs + z->max_q + 1; }
        else { ri.qe = r2.qs - q_off; ri.qs = ri.qe - (z->max_q + 1); }
        ri.rs = r1.re + t_off; ri.re = ri.rs + z->max_t + 1;
        T.get(r1.rid, r1.re + t_off, r2.rs, tt.data());
        update_extra(ri, qseq + q_off, tt.data(), C.mat, o.q, o.e);
        return 1;
    }
};

// Advance one task as far as the memo allows (the loop body of mm_align_skeleton for one starting hit,
// ref: align.c:876-905).  Returns true when the task is finished.
inline bool advance(const Ctx &C, ReadJob &J, Task &K, int task_id, std::vector<DpReq> &reqs, std::vector<uint8_t> &tbuf, std::vector<int> &scratch)
{
    while (K.cur < (int)K.out.size()) {
        Walker wk{C, J, K, task_id, reqs, tbuf, scratch};
        if (K.phase == 0) {
            Reg r2;
            if (!wk.align1(K.out[K.cur], r2)) return false;
            K.reset_call_state();
            if (r2.cnt > 0) K.out.insert(K.out.begin() + K.cur + 1, r2);
            K.phase = 1;
        }
        if (K.phase == 1) {
            if (K.cur > 0 && K.out[K.cur].split_inv) {
                Reg ri;
                const int ret = wk.align1_inv(K.out[K.cur - 1], K.out[K.cur], ri);
                if (ret < 0) return false;
                K.reset_call_state();
                if (ret > 0) { K.out.insert(K.out.begin() + K.cur + 1, ri); ++K.cur; }
            }
            K.phase = 0; ++K.cur;
        }
    }
    K.done = true;
    return true;
}

// after all tasks of a read: mm_filter_regs + mm_hit_sort (ref: hit.c:257-276,169-201)
inline void finish_read(const Opt &o, int qlen, std::vector<Reg> &regs)
{
    size_t k = 0;
    for (size_t i = 0; i < regs.size(); ++i) {
        Reg &r = regs[i]; bool flt = false;
        if (!r.inv && r.cnt < o.min_cnt) flt = true;
        if (r.has_p) {
            if (r.mlen < o.min_chain_score) flt = true;
            else if (r.dp_max < o.min_dp_max) flt = true;
            else if (r.qs > qlen * o.max_clip_ratio && qlen - r.qe > qlen * o.max_clip_ratio) flt = true;
        }
        if (!flt) { if (k < i) regs[k] = regs[i]; ++k; }
    }
    regs.resize(k);
    const int n = (int)regs.size();
    if (n <= 1) return;
    std::vector<nd_u128> aux; aux.reserve((size_t)n);
    for (int i = 0; i < n; ++i)
        if (regs[i].inv || regs[i].cnt > 0) {
            nd_u128 z; z.x = (uint64_t)(uint32_t)(regs[i].has_p ? regs[i].dp_max : regs[i].score) << 32 | regs[i].hash; z.y = (uint64_t)i;
            aux.push_back(z);
        }
    std::vector<nd_rs_job> jobs(2 * (aux.size() / (ND_RS_SMALL + 1) + 1) + 2); uint32_t head[256], tail[256];
    nd_rs_sort_128x_seq(aux.data(), (uint32_t)aux.size(), jobs.data(), head, tail);
    std::vector<Reg> t; t.reserve(aux.size());
    for (int i = (int)aux.size() - 1; i >= 0; --i) t.push_back(regs[(size_t)aux[i].y]);
    regs.swap(t);
}

// ---- batch driver ----------------------------------------------------------------------------------------

struct Backend {                                   // runs ksw_extd2 for a list of calls; cigars[res[i].cig_off ..] holds res[i].n_cigar ops
    virtual void run(const std::vector<DpReq> &reqs, const SeqStore &T, std::vector<DpRes> &res, std::vector<uint32_t> &cigars) = 0;
    virtual ~Backend() {}
};

struct Stats { uint64_t n_tasks = 0, n_calls = 0, n_rounds = 0, cells = 0; double t_walk = 0, t_dp = 0, t_hand = 0, t_finish = 0; };

// jobs[r].regs are replaced by the aligned, filtered and ordered hits (what mm_align_skeleton returns)
inline void align_batch(const Opt &o, const SeqStore &T, std::vector<ReadJob> &jobs, Backend &be, Stats *st)
{
    Ctx C; C.o = &o; C.T = &T; simple_mat(C.mat, o);
    std::vector<Task> tasks;
    for (size_t r = 0; r < jobs.size(); ++r)
        for (size_t i = 0; i < jobs[r].regs.size(); ++i) {
            tasks.emplace_back();
            tasks.back().read = (int)r; tasks.back().out.push_back(jobs[r].regs[i]);
        }
    std::vector<int> active(tasks.size());
    for (size_t i = 0; i < tasks.size(); ++i) active[i] = (int)i;
    if (st) st->n_tasks += tasks.size();
    std::vector<DpReq> reqs; std::vector<DpRes> res; std::vector<uint32_t> cigs;
    while (!active.empty()) {
        reqs.clear();
        double t0 = wall();
        const bool dbg = getenv("NDAL_DEBUG") != 0;
        WalkProf tot; double busy_sum = 0, busy_max = 0; int n_thr = 0;
#pragma omp parallel
        {
            std::vector<DpReq> mine; std::vector<uint8_t> tbuf; std::vector<int> scratch;
            WalkProf wp; if (dbg) walk_prof() = &wp;
            double busy = 0;
#pragma omp for schedule(dynamic, 1) nowait
            for (long long k = 0; k < (long long)active.size(); ++k) {
                Task &K = tasks[(size_t)active[(size_t)k]];
                const double ta = dbg ? wall() : 0;
                advance(C, jobs[(size_t)K.read], K, active[(size_t)k], mine, tbuf, scratch);
                if (dbg) { const double d = wall() - ta; busy += d; if (d > wp.t_task_max) wp.t_task_max = d; }
            }
            walk_prof() = 0;
#pragma omp critical
            {
                reqs.insert(reqs.end(), mine.begin(), mine.end());
                if (dbg) { tot.t_plan += wp.t_plan; tot.t_zdrop += wp.t_zdrop; tot.t_extra += wp.t_extra; tot.n_zdrop += wp.n_zdrop; tot.n_extra += wp.n_extra; tot.n_inv += wp.n_inv;
                           if (wp.t_task_max > tot.t_task_max) tot.t_task_max = wp.t_task_max; busy_sum += busy; if (busy > busy_max) busy_max = busy; ++n_thr; }
            }
        }
        if (dbg) fprintf(stderr, "[ndal] walk detail: %d threads, busy sum %.2f s (max thread %.2f s, slowest task %.3f s); plan %.2f s, zdrop tests %ld in %.2f s (%ld inversion candidates), update_extra %ld in %.2f s\n",
                         n_thr, busy_sum, busy_max, tot.t_task_max, tot.t_plan, tot.n_zdrop, tot.t_zdrop, tot.n_inv, tot.n_extra, tot.t_extra);
        std::vector<int> next;
        for (size_t k = 0; k < active.size(); ++k) if (!tasks[(size_t)active[k]].done) next.push_back(active[k]);
        active.swap(next);
        if (st) st->t_walk += wall() - t0;
        if (getenv("NDAL_DEBUG")) fprintf(stderr, "[ndal] round: %zu active tasks walked in %.3f s -> %zu calls\n", active.size() + 0, wall() - t0, reqs.size());
        if (reqs.empty()) break;
        if (st) { st->n_calls += reqs.size(); ++st->n_rounds; for (const DpReq &q : reqs) st->cells += (uint64_t)q.qlen * q.tlen; }
        res.clear(); cigs.clear();
        t0 = wall();
        be.run(reqs, T, res, cigs);
        if (st) st->t_dp += wall() - t0;
        t0 = wall();
        // hand results to their tasks; the calls of one task are contiguous in reqs
        std::vector<size_t> runs;
        for (size_t i = 0; i < reqs.size(); ++i) if (i == 0 || reqs[i].task != reqs[i - 1].task) runs.push_back(i);
        runs.push_back(reqs.size());
#pragma omp parallel for schedule(dynamic, 64)
        for (long long u = 0; u < (long long)runs.size() - 1; ++u) {
            Task &K = tasks[(size_t)reqs[runs[(size_t)u]].task];
            for (size_t i = runs[(size_t)u]; i < runs[(size_t)u + 1]; ++i) {
                DpRes z = res[i];
                const uint32_t src = z.cig_off;
                z.cig_off = (uint32_t)K.cigpool.size(); z.zcode = -1;
                K.cigpool.insert(K.cigpool.end(), cigs.begin() + src, cigs.begin() + src + z.n_cigar);
                K.memo.push_back(std::make_pair(reqs[i].key, z));
            }
        }
        if (st) st->t_hand += wall() - t0;
    }
    const double t1 = wall();
    size_t t = 0;
    for (size_t r = 0; r < jobs.size(); ++r) {
        const size_t n0 = jobs[r].regs.size();
        std::vector<Reg> all;
        for (size_t i = 0; i < n0; ++i, ++t) for (Reg &x : tasks[t].out) all.push_back(std::move(x));
        jobs[r].regs.swap(all);
    }
#pragma omp parallel for schedule(dynamic, 4)
    for (long long r = 0; r < (long long)jobs.size(); ++r) finish_read(o, jobs[(size_t)r].qlen, jobs[(size_t)r].regs);
    if (st) st->t_finish += wall() - t1;
}

// chains of one read as the GPU path hands them over -> chain array + hits in the reference's order
// (the two unstable sorts of chain.c:145-160 and hit.c:62-71 replayed exactly).
struct ChainIn { uint64_t zkey; int32_t score, cnt, fend, jend; const nd_u128 *anchors; };

inline void build_read_job(ReadJob &J, std::vector<ChainIn> &ch)
{
    const int n = (int)ch.size();
    J.a.clear(); J.regs.clear();
    if (n == 0) return;
    std::sort(ch.begin(), ch.end(), [](const ChainIn &x, const ChainIn &y) {      // backtracking order: (f, j) descending, all keys distinct
        return x.fend != y.fend ? x.fend > y.fend : x.jend > y.jend; });
    std::vector<nd_u128> w((size_t)n);
    for (int i = 0; i < n; ++i) { w[i].x = ch[i].anchors[0].x; w[i].y = (uint64_t)i; }
    std::vector<nd_rs_job> jobs(2 * ((size_t)n / (ND_RS_SMALL + 1) + 1) + 2); uint32_t head[256], tail[256];
    nd_rs_sort_128x_seq(w.data(), (uint32_t)n, jobs.data(), head, tail);
    std::vector<nd_u128> z((size_t)n);
    size_t k = 0;
    for (int i = 0; i < n; ++i) {
        const ChainIn &c = ch[(size_t)w[i].y];
        z[i].x = c.zkey; z[i].y = (uint64_t)k << 32 | (uint32_t)c.cnt;
        J.a.insert(J.a.end(), c.anchors, c.anchors + c.cnt);
        k += (size_t)c.cnt;
    }
    nd_rs_sort_128x_seq(z.data(), (uint32_t)n, jobs.data(), head, tail);
    J.regs.resize((size_t)n);
    for (int i = 0; i < n; ++i) {
        const nd_u128 &e = z[(size_t)(n - 1 - i)];
        Reg &r = J.regs[i];
        r.id = i; r.parent = PARENT_UNSET; r.score = r.score0 = (int32_t)(e.x >> 32); r.hash = (uint32_t)e.x;
        r.cnt = (int32_t)e.y; r.as = (int32_t)(e.y >> 32);
        set_coor(r, J.qlen, J.a.data());
    }
}

} // namespace ndal
