// nd_ksw.cuh -- batched dual-affine banded extension DP (ksw_extd2), one warp per alignment.
// ref: minimap2/ksw2_extd2_sse.c:1-394, ksw2.h:119-176.
//
// The reference is a 16-lane int8 SSE kernel over anti-diagonals whose results depend on three
// implementation details, all reproduced here so that score, end points, z-drop
// decisions and CIGAR are bit-identical:
//   (1) a diagonal's band [st0,en0] is widened to 16-cell boundaries [st,en]; the extra cells are
//       computed from stale scores/state and feed the next diagonal (:139-152);
//   (2) the score vector is rewritten only on 16-wide unaligned blocks starting at st0, reading the
//       target / reversed query past their ends inside one zero-initialised block whose layout
//       (u v x y x2 y2 s sf qr) we keep, so the overruns land where the reference's do (:99-103,158-176);
//   (3) the exact maximum of H breaks ties in the order of the 4-lane SIMD search (:320-348).
// Mapping: lanes = consecutive target positions t of the current anti-diagonal (32 per step), state
// bytes in a per-warp global scratch that stays L1-resident across diagonals, neighbours (t-1) by
// shuffle with a carry between 32-cell steps, one backtrack byte per cell streamed to HBM, CIGAR
// walked by lane 0.  No tensor cores: this is integer DP.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>

#define NDK_NEG_INF (-0x40000000)
#define NDK_SCORE_ONLY  0x01
#define NDK_RIGHT       0x02
#define NDK_GENERIC_SC  0x04
#define NDK_APPROX_MAX  0x08
#define NDK_APPROX_DROP 0x10
#define NDK_EXTZ_ONLY   0x40
#define NDK_REV_CIGAR   0x80
#define NDK_FULL 0xffffffffu

struct ndk_call { int32_t qlen, tlen; unsigned long long q_off, t_off; int32_t w, zdrop, end_bonus, flag; unsigned long long p_off, cigar_off; };
struct ndk_res { int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar, reach_end, pad; };
struct ndk_par { int m; int8_t mat[25]; int q, e, q2, e2; };

__device__ __forceinline__ int ndk_w8(int v) { return (int)(int8_t)(uint8_t)v; }

// ref: ksw2.h:160-176, is_rot = 1
__device__ __forceinline__ bool ndk_zdrop(ndk_res &ez, int H, int r, int t, int zdrop, int e)
{
    if (H > ez.max) { ez.max = H; ez.max_t = t; ez.max_q = r - t; }
    else if (t >= ez.max_t && r - t >= ez.max_q) {
        int tl = t - ez.max_t, ql = (r - t) - ez.max_q, l = tl > ql ? tl - ql : ql - tl;
        if (zdrop >= 0 && ez.max - H > zdrop + l * e) { ez.zdropped = 1; return true; }
    }
    return false;
}

// ref: ksw2.h:119-152, is_rot = 1, min_intron_len = 0 (lane 0)
// rowtab != 0: the backtrack matrix lives in 64-row chunks, rowtab[r >> 6] = byte offset of the chunk from p
__device__ int ndk_backtrack(bool is_rev, const uint8_t *p, const int *off, const int *off_end, int n_col, int i0, int j0, uint32_t *cigar,
                             const unsigned long long *rowtab = nullptr)
{
    int n = 0, i = i0, j = j0, state = 0;
    while (i >= 0 && j >= 0) {
        int force = -1, r = i + j;
        int o = off[r];
        if (i < o) force = 2;
        if (i > off_end[r]) force = 1;
        uint32_t tmp = force < 0 ? (rowtab ? p[rowtab[r >> 6] + (size_t)(r & 63) * n_col + i - o] : p[(size_t)r * n_col + i - o]) : 0;
        if (state == 0) state = tmp & 7;
        else if (!(tmp >> (state + 2) & 1)) state = 0;
        if (state == 0) state = tmp & 7;
        if (force >= 0) state = force;
        int op = state == 0 ? 0 : (state == 1 || state == 3) ? 2 : 1;
        if (n == 0 || op != (int)(cigar[n - 1] & 0xf)) cigar[n++] = 1u << 4 | (uint32_t)op; else cigar[n - 1] += 1u << 4;
        if (op == 0) { --i; --j; } else if (op == 2) --i; else --j;
    }
    if (i >= 0) { if (n == 0 || 2 != (int)(cigar[n - 1] & 0xf)) cigar[n++] = (uint32_t)(i + 1) << 4 | 2u; else cigar[n - 1] += (uint32_t)(i + 1) << 4; }
    if (j >= 0) { if (n == 0 || 1 != (int)(cigar[n - 1] & 0xf)) cigar[n++] = (uint32_t)(j + 1) << 4 | 1u; else cigar[n - 1] += (uint32_t)(j + 1) << 4; }
    if (!is_rev) for (int k = 0; k < n >> 1; ++k) { uint32_t t = cigar[k]; cigar[k] = cigar[n - 1 - k]; cigar[n - 1 - k] = t; }
    return n;
}

#define NDK_WARPS 2

// scratch per warp slot: state_stride bytes of int8 state + H (int32 x Tmax) + off/off_end (int x 2 x rmax)
template <bool RIGHTV>
__device__ __forceinline__ void ndk_cell(int xt1, int vt1, int x2t1, int ut, int yo, int y2o, int &z, int q, int q2, int qe, int qe2, int sc_mch,
                                         int &un, int &vn, int &xn, int &yn, int &x2n, int &y2n, int &d)
{
    // int8 wrap-around of the reference (_mm_add_epi8/_mm_sub_epi8) cannot trigger while (q+e)+(q2+e2) <= 127 and |score| <= 127
    // (minimap2/options.c:204): every intermediate below then stays inside [-128,127], so plain int arithmetic is identical.
    int a = xt1 + vt1, b = yo + ut, a2 = x2t1 + vt1, b2 = y2o + ut;
    if (!RIGHTV) {
        d = a > z ? 1 : 0;  z = max(z, a);
        d = b > z ? 2 : d;  z = max(z, b);
        d = a2 > z ? 3 : d; z = max(z, a2);
        d = b2 > z ? 4 : d; z = max(z, b2);
    } else {
        d = z > a ? 0 : 1;  z = max(z, a);
        d = z > b ? d : 2;  z = max(z, b);
        d = z > a2 ? d : 3; z = max(z, a2);
        d = z > b2 ? d : 4; z = max(z, b2);
    }
    z = min(z, sc_mch);
    un = z - vt1; vn = z - ut;
    int tmp = z - q;  a -= tmp;  b -= tmp;
    tmp = z - q2;     a2 -= tmp; b2 -= tmp;
    if (!RIGHTV) {
        xn  = max(a, 0) - qe;   d |= a  > 0 ? 0x08 : 0;
        yn  = max(b, 0) - qe;   d |= b  > 0 ? 0x10 : 0;
        x2n = max(a2, 0) - qe2; d |= a2 > 0 ? 0x20 : 0;
        y2n = max(b2, 0) - qe2; d |= b2 > 0 ? 0x40 : 0;
    } else {
        xn  = max(a, 0) - qe;   d |= a  < 0 ? 0 : 0x08;
        yn  = max(b, 0) - qe;   d |= b  < 0 ? 0 : 0x10;
        x2n = max(a2, 0) - qe2; d |= a2 < 0 ? 0 : 0x20;
        y2n = max(b2, 0) - qe2; d |= b2 < 0 ? 0 : 0x40;
    }
}

__global__ void __launch_bounds__(NDK_WARPS * 32)
k_ksw_extd2(const ndk_call *calls, int n_calls, const uint8_t *seqs, ndk_par P, ndk_res *res, uint8_t *pmat, uint32_t *cigars,
            uint8_t *scratch, size_t slot_bytes, size_t state_bytes, size_t h_bytes, uint32_t *work)
{
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * NDK_WARPS + (threadIdx.x >> 5);
    uint8_t *base = scratch + (size_t)slot * slot_bytes;
    int32_t *H = reinterpret_cast<int32_t *>(base + state_bytes);
    int *off = reinterpret_cast<int *>(base + state_bytes + h_bytes);

    for (;;) {
        int ci = 0;
        if (lane == 0) ci = (int)atomicAdd(work, 1u);
        ci = __shfl_sync(NDK_FULL, ci, 0);
        if (ci >= n_calls) break;
        const ndk_call c = calls[ci];
        const int qlen = c.qlen, tlen = c.tlen, flag = c.flag, zdrop = c.zdrop;
        const uint8_t *query = seqs + c.q_off, *target = seqs + c.t_off;
        const bool with_cigar = !(flag & NDK_SCORE_ONLY), approx_max = (flag & NDK_APPROX_MAX) != 0, right = with_cigar && (flag & NDK_RIGHT);
        ndk_res ez;
        ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1; ez.max = 0; ez.score = ez.mqe = ez.mte = NDK_NEG_INF;
        ez.n_cigar = 0; ez.zdropped = 0; ez.reach_end = 0; ez.pad = 0;
        int q = P.q, e = P.e, q2 = P.q2, e2 = P.e2;
        const int qe0 = q + e;                              // the reference's scalar qe predates the swap (:58)
        const int m = P.m;
        bool run = !(m <= 1 || qlen <= 0 || tlen <= 0);
        if (q2 + e2 < q + e) { int t = q; q = q2; q2 = t; t = e; e = e2; e2 = t; }
        const int qe = q + e, qe2 = q2 + e2;
        const int sc_mch = P.mat[0], sc_mis = P.mat[1], sc_N = P.mat[m * m - 1] == 0 ? -e2 : P.mat[m * m - 1];
        int w = c.w;
        if (w < 0) w = tlen > qlen ? tlen : qlen;
        const int T = (tlen + 15) / 16 * 16;
        int n_col_ = qlen < tlen ? qlen : tlen;
        n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;
        const int n_col = n_col_ * 16;
        {
            int min_sc = P.mat[1];
            for (int t = 1; t < m * m; ++t) min_sc = min_sc < P.mat[t] ? min_sc : P.mat[t];
            if (-min_sc > 2 * (q + e)) run = false;
        }
        int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
        if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
        const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;

        // per-cell state packed in 8 bytes: x v x2 u | y y2 s - ; then the target copy sf (indices >= T alias the
        // reversed query, as in the reference's single block) and the reversed query qr
        uint2 *cell = reinterpret_cast<uint2 *>(base);
        int8_t *cb = reinterpret_cast<int8_t *>(base);
        uint8_t *sf = base + (size_t)8 * T, *qr = sf + T;
#define NDK_X(t)  cb[8 * (t) + 0]
#define NDK_V(t)  cb[8 * (t) + 1]
#define NDK_X2(t) cb[8 * (t) + 2]
#define NDK_U(t)  cb[8 * (t) + 3]
#define NDK_Y(t)  cb[8 * (t) + 4]
#define NDK_Y2(t) cb[8 * (t) + 5]
#define NDK_S(t)  cb[8 * (t) + 6]
        int *off_end = off + qlen + tlen - 1;
        uint8_t *p = pmat + c.p_off;
        uint32_t *cigar = cigars + c.cigar_off;

        if (run) {
            const int qpad = ((qlen + 15) / 16 + 1) * 16 + 32;
            {
                const unsigned b1 = (unsigned)(uint8_t)(-q - e), b2 = (unsigned)(uint8_t)(-q2 - e2);
                const uint2 init = make_uint2(b1 | b1 << 8 | b2 << 16 | b1 << 24, b1 | b2 << 8);
                for (int t = lane; t < T; t += 32) {
                    cell[t] = init;
                    sf[t] = t < tlen ? target[t] : 0;
                    if (!approx_max) H[t] = NDK_NEG_INF;
                }
            }
            for (int t = lane; t < qpad; t += 32) qr[t] = t < qlen ? query[qlen - 1 - t] : 0;
            __syncwarp();

            int last_st = -1, last_en = -1, H0 = 0, last_H0_t = 0;
            for (int r = 0; r < qlen + tlen - 1; ++r) {
                int st = 0, en = tlen - 1;
                if (st < r - qlen + 1) st = r - qlen + 1;
                if (en > r) en = r;
                if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
                if (en > (r + w) >> 1) en = (r + w) >> 1;
                if (st > en) { ez.zdropped = 1; break; }
                const int st0 = st, en0 = en;
                st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;
                unsigned carry;                                     // x | v<<8 | x2<<16 of cell st-1 on the previous diagonal
                const int vb = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
                {
                    const unsigned d1 = (unsigned)(uint8_t)(-q - e), d2 = (unsigned)(uint8_t)(-q2 - e2);
                    if (st > 0) {
                        if (st - 1 >= last_st && st - 1 <= last_en) carry = cell[st - 1].x & 0xffffffu;
                        else carry = d1 | d1 << 8 | d2 << 16;
                    } else carry = d1 | (unsigned)(uint8_t)vb << 8 | d2 << 16;
                }
                if (en >= r && lane == 0) { NDK_Y(r) = (int8_t)(-q - e); NDK_Y2(r) = (int8_t)(-q2 - e2); NDK_U(r) = (int8_t)vb; }
                __syncwarp();
                if (with_cigar && lane == 0) { off[r] = st; off_end[r] = en; }
                const uint8_t *qrr = qr + (qlen - 1 - r);
                const int sc_end = st0 + (en0 - st0 + 16) / 16 * 16;   // scores are rewritten on [st0, sc_end) (:158-176)
                const bool generic = (flag & NDK_GENERIC_SC) != 0;
                // core: 32 cells per step; neighbour (t-1) words by shuffle with a carry across steps; score fused in
                {
                    uint8_t *pr = p + (size_t)r * n_col - st;
                    for (int t0 = st; t0 <= en; t0 += 32) {
                        const int t = t0 + lane;
                        const bool act = t <= en;
                        uint2 cw = make_uint2(0, 0);
                        int z = 0;
                        if (act) {
                            cw = cell[t];
                            z = (int)(int8_t)(cw.y >> 16);                  // stale score unless rewritten this diagonal
                            if (t >= st0 && (generic ? t <= en0 : t < sc_end)) {
                                const int sq = sf[t], sr = qrr[t];
                                if (generic) z = P.mat[sq * m + sr];
                                else { z = sq == sr ? sc_mch : sc_mis; if (sq == m - 1 || sr == m - 1) z = sc_N; }
                            }
                        }
                        const int sc_keep = z;
                        unsigned nb = __shfl_up_sync(NDK_FULL, cw.x, 1);
                        if (lane == 0) nb = carry;
                        carry = __shfl_sync(NDK_FULL, cw.x, 31) & 0xffffffu;
                        const int xt1 = (int)(int8_t)nb, vt1 = (int)(int8_t)(nb >> 8), x2t1 = (int)(int8_t)(nb >> 16);
                        const int ut = (int)cw.x >> 24, yo = (int)(int8_t)cw.y, y2o = (int)(int8_t)(cw.y >> 8);
                        int un, vn, xn, yn, x2n, y2n, d;
                        if (right) ndk_cell<true>(xt1, vt1, x2t1, ut, yo, y2o, z, q, q2, qe, qe2, sc_mch, un, vn, xn, yn, x2n, y2n, d);
                        else ndk_cell<false>(xt1, vt1, x2t1, ut, yo, y2o, z, q, q2, qe, qe2, sc_mch, un, vn, xn, yn, x2n, y2n, d);
                        if (act) {
                            cell[t] = make_uint2(((unsigned)xn & 255u) | ((unsigned)vn & 255u) << 8 | ((unsigned)x2n & 255u) << 16 | (unsigned)un << 24,
                                                 ((unsigned)yn & 255u) | ((unsigned)y2n & 255u) << 8 | ((unsigned)sc_keep & 255u) << 16);
                            if (with_cigar) pr[t] = (uint8_t)d;
                        }
                    }
                    // score cells past the rounded band end: only s[] changes; beyond T the block spills into sf[] (:103,:175)
                    if (!generic && sc_end - 1 > en) {
                        const int t = en + 1 + lane;
                        int z = 0;
                        const bool on = t < sc_end;
                        if (on) {
                            const int sq = t < T ? sf[t] : qr[t - T], sr = qrr[t];
                            z = sq == sr ? sc_mch : sc_mis; if (sq == m - 1 || sr == m - 1) z = sc_N;
                        }
                        __syncwarp();
                        if (on) { if (t < T) NDK_S(t) = (int8_t)z; else sf[t - T] = (uint8_t)z; }
                    }
                }
                __syncwarp();
                if (!approx_max) {
                    int max_H, max_t;
                    if (r > 0) {
                        const int en1 = st0 + (en0 - st0) / 4 * 4;
                        const int Hlast = en0 > 0 ? H[en0 - 1] + NDK_U(en0) : H[en0] + NDK_V(en0);
                        __syncwarp();
                        long long best = ((long long)Hlast << 32) | 0xffffffffLL;          // rank 0 = en0
                        for (int t0 = st0; t0 <= en0; t0 += 32) {
                            const int t = t0 + lane;
                            if (t < en0) {
                                const int h = H[t] + NDK_V(t);
                                H[t] = h;
                                const unsigned rank = t < en1 ? 1u + ((unsigned)((t - st0) & 3) << 20) + (unsigned)((t - st0) >> 2)
                                                              : 1u + (4u << 20) + (unsigned)(t - en1);
                                const long long key = ((long long)h << 32) | (long long)(0xffffffffu - rank);
                                best = key > best ? key : best;
                            } else if (t == en0) H[t] = Hlast;
                        }
#pragma unroll
                        for (int o = 16; o; o >>= 1) { long long y2k = __shfl_xor_sync(NDK_FULL, best, o); best = y2k > best ? y2k : best; }
                        max_H = (int)(best >> 32);
                        const unsigned rank = 0xffffffffu - (unsigned)(best & 0xffffffffLL);
                        if (rank == 0) max_t = en0;
                        else if (rank - 1 < (4u << 20)) max_t = st0 + (int)((rank - 1) & 0xfffff) * 4 + (int)((rank - 1) >> 20);
                        else max_t = en1 + (int)(rank - 1 - (4u << 20));
                    } else {
                        if (lane == 0) H[0] = NDK_V(0) - qe0;
                        max_H = NDK_V(0) - qe0; max_t = 0;
                    }
                    __syncwarp();
                    if (en0 == tlen - 1) { int h = H[en0]; if (h > ez.mte) { ez.mte = h; ez.mte_q = r - en; } }
                    if (r - st0 == qlen - 1) { int h = H[st0]; if (h > ez.mqe) { ez.mqe = h; ez.mqe_t = st0; } }
                    if (ndk_zdrop(ez, max_H, r, max_t, zdrop, (int)(int8_t)e2)) break;
                    if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H[tlen - 1];
                } else {
                    if (r > 0) {
                        if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
                            int d0 = NDK_V(last_H0_t), d1 = NDK_U(last_H0_t + 1);
                            if (d0 > d1) H0 += d0; else { H0 += d1; ++last_H0_t; }
                        } else if (last_H0_t >= st0 && last_H0_t <= en0) H0 += NDK_V(last_H0_t);
                        else { ++last_H0_t; H0 += NDK_U(last_H0_t); }
                    } else { H0 = NDK_V(0) - qe0; last_H0_t = 0; }
                    if ((flag & NDK_APPROX_DROP) && ndk_zdrop(ez, H0, r, last_H0_t, zdrop, (int)(int8_t)e2)) break;
                    if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H0;
                }
                last_st = st; last_en = en;
            }
            __syncwarp();
            if (with_cigar && lane == 0) {
                const bool rev = (flag & NDK_REV_CIGAR) != 0;
                if (!ez.zdropped && !(flag & NDK_EXTZ_ONLY))
                    ez.n_cigar = ndk_backtrack(rev, p, off, off_end, n_col, tlen - 1, qlen - 1, cigar);
                else if (!ez.zdropped && (flag & NDK_EXTZ_ONLY) && ez.mqe + c.end_bonus > ez.max) {
                    ez.reach_end = 1;
                    ez.n_cigar = ndk_backtrack(rev, p, off, off_end, n_col, ez.mqe_t, qlen - 1, cigar);
                } else if (ez.max_t >= 0 && ez.max_q >= 0)
                    ez.n_cigar = ndk_backtrack(rev, p, off, off_end, n_col, ez.max_t, ez.max_q, cigar);
            }
        }
        if (lane == 0) res[ci] = ez;
        __syncwarp();
    }
}



// ------------------------------------------------------------------------------------------------
// Packed generation: two adjacent target cells per lane in s16x2 registers, DPX integer SIMD
// (VIADD.16x2, VIMNMX.S16x2 with predicate outputs, VIMNMX3.S16x2 on sm_100a).  64 cells per warp step.
// Per-pair state in the scratch: 8 words {x v x2 u | y y2 s -}, low half = even cell.  Same observable
// behaviour as above (16-cell rounding, stale scores, spill into sf[], SIMD tie order of the exact max).
// ------------------------------------------------------------------------------------------------
#define NDK_H16(t, f) h16[((t) >> 1) * 16 + (f) * 2 + ((t) & 1)]
__device__ __forceinline__ unsigned ndk_pk(int lo, int hi) { return ((unsigned)lo & 0xffffu) | ((unsigned)hi << 16); }

template <bool RIGHTV>
__device__ __forceinline__ void ndk_pair(unsigned xt1, unsigned vt1, unsigned x2t1, unsigned ut, unsigned yo, unsigned y2o, unsigned z,
                                         unsigned Q, unsigned Q2, unsigned NQE, unsigned NQE2, unsigned SCM,
                                         unsigned &un, unsigned &vn, unsigned &xn, unsigned &yn, unsigned &x2n, unsigned &y2n, unsigned &d)
{
    unsigned a = __vadd2(xt1, vt1), b = __vadd2(yo, ut), a2 = __vadd2(x2t1, vt1), b2 = __vadd2(y2o, ut);
    bool ph, pl; int dh, dl;
    if (!RIGHTV) {                                   // d = cand > z ? k : d  (strict), ksw2_extd2_sse.c:220-228
        z = __vibmax_s16x2(z, a, &ph, &pl);  dh = ph ? 0 : 1;  dl = pl ? 0 : 1;
        z = __vibmax_s16x2(z, b, &ph, &pl);  dh = ph ? dh : 2; dl = pl ? dl : 2;
        z = __vibmax_s16x2(z, a2, &ph, &pl); dh = ph ? dh : 3; dl = pl ? dl : 3;
        z = __vibmax_s16x2(z, b2, &ph, &pl); dh = ph ? dh : 4; dl = pl ? dl : 4;
    } else {                                         // d = z > cand ? d : k, :267-275
        z = __vibmax_s16x2(a, z, &ph, &pl);  dh = ph ? 1 : 0;  dl = pl ? 1 : 0;
        z = __vibmax_s16x2(b, z, &ph, &pl);  dh = ph ? 2 : dh; dl = pl ? 2 : dl;
        z = __vibmax_s16x2(a2, z, &ph, &pl); dh = ph ? 3 : dh; dl = pl ? 3 : dl;
        z = __vibmax_s16x2(b2, z, &ph, &pl); dh = ph ? 4 : dh; dl = pl ? 4 : dl;
    }
    z = __vimin3_s16x2(z, SCM, SCM);
    un = __vsub2(z, vt1); vn = __vsub2(z, ut);
    const unsigned nz = __vsub2(0u, z);
    const unsigned t1 = __vadd2(nz, Q), t2 = __vadd2(nz, Q2);
    a = __vadd2(a, t1); b = __vadd2(b, t1); a2 = __vadd2(a2, t2); b2 = __vadd2(b2, t2);
    // continuation flags without predicate outputs (arithmetic on the packed halves): relu, then
    //   left-aligned gaps:  flag = value > 0   <=> relu(value) >= 1 <=> bit 15 of relu + 0x7fff
    //   right-aligned gaps: flag = value >= 0  <=> sign bit clear
    const unsigned ra = __vimax_s16x2_relu(a, a), rb = __vimax_s16x2_relu(b, b), ra2 = __vimax_s16x2_relu(a2, a2), rb2 = __vimax_s16x2_relu(b2, b2);
    xn = __vadd2(ra, NQE); yn = __vadd2(rb, NQE); x2n = __vadd2(ra2, NQE2); y2n = __vadd2(rb2, NQE2);
    unsigned fa, fb, fa2, fb2;
    if (!RIGHTV) {
        fa = __vadd2(ra, 0x7fff7fffu) >> 15; fb = __vadd2(rb, 0x7fff7fffu) >> 15; fa2 = __vadd2(ra2, 0x7fff7fffu) >> 15; fb2 = __vadd2(rb2, 0x7fff7fffu) >> 15;
    } else {
        fa = ~a >> 15; fb = ~b >> 15; fa2 = ~a2 >> 15; fb2 = ~b2 >> 15;
    }
    const unsigned fl = ((fa & 0x00010001u) << 3) | ((fb & 0x00010001u) << 4) | ((fa2 & 0x00010001u) << 5) | ((fb2 & 0x00010001u) << 6);
    dl |= (int)(fl & 0xffu); dh |= (int)(fl >> 16 & 0xffu);
    d = (unsigned)dl | (unsigned)dh << 8;
}

__global__ void __launch_bounds__(NDK_WARPS * 32)
k_ksw_extd2_px(const ndk_call *calls, int n_calls, const uint8_t *seqs, ndk_par P, ndk_res *res, uint8_t *pmat, uint32_t *cigars,
               uint8_t *scratch, size_t slot_bytes, size_t state_bytes, size_t h_bytes, uint32_t *work,
               unsigned long long *arena_cur = nullptr, unsigned long long arena_bytes = 0, size_t off_bytes = 0)
{
    // arena_cur != 0: backtrack rows are carved from pmat[0, arena_bytes) on demand, 64 rows at a time (a z-dropped
    // extension touches a fraction of its worst-case matrix); a call that finds the arena full reports pad = 1 (retry)
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * NDK_WARPS + (threadIdx.x >> 5);
    uint8_t *base = scratch + (size_t)slot * slot_bytes;
    int32_t *H = reinterpret_cast<int32_t *>(base + state_bytes);
    int *off = reinterpret_cast<int *>(base + state_bytes + h_bytes);
    unsigned long long *rowtab = arena_cur ? reinterpret_cast<unsigned long long *>(base + state_bytes + h_bytes + off_bytes) : nullptr;

    for (;;) {
        int ci = 0;
        if (lane == 0) ci = (int)atomicAdd(work, 1u);
        ci = __shfl_sync(NDK_FULL, ci, 0);
        if (ci >= n_calls) break;
        const ndk_call c = calls[ci];
        const int qlen = c.qlen, tlen = c.tlen, flag = c.flag, zdrop = c.zdrop;
        const uint8_t *query = seqs + c.q_off, *target = seqs + c.t_off;
        const bool with_cigar = !(flag & NDK_SCORE_ONLY), approx_max = (flag & NDK_APPROX_MAX) != 0, right = with_cigar && (flag & NDK_RIGHT);
        ndk_res ez;
        ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1; ez.max = 0; ez.score = ez.mqe = ez.mte = NDK_NEG_INF;
        ez.n_cigar = 0; ez.zdropped = 0; ez.reach_end = 0; ez.pad = 0;
        int q = P.q, e = P.e, q2 = P.q2, e2 = P.e2;
        const int qe0 = q + e;
        const int m = P.m;
        bool run = !(m <= 1 || qlen <= 0 || tlen <= 0);
        if (q2 + e2 < q + e) { int t = q; q = q2; q2 = t; t = e; e = e2; e2 = t; }
        const int qe = q + e, qe2 = q2 + e2;
        const int sc_mch = P.mat[0], sc_mis = P.mat[1], sc_N = P.mat[m * m - 1] == 0 ? -e2 : P.mat[m * m - 1];
        int w = c.w;
        if (w < 0) w = tlen > qlen ? tlen : qlen;
        const int T = (tlen + 15) / 16 * 16;
        int n_col_ = qlen < tlen ? qlen : tlen;
        n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;
        const int n_col = n_col_ * 16;
        {
            int min_sc = P.mat[1];
            for (int t = 1; t < m * m; ++t) min_sc = min_sc < P.mat[t] ? min_sc : P.mat[t];
            if (-min_sc > 2 * (q + e)) run = false;
        }
        int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
        if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
        const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;

        uint4 *pairw = reinterpret_cast<uint4 *>(base);                 // 2 x uint4 per pair of cells
        int16_t *h16 = reinterpret_cast<int16_t *>(base);
        uint8_t *sf = base + (size_t)16 * T, *qr = sf + T;
        int *off_end = off + qlen + tlen - 1;
        uint8_t *p = rowtab ? pmat : pmat + c.p_off;
        uint32_t *cigar = cigars + c.cigar_off;
        const unsigned Q = ndk_pk(q, q), Q2 = ndk_pk(q2, q2), NQE = ndk_pk(-qe, -qe), NQE2 = ndk_pk(-qe2, -qe2), SCM = ndk_pk(sc_mch, sc_mch);
        bool arena_full = false;
        uint8_t *chunk = nullptr;

        if (run) {
            const int qpad = ((qlen + 15) / 16 + 1) * 16 + 32;
            {
                const unsigned i1 = ndk_pk(-q - e, -q - e), i2 = ndk_pk(-q2 - e2, -q2 - e2);
                for (int j = lane; j < T / 2; j += 32) {
                    pairw[2 * j] = make_uint4(i1, i1, i2, i1);          // x v x2 u
                    pairw[2 * j + 1] = make_uint4(i1, i2, 0u, 0u);      // y y2 s -
                }
                for (int t = lane; t < T; t += 32) {
                    sf[t] = t < tlen ? target[t] : 0;
                    if (!approx_max) H[t] = NDK_NEG_INF;
                }
            }
            for (int t = lane; t < qpad; t += 32) qr[t] = t < qlen ? query[qlen - 1 - t] : 0;
            __syncwarp();

            int last_st = -1, last_en = -1, H0 = 0, last_H0_t = 0;
            for (int r = 0; r < qlen + tlen - 1; ++r) {
                int st = 0, en = tlen - 1;
                if (st < r - qlen + 1) st = r - qlen + 1;
                if (en > r) en = r;
                if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
                if (en > (r + w) >> 1) en = (r + w) >> 1;
                if (st > en) { ez.zdropped = 1; break; }
                const int st0 = st, en0 = en;
                st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;
                const int vb = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
                unsigned cx, cv, cx2;                                   // previous-diagonal x, v, x2 of cell st-1, in the HIGH halves
                if (st > 0) {
                    if (st - 1 >= last_st && st - 1 <= last_en) { cx = (unsigned)NDK_H16(st - 1, 0) << 16; cv = (unsigned)NDK_H16(st - 1, 1) << 16; cx2 = (unsigned)NDK_H16(st - 1, 2) << 16; }
                    else { cx = (unsigned)(-q - e) << 16; cv = cx; cx2 = (unsigned)(-q2 - e2) << 16; }
                } else { cx = (unsigned)(-q - e) << 16; cv = (unsigned)vb << 16; cx2 = (unsigned)(-q2 - e2) << 16; }
                if (en >= r && lane == 0) { NDK_H16(r, 4) = (int16_t)(-q - e); NDK_H16(r, 5) = (int16_t)(-q2 - e2); NDK_H16(r, 3) = (int16_t)vb; }
                __syncwarp();
                if (with_cigar && lane == 0) { off[r] = st; off_end[r] = en; }
                const uint8_t *qrr = qr + (qlen - 1 - r);
                const int sc_end = st0 + (en0 - st0 + 16) / 16 * 16;
                const bool generic = (flag & NDK_GENERIC_SC) != 0;
                const int sc_hi = generic ? en0 + 1 : sc_end;           // scores rewritten on [st0, sc_hi)
                if (rowtab && with_cigar && (r & 63) == 0) {
                    unsigned long long o = 0;
                    if (lane == 0) { o = atomicAdd(arena_cur, 64ull * (unsigned long long)n_col); rowtab[r >> 6] = o; }
                    o = __shfl_sync(NDK_FULL, o, 0);
                    if (o + 64ull * (unsigned long long)n_col > arena_bytes) { arena_full = true; break; }
                    chunk = pmat + o;
                }
                const bool exact_h = !approx_max && r > 0;             // exact maximum: H[t] += v[t] fused into the sweep below
                const int en1 = st0 + (en0 - st0) / 4 * 4;
                long long best = (long long)0x8000000000000000ULL;
                {
                    uint8_t *pr = (rowtab ? chunk + (size_t)(r & 63) * n_col : p + (size_t)r * n_col) - st;
                    int ch = (exact_h && st > 0) ? H[st - 1] : 0;        // old H of the cell left of the sweep (for H[en0], ref :352)
                    // software pipeline: the state of the next 64 cells is requested before the current 64 are computed
                    uint4 A = make_uint4(0, 0, 0, 0), Bw = make_uint4(0, 0, 0, 0); unsigned sq2 = 0; int sr0 = 0, sr1 = 0; int2 hh = make_int2(0, 0);
                    {
                        const int t = st + 2 * lane;
                        if (t <= en) {
                            A = pairw[t]; Bw = pairw[t + 1];
                            if (t + 1 >= st0 && t < sc_hi) { sq2 = *reinterpret_cast<const uint16_t *>(sf + t); sr0 = qrr[t]; sr1 = qrr[t + 1]; }
                            if (exact_h) hh = *reinterpret_cast<const int2 *>(H + t);
                        }
                    }
                    for (int t0 = st; t0 <= en; t0 += 64) {
                        const int t = t0 + 2 * lane;
                        const bool act = t <= en;
                        uint4 An = make_uint4(0, 0, 0, 0), Bn = make_uint4(0, 0, 0, 0); unsigned sq2n = 0; int sr0n = 0, sr1n = 0; int2 hhn = make_int2(0, 0);
                        {
                            const int tn = t + 64;
                            if (tn <= en) {
                                An = pairw[tn]; Bn = pairw[tn + 1];
                                if (tn + 1 >= st0 && tn < sc_hi) { sq2n = *reinterpret_cast<const uint16_t *>(sf + tn); sr0n = qrr[tn]; sr1n = qrr[tn + 1]; }
                                if (exact_h) hhn = *reinterpret_cast<const int2 *>(H + tn);
                            }
                        }
                        unsigned z = 0;
                        if (act) {
                            int zl = (int)(int16_t)(Bw.z & 0xffffu), zh = (int)Bw.z >> 16;
                            if (t + 1 >= st0 && t < sc_hi) {
                                const int sq0 = sq2 & 255, sq1 = sq2 >> 8;
                                if (t >= st0) {
                                    if (generic) zl = P.mat[sq0 * m + sr0];
                                    else { zl = sq0 == sr0 ? sc_mch : sc_mis; if (sq0 == m - 1 || sr0 == m - 1) zl = sc_N; }
                                }
                                if (t + 1 < sc_hi) {
                                    if (generic) zh = P.mat[sq1 * m + sr1];
                                    else { zh = sq1 == sr1 ? sc_mch : sc_mis; if (sq1 == m - 1 || sr1 == m - 1) zh = sc_N; }
                                }
                            }
                            z = ndk_pk(zl, zh);
                        }
                        const unsigned zkeep = z;
                        unsigned px = __shfl_up_sync(NDK_FULL, A.x, 1), pv = __shfl_up_sync(NDK_FULL, A.y, 1), px2 = __shfl_up_sync(NDK_FULL, A.z, 1);
                        if (lane == 0) { px = cx; pv = cv; px2 = cx2; }
                        cx = __shfl_sync(NDK_FULL, A.x, 31); cv = __shfl_sync(NDK_FULL, A.y, 31); cx2 = __shfl_sync(NDK_FULL, A.z, 31);
                        const unsigned xt1 = __byte_perm(px, A.x, 0x5432), vt1 = __byte_perm(pv, A.y, 0x5432), x2t1 = __byte_perm(px2, A.z, 0x5432);
                        unsigned un, vn, xn, yn, x2n, y2n, d;
                        if (right) ndk_pair<true>(xt1, vt1, x2t1, A.w, Bw.x, Bw.y, z, Q, Q2, NQE, NQE2, SCM, un, vn, xn, yn, x2n, y2n, d);
                        else ndk_pair<false>(xt1, vt1, x2t1, A.w, Bw.x, Bw.y, z, Q, Q2, NQE, NQE2, SCM, un, vn, xn, yn, x2n, y2n, d);
                        if (act) {
                            pairw[t] = make_uint4(xn, vn, x2n, un);
                            pairw[t + 1] = make_uint4(yn, y2n, zkeep, 0u);
                            if (with_cigar) *reinterpret_cast<uint16_t *>(pr + t) = (uint16_t)d;
                        }
                        if (exact_h) {                                  // ref: ksw2_extd2_sse.c:346-371 (H update + SIMD-ordered maximum)
                            int ph = __shfl_up_sync(NDK_FULL, hh.y, 1);
                            if (lane == 0) ph = ch;
                            ch = __shfl_sync(NDK_FULL, hh.y, 31);
                            if (act && t + 1 >= st0 && t <= en0) {
                                const int v0 = (int)(int16_t)(vn & 0xffffu), v1 = (int)vn >> 16, u0 = (int)(int16_t)(un & 0xffffu), u1 = (int)un >> 16;
                                int h0 = hh.x, h1 = hh.y;
                                if (t >= st0) {
                                    if (t < en0) {
                                        h0 = hh.x + v0;
                                        const unsigned rank = t < en1 ? 1u + ((unsigned)((t - st0) & 3) << 20) + (unsigned)((t - st0) >> 2) : 1u + (4u << 20) + (unsigned)(t - en1);
                                        const long long key = ((long long)h0 << 32) | (long long)(0xffffffffu - rank);
                                        best = key > best ? key : best;
                                    } else {                            // t == en0
                                        h0 = en0 > 0 ? ph + u0 : hh.x + v0;
                                        const long long key = ((long long)h0 << 32) | 0xffffffffLL;
                                        best = key > best ? key : best;
                                    }
                                }
                                if (t + 1 <= en0) {
                                    if (t + 1 < en0) {
                                        h1 = hh.y + v1;
                                        const int tt = t + 1;
                                        const unsigned rank = tt < en1 ? 1u + ((unsigned)((tt - st0) & 3) << 20) + (unsigned)((tt - st0) >> 2) : 1u + (4u << 20) + (unsigned)(tt - en1);
                                        const long long key = ((long long)h1 << 32) | (long long)(0xffffffffu - rank);
                                        best = key > best ? key : best;
                                    } else {                            // t + 1 == en0 (> 0): H[en0] = H[en0-1] + u[en0], H[en0-1] being this lane's old low cell
                                        h1 = hh.x + u1;
                                        const long long key = ((long long)h1 << 32) | 0xffffffffLL;
                                        best = key > best ? key : best;
                                    }
                                }
                                *reinterpret_cast<int2 *>(H + t) = make_int2(h0, h1);
                            }
                        }
                        A = An; Bw = Bn; sq2 = sq2n; sr0 = sr0n; sr1 = sr1n; hh = hhn;
                    }
                    if (!generic && sc_end - 1 > en) {                  // score cells past the rounded band end (spill into sf[] beyond T)
                        const int t = en + 1 + lane;
                        int zz = 0;
                        const bool on = t < sc_end;
                        if (on) {
                            const int sq = t < T ? sf[t] : qr[t - T], sr = qrr[t];
                            zz = sq == sr ? sc_mch : sc_mis; if (sq == m - 1 || sr == m - 1) zz = sc_N;
                        }
                        __syncwarp();
                        if (on) { if (t < T) NDK_H16(t, 6) = (int16_t)zz; else sf[t - T] = (uint8_t)zz; }
                    }
                }
                __syncwarp();
                if (!approx_max) {
                    int max_H, max_t;
                    if (r > 0) {
#pragma unroll
                        for (int o = 16; o; o >>= 1) { long long y2k = __shfl_xor_sync(NDK_FULL, best, o); best = y2k > best ? y2k : best; }
                        max_H = (int)(best >> 32);
                        const unsigned rank = 0xffffffffu - (unsigned)(best & 0xffffffffLL);
                        if (rank == 0) max_t = en0;
                        else if (rank - 1 < (4u << 20)) max_t = st0 + (int)((rank - 1) & 0xfffff) * 4 + (int)((rank - 1) >> 20);
                        else max_t = en1 + (int)(rank - 1 - (4u << 20));
                    } else {
                        if (lane == 0) H[0] = NDK_H16(0, 1) - qe0;
                        max_H = NDK_H16(0, 1) - qe0; max_t = 0;
                    }
                    __syncwarp();
                    if (en0 == tlen - 1) { int h = H[en0]; if (h > ez.mte) { ez.mte = h; ez.mte_q = r - en; } }
                    if (r - st0 == qlen - 1) { int h = H[st0]; if (h > ez.mqe) { ez.mqe = h; ez.mqe_t = st0; } }
                    if (ndk_zdrop(ez, max_H, r, max_t, zdrop, (int)(int8_t)e2)) break;
                    if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H[tlen - 1];
                } else {
                    if (r > 0) {
                        if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
                            int d0 = NDK_H16(last_H0_t, 1), d1 = NDK_H16(last_H0_t + 1, 3);
                            if (d0 > d1) H0 += d0; else { H0 += d1; ++last_H0_t; }
                        } else if (last_H0_t >= st0 && last_H0_t <= en0) H0 += NDK_H16(last_H0_t, 1);
                        else { ++last_H0_t; H0 += NDK_H16(last_H0_t, 3); }
                    } else { H0 = NDK_H16(0, 1) - qe0; last_H0_t = 0; }
                    if ((flag & NDK_APPROX_DROP) && ndk_zdrop(ez, H0, r, last_H0_t, zdrop, (int)(int8_t)e2)) break;
                    if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H0;
                }
                last_st = st; last_en = en;
            }
            __syncwarp();
            if (arena_full) ez.pad = 1;
            else if (with_cigar && lane == 0) {
                const bool rev = (flag & NDK_REV_CIGAR) != 0;
                if (!ez.zdropped && !(flag & NDK_EXTZ_ONLY))
                    ez.n_cigar = ndk_backtrack(rev, p, off, off_end, n_col, tlen - 1, qlen - 1, cigar, rowtab);
                else if (!ez.zdropped && (flag & NDK_EXTZ_ONLY) && ez.mqe + c.end_bonus > ez.max) {
                    ez.reach_end = 1;
                    ez.n_cigar = ndk_backtrack(rev, p, off, off_end, n_col, ez.mqe_t, qlen - 1, cigar, rowtab);
                } else if (ez.max_t >= 0 && ez.max_q >= 0)
                    ez.n_cigar = ndk_backtrack(rev, p, off, off_end, n_col, ez.max_t, ez.max_q, cigar, rowtab);
            }
        }
        if (lane == 0) res[ci] = ez;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Cooperative generation for wide bands (extensions with w = 1.5*bw+1, long joins): the same cell code, one call per block.
// A lone warp advances a 3000-cell diagonal in 47 dependent round trips to its L2-resident state; four warps do it in 12 (8 warps
// per call were measured too: faster for a single call, 17 % slower for the full test set, where thousands of calls are in flight).
// ------------------------------------------------------------------------------------------------
#ifndef NDK_COOP_WARPS
#define NDK_COOP_WARPS 4
#endif
__global__ void __launch_bounds__(NDK_COOP_WARPS * 32)
k_ksw_extd2_coop(const ndk_call *calls, int n_calls, const uint8_t *seqs, ndk_par P, ndk_res *res, uint8_t *pmat, uint32_t *cigars,
               uint8_t *scratch, size_t slot_bytes, size_t state_bytes, size_t h_bytes, uint32_t *work,
               unsigned long long *arena_cur = nullptr, unsigned long long arena_bytes = 0, size_t off_bytes = 0)
{
    // One call per BLOCK: every anti-diagonal is cut into NDK_COOP_WARPS runs of 64-cell steps, one run per warp.  The only
    // values a run needs from its left neighbour are last diagonal's x, v, x2 (and H) of the cell just before it; they are read
    // before the first barrier, the in-place sweeps happen between the two barriers of a diagonal.
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, gtid = threadIdx.x;
    const int slot = blockIdx.x;
    __shared__ long long s_best[NDK_COOP_WARPS];
    __shared__ unsigned long long s_o;
    __shared__ int s_ci;
    uint8_t *base = scratch + (size_t)slot * slot_bytes;
    int32_t *H = reinterpret_cast<int32_t *>(base + state_bytes);
    int *off = reinterpret_cast<int *>(base + state_bytes + h_bytes);
    unsigned long long *rowtab = arena_cur ? reinterpret_cast<unsigned long long *>(base + state_bytes + h_bytes + off_bytes) : nullptr;

    for (;;) {
        if (gtid == 0) s_ci = (int)atomicAdd(work, 1u);
        __syncthreads();
        const int ci = s_ci;
        __syncthreads();
        if (ci >= n_calls) break;
        const ndk_call c = calls[ci];
        const int qlen = c.qlen, tlen = c.tlen, flag = c.flag, zdrop = c.zdrop;
        const uint8_t *query = seqs + c.q_off, *target = seqs + c.t_off;
        const bool with_cigar = !(flag & NDK_SCORE_ONLY), approx_max = (flag & NDK_APPROX_MAX) != 0, right = with_cigar && (flag & NDK_RIGHT);
        ndk_res ez;
        ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1; ez.max = 0; ez.score = ez.mqe = ez.mte = NDK_NEG_INF;
        ez.n_cigar = 0; ez.zdropped = 0; ez.reach_end = 0; ez.pad = 0;
        int q = P.q, e = P.e, q2 = P.q2, e2 = P.e2;
        const int qe0 = q + e;
        const int m = P.m;
        bool run = !(m <= 1 || qlen <= 0 || tlen <= 0);
        if (q2 + e2 < q + e) { int t = q; q = q2; q2 = t; t = e; e = e2; e2 = t; }
        const int qe = q + e, qe2 = q2 + e2;
        const int sc_mch = P.mat[0], sc_mis = P.mat[1], sc_N = P.mat[m * m - 1] == 0 ? -e2 : P.mat[m * m - 1];
        int w = c.w;
        if (w < 0) w = tlen > qlen ? tlen : qlen;
        const int T = (tlen + 15) / 16 * 16;
        int n_col_ = qlen < tlen ? qlen : tlen;
        n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;
        const int n_col = n_col_ * 16;
        {
            int min_sc = P.mat[1];
            for (int t = 1; t < m * m; ++t) min_sc = min_sc < P.mat[t] ? min_sc : P.mat[t];
            if (-min_sc > 2 * (q + e)) run = false;
        }
        int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
        if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
        const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;

        uint4 *pairw = reinterpret_cast<uint4 *>(base);                 // 2 x uint4 per pair of cells
        int16_t *h16 = reinterpret_cast<int16_t *>(base);
        uint8_t *sf = base + (size_t)16 * T, *qr = sf + T;
        int *off_end = off + qlen + tlen - 1;
        uint8_t *p = rowtab ? pmat : pmat + c.p_off;
        uint32_t *cigar = cigars + c.cigar_off;
        const unsigned Q = ndk_pk(q, q), Q2 = ndk_pk(q2, q2), NQE = ndk_pk(-qe, -qe), NQE2 = ndk_pk(-qe2, -qe2), SCM = ndk_pk(sc_mch, sc_mch);
        bool arena_full = false;
        uint8_t *chunk = nullptr;

        if (run) {
            const int qpad = ((qlen + 15) / 16 + 1) * 16 + 32;
            {
                const unsigned i1 = ndk_pk(-q - e, -q - e), i2 = ndk_pk(-q2 - e2, -q2 - e2);
                for (int j = gtid; j < T / 2; j += NDK_COOP_WARPS * 32) {
                    pairw[2 * j] = make_uint4(i1, i1, i2, i1);          // x v x2 u
                    pairw[2 * j + 1] = make_uint4(i1, i2, 0u, 0u);      // y y2 s -
                }
                for (int t = gtid; t < T; t += NDK_COOP_WARPS * 32) {
                    sf[t] = t < tlen ? target[t] : 0;
                    if (!approx_max) H[t] = NDK_NEG_INF;
                }
            }
            for (int t = gtid; t < qpad; t += NDK_COOP_WARPS * 32) qr[t] = t < qlen ? query[qlen - 1 - t] : 0;
            __syncthreads();

            int last_st = -1, last_en = -1, H0 = 0, last_H0_t = 0;
            for (int r = 0; r < qlen + tlen - 1; ++r) {
                int st = 0, en = tlen - 1;
                if (st < r - qlen + 1) st = r - qlen + 1;
                if (en > r) en = r;
                if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
                if (en > (r + w) >> 1) en = (r + w) >> 1;
                if (st > en) { ez.zdropped = 1; break; }
                const int st0 = st, en0 = en;
                st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;
                const int vb = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
                const int n_steps = (en - st) / 64 + 1, spw = (n_steps + NDK_COOP_WARPS - 1) / NDK_COOP_WARPS;
                const int lo = st + 64 * spw * wid;                     // this warp sweeps [lo, my_en]
                const int my_en = en < lo + 64 * spw - 1 ? en : lo + 64 * spw - 1;
                unsigned cx, cv, cx2;                                   // previous-diagonal x, v, x2 of cell lo-1, in the HIGH halves
                if (lo > st) {
                    if (lo <= en) { cx = (unsigned)NDK_H16(lo - 1, 0) << 16; cv = (unsigned)NDK_H16(lo - 1, 1) << 16; cx2 = (unsigned)NDK_H16(lo - 1, 2) << 16; }
                    else { cx = cv = cx2 = 0; }
                } else if (st > 0) {
                    if (st - 1 >= last_st && st - 1 <= last_en) { cx = (unsigned)NDK_H16(st - 1, 0) << 16; cv = (unsigned)NDK_H16(st - 1, 1) << 16; cx2 = (unsigned)NDK_H16(st - 1, 2) << 16; }
                    else { cx = (unsigned)(-q - e) << 16; cv = cx; cx2 = (unsigned)(-q2 - e2) << 16; }
                } else { cx = (unsigned)(-q - e) << 16; cv = (unsigned)vb << 16; cx2 = (unsigned)(-q2 - e2) << 16; }
                if (en >= r && gtid == 0) { NDK_H16(r, 4) = (int16_t)(-q - e); NDK_H16(r, 5) = (int16_t)(-q2 - e2); NDK_H16(r, 3) = (int16_t)vb; }
                if (with_cigar && gtid == 0) { off[r] = st; off_end[r] = en; }
                const bool exact_h = !approx_max && r > 0;             // exact maximum: H[t] += v[t] fused into the sweep below
                int ch = (exact_h && lo > 0 && lo <= en) ? H[lo - 1] : 0;   // old H of the cell left of this warp's run (for H[en0], ref :352)
                const uint8_t *qrr = qr + (qlen - 1 - r);
                const int sc_end = st0 + (en0 - st0 + 16) / 16 * 16;
                const bool generic = (flag & NDK_GENERIC_SC) != 0;
                const int sc_hi = generic ? en0 + 1 : sc_end;           // scores rewritten on [st0, sc_hi)
                if (rowtab && with_cigar && (r & 63) == 0 && gtid == 0) {
                    const unsigned long long o = atomicAdd(arena_cur, 64ull * (unsigned long long)n_col);
                    rowtab[r >> 6] = o; s_o = o;
                }
                __syncthreads();                                        // barrier 1: boundary values read, cell r initialised
                if (rowtab && with_cigar && (r & 63) == 0) {
                    const unsigned long long o = s_o;
                    if (o + 64ull * (unsigned long long)n_col > arena_bytes) { arena_full = true; break; }
                    chunk = pmat + o;
                }
                const int en1 = st0 + (en0 - st0) / 4 * 4;
                long long best = (long long)0x8000000000000000ULL;
                {
                    uint8_t *pr = (rowtab ? chunk + (size_t)(r & 63) * n_col : p + (size_t)r * n_col) - st;
                    // software pipeline: the state of the next 64 cells is requested before the current 64 are computed
                    uint4 A = make_uint4(0, 0, 0, 0), Bw = make_uint4(0, 0, 0, 0); unsigned sq2 = 0; int sr0 = 0, sr1 = 0; int2 hh = make_int2(0, 0);
                    {
                        const int t = lo + 2 * lane;
                        if (t <= my_en) {
                            A = pairw[t]; Bw = pairw[t + 1];
                            if (t + 1 >= st0 && t < sc_hi) { sq2 = *reinterpret_cast<const uint16_t *>(sf + t); sr0 = qrr[t]; sr1 = qrr[t + 1]; }
                            if (exact_h) hh = *reinterpret_cast<const int2 *>(H + t);
                        }
                    }
                    for (int t0 = lo; t0 <= my_en; t0 += 64) {
                        const int t = t0 + 2 * lane;
                        const bool act = t <= en;
                        uint4 An = make_uint4(0, 0, 0, 0), Bn = make_uint4(0, 0, 0, 0); unsigned sq2n = 0; int sr0n = 0, sr1n = 0; int2 hhn = make_int2(0, 0);
                        {
                            const int tn = t + 64;
                            if (tn <= my_en) {
                                An = pairw[tn]; Bn = pairw[tn + 1];
                                if (tn + 1 >= st0 && tn < sc_hi) { sq2n = *reinterpret_cast<const uint16_t *>(sf + tn); sr0n = qrr[tn]; sr1n = qrr[tn + 1]; }
                                if (exact_h) hhn = *reinterpret_cast<const int2 *>(H + tn);
                            }
                        }
                        unsigned z = 0;
                        if (act) {
                            int zl = (int)(int16_t)(Bw.z & 0xffffu), zh = (int)Bw.z >> 16;
                            if (t + 1 >= st0 && t < sc_hi) {
                                const int sq0 = sq2 & 255, sq1 = sq2 >> 8;
                                if (t >= st0) {
                                    if (generic) zl = P.mat[sq0 * m + sr0];
                                    else { zl = sq0 == sr0 ? sc_mch : sc_mis; if (sq0 == m - 1 || sr0 == m - 1) zl = sc_N; }
                                }
                                if (t + 1 < sc_hi) {
                                    if (generic) zh = P.mat[sq1 * m + sr1];
                                    else { zh = sq1 == sr1 ? sc_mch : sc_mis; if (sq1 == m - 1 || sr1 == m - 1) zh = sc_N; }
                                }
                            }
                            z = ndk_pk(zl, zh);
                        }
                        const unsigned zkeep = z;
                        unsigned px = __shfl_up_sync(NDK_FULL, A.x, 1), pv = __shfl_up_sync(NDK_FULL, A.y, 1), px2 = __shfl_up_sync(NDK_FULL, A.z, 1);
                        if (lane == 0) { px = cx; pv = cv; px2 = cx2; }
                        cx = __shfl_sync(NDK_FULL, A.x, 31); cv = __shfl_sync(NDK_FULL, A.y, 31); cx2 = __shfl_sync(NDK_FULL, A.z, 31);
                        const unsigned xt1 = __byte_perm(px, A.x, 0x5432), vt1 = __byte_perm(pv, A.y, 0x5432), x2t1 = __byte_perm(px2, A.z, 0x5432);
                        unsigned un, vn, xn, yn, x2n, y2n, d;
                        if (right) ndk_pair<true>(xt1, vt1, x2t1, A.w, Bw.x, Bw.y, z, Q, Q2, NQE, NQE2, SCM, un, vn, xn, yn, x2n, y2n, d);
                        else ndk_pair<false>(xt1, vt1, x2t1, A.w, Bw.x, Bw.y, z, Q, Q2, NQE, NQE2, SCM, un, vn, xn, yn, x2n, y2n, d);
                        if (act) {
                            pairw[t] = make_uint4(xn, vn, x2n, un);
                            pairw[t + 1] = make_uint4(yn, y2n, zkeep, 0u);
                            if (with_cigar) *reinterpret_cast<uint16_t *>(pr + t) = (uint16_t)d;
                        }
                        if (exact_h) {                                  // ref: ksw2_extd2_sse.c:346-371 (H update + SIMD-ordered maximum)
                            int ph = __shfl_up_sync(NDK_FULL, hh.y, 1);
                            if (lane == 0) ph = ch;
                            ch = __shfl_sync(NDK_FULL, hh.y, 31);
                            if (act && t + 1 >= st0 && t <= en0) {
                                const int v0 = (int)(int16_t)(vn & 0xffffu), v1 = (int)vn >> 16, u0 = (int)(int16_t)(un & 0xffffu), u1 = (int)un >> 16;
                                int h0 = hh.x, h1 = hh.y;
                                if (t >= st0) {
                                    if (t < en0) {
                                        h0 = hh.x + v0;
                                        const unsigned rank = t < en1 ? 1u + ((unsigned)((t - st0) & 3) << 20) + (unsigned)((t - st0) >> 2) : 1u + (4u << 20) + (unsigned)(t - en1);
                                        const long long key = ((long long)h0 << 32) | (long long)(0xffffffffu - rank);
                                        best = key > best ? key : best;
                                    } else {                            // t == en0
                                        h0 = en0 > 0 ? ph + u0 : hh.x + v0;
                                        const long long key = ((long long)h0 << 32) | 0xffffffffLL;
                                        best = key > best ? key : best;
                                    }
                                }
                                if (t + 1 <= en0) {
                                    if (t + 1 < en0) {
                                        h1 = hh.y + v1;
                                        const int tt = t + 1;
                                        const unsigned rank = tt < en1 ? 1u + ((unsigned)((tt - st0) & 3) << 20) + (unsigned)((tt - st0) >> 2) : 1u + (4u << 20) + (unsigned)(tt - en1);
                                        const long long key = ((long long)h1 << 32) | (long long)(0xffffffffu - rank);
                                        best = key > best ? key : best;
                                    } else {                            // t + 1 == en0 (> 0): H[en0] = H[en0-1] + u[en0], H[en0-1] being this lane's old low cell
                                        h1 = hh.x + u1;
                                        const long long key = ((long long)h1 << 32) | 0xffffffffLL;
                                        best = key > best ? key : best;
                                    }
                                }
                                *reinterpret_cast<int2 *>(H + t) = make_int2(h0, h1);
                            }
                        }
                        A = An; Bw = Bn; sq2 = sq2n; sr0 = sr0n; sr1 = sr1n; hh = hhn;
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) { long long y2k = __shfl_xor_sync(NDK_FULL, best, o); best = y2k > best ? y2k : best; }
                    if (lane == 0) s_best[wid] = best;
                    __syncthreads();                                    // barrier 2: every run of this diagonal is written
                    if (wid == 0 && !generic && sc_end - 1 > en) {      // score cells past the rounded band end (spill into sf[] beyond T)
                        const int t = en + 1 + lane;
                        int zz = 0;
                        const bool on = t < sc_end;
                        if (on) {
                            const int sq = t < T ? sf[t] : qr[t - T], sr = qrr[t];
                            zz = sq == sr ? sc_mch : sc_mis; if (sq == m - 1 || sr == m - 1) zz = sc_N;
                        }
                        __syncwarp();
                        if (on) { if (t < T) NDK_H16(t, 6) = (int16_t)zz; else sf[t - T] = (uint8_t)zz; }
                    }
                }
                if (!approx_max) {
                    int max_H, max_t;
                    if (r > 0) {
                        best = s_best[0];
#pragma unroll
                        for (int o = 1; o < NDK_COOP_WARPS; ++o) best = s_best[o] > best ? s_best[o] : best;
                        max_H = (int)(best >> 32);
                        const unsigned rank = 0xffffffffu - (unsigned)(best & 0xffffffffLL);
                        if (rank == 0) max_t = en0;
                        else if (rank - 1 < (4u << 20)) max_t = st0 + (int)((rank - 1) & 0xfffff) * 4 + (int)((rank - 1) >> 20);
                        else max_t = en1 + (int)(rank - 1 - (4u << 20));
                    } else {
                        if (gtid == 0) H[0] = NDK_H16(0, 1) - qe0;
                        max_H = NDK_H16(0, 1) - qe0; max_t = 0;
                    }
                    if (en0 == tlen - 1) { int h = r == 0 ? max_H : H[en0]; if (h > ez.mte) { ez.mte = h; ez.mte_q = r - en; } }
                    if (r - st0 == qlen - 1) { int h = r == 0 ? max_H : H[st0]; if (h > ez.mqe) { ez.mqe = h; ez.mqe_t = st0; } }
                    if (ndk_zdrop(ez, max_H, r, max_t, zdrop, (int)(int8_t)e2)) break;
                    if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = r == 0 ? max_H : H[tlen - 1];
                } else {
                    if (r > 0) {
                        if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
                            int d0 = NDK_H16(last_H0_t, 1), d1 = NDK_H16(last_H0_t + 1, 3);
                            if (d0 > d1) H0 += d0; else { H0 += d1; ++last_H0_t; }
                        } else if (last_H0_t >= st0 && last_H0_t <= en0) H0 += NDK_H16(last_H0_t, 1);
                        else { ++last_H0_t; H0 += NDK_H16(last_H0_t, 3); }
                    } else { H0 = NDK_H16(0, 1) - qe0; last_H0_t = 0; }
                    if ((flag & NDK_APPROX_DROP) && ndk_zdrop(ez, H0, r, last_H0_t, zdrop, (int)(int8_t)e2)) break;
                    if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H0;
                }
                last_st = st; last_en = en;
            }
            __syncthreads();
            if (arena_full) ez.pad = 1;
            else if (with_cigar && gtid == 0) {
                const bool rev = (flag & NDK_REV_CIGAR) != 0;
                if (!ez.zdropped && !(flag & NDK_EXTZ_ONLY))
                    ez.n_cigar = ndk_backtrack(rev, p, off, off_end, n_col, tlen - 1, qlen - 1, cigar, rowtab);
                else if (!ez.zdropped && (flag & NDK_EXTZ_ONLY) && ez.mqe + c.end_bonus > ez.max) {
                    ez.reach_end = 1;
                    ez.n_cigar = ndk_backtrack(rev, p, off, off_end, n_col, ez.mqe_t, qlen - 1, cigar, rowtab);
                } else if (ez.max_t >= 0 && ez.max_q >= 0)
                    ez.n_cigar = ndk_backtrack(rev, p, off, off_end, n_col, ez.max_t, ez.max_q, cigar, rowtab);
            }
        }
        if (gtid == 0) res[ci] = ez;
        __syncthreads();
    }
}

// ---- compact CIGAR hand-back: per-call capacity buffers -> one dense array ---------------------------------
__global__ void k_ksw_ncigar(const ndk_res *res, int n, uint32_t *ncig)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ncig[i] = (uint32_t)res[i].n_cigar;
    if (i == n) ncig[i] = 0;
}
__global__ void k_ksw_cigar_gather(const ndk_call *calls, const ndk_res *res, const unsigned long long *dst_off, int n, const uint32_t *cigars, uint32_t *out)
{
    const int i = (int)((blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint32_t *src = cigars + calls[i].cigar_off; uint32_t *dst = out + dst_off[i];
    for (int k = lane; k < res[i].n_cigar; k += 32) dst[k] = src[k];
}
