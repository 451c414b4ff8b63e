/*
 * nd_ksw.c -- scalar CPU restatement of the reference's dual-affine banded extension DP
 * (ksw_extd2_sse, ref: minimap2/ksw2_extd2_sse.c:1-394; helpers ksw2.h:119-176).
 *
 * TEST INFRASTRUCTURE ONLY (same rules as nd_oracle.c).  Pinned by tests/test_ksw_oracle.py against
 * the unmodified reference function through oracle/_ref/libminimap2_ref.so.
 *
 * The reference is a 16-lane int8 SSE kernel over anti-diagonals; three of its properties are
 * observable and restated literally here:
 *   (1) the band [st0,en0] of a diagonal is widened to 16-lane boundaries [st,en] and those extra
 *       cells ARE computed (from stale scores and stale state) and feed the next diagonal;
 *   (2) the score vector s[] is only rewritten on [st0, st0+16*ceil) each diagonal, from the target and
 *       the reversed query padded with zeros (kcalloc), and keeps older values elsewhere;
 *   (3) the exact-max search breaks ties in the order of its 4-lane SIMD loop.
 * All state is int8 with wrap-around, as _mm_add_epi8/_mm_sub_epi8.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NDK_NEG_INF (-0x40000000)
#define NDK_SCORE_ONLY  0x01
#define NDK_RIGHT       0x02
#define NDK_GENERIC_SC  0x04
#define NDK_APPROX_MAX  0x08
#define NDK_APPROX_DROP 0x10
#define NDK_EXTZ_ONLY   0x40
#define NDK_REV_CIGAR   0x80

typedef struct {
    int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar, reach_end;
} ndk_result;

static inline int8_t w8(int v) { return (int8_t)(uint8_t)v; }

/* ref: ksw2.h:160-176 (ksw_apply_zdrop), is_rot = 1 */
static int apply_zdrop(ndk_result *ez, int32_t H, int r, int t, int zdrop, int8_t e)
{
    if (H > ez->max) { ez->max = H; ez->max_t = t; ez->max_q = r - t; }
    else if (t >= ez->max_t && r - t >= ez->max_q) {
        int tl = t - ez->max_t, ql = (r - t) - ez->max_q, l = tl > ql ? tl - ql : ql - tl;
        if (zdrop >= 0 && ez->max - H > zdrop + l * e) { ez->zdropped = 1; return 1; }
    }
    return 0;
}

/* ref: ksw2.h:119-152 (ksw_backtrack), is_rot = 1, min_intron_len = 0.  cigar[] must hold qlen+tlen entries */
static int backtrack(int is_rev, const uint8_t *p, const int *off, const int *off_end, int n_col, int i0, int j0, uint32_t *cigar)
{
    int n = 0, i = i0, j = j0, state = 0, k;
#define PUSH(op, len) do { if (n == 0 || (op) != (int)(cigar[n - 1] & 0xf)) cigar[n++] = (uint32_t)(len) << 4 | (op); else cigar[n - 1] += (uint32_t)(len) << 4; } while (0)
    while (i >= 0 && j >= 0) {
        int force = -1, r = i + j;
        uint32_t tmp;
        if (i < off[r]) force = 2;
        if (i > off_end[r]) force = 1;
        tmp = force < 0 ? p[(size_t)r * n_col + i - off[r]] : 0;
        if (state == 0) state = tmp & 7;
        else if (!(tmp >> (state + 2) & 1)) state = 0;
        if (state == 0) state = tmp & 7;
        if (force >= 0) state = force;
        if (state == 0) { PUSH(0, 1); --i; --j; }
        else if (state == 1 || state == 3) { PUSH(2, 1); --i; }
        else { PUSH(1, 1); --j; }
    }
    if (i >= 0) PUSH(2, i + 1);
    if (j >= 0) PUSH(1, j + 1);
#undef PUSH
    if (!is_rev) for (k = 0; k < n >> 1; ++k) { uint32_t t = cigar[k]; cigar[k] = cigar[n - 1 - k]; cigar[n - 1 - k] = t; }
    return n;
}

/* returns 0; fills *ez and cigar (capacity qlen+tlen+2) */
int ndk_extd2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
              int q, int e, int q2, int e2, int w, int zdrop, int end_bonus, int flag, ndk_result *ez, uint32_t *cigar)
{
    int r, t, qe, qe2, qe0 = q + e /* scalar qe is taken before the swap (:58) */, n_col_, tlen_, T, last_st, last_en, wl, wr, min_sc, long_thres, long_diff;
    const int with_cigar = !(flag & NDK_SCORE_ONLY), approx_max = !!(flag & NDK_APPROX_MAX);
    int32_t *H = NULL, H0 = 0, last_H0_t = 0;
    int8_t *u, *v, *x, *y, *x2, *y2, *s, sc_mch, sc_mis, sc_N;
    uint8_t *sf, *qr, *p = NULL;
    int *off = NULL, *off_end = NULL;

    ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;           /* ksw_reset_extz, ksw2.h:153-158 */
    ez->max = 0; ez->score = ez->mqe = ez->mte = NDK_NEG_INF;
    ez->n_cigar = 0; ez->zdropped = 0; ez->reach_end = 0;
    if (m <= 1 || qlen <= 0 || tlen <= 0) return 0;
    if (q2 + e2 < q + e) { t = q; q = q2; q2 = t; t = e; e = e2; e2 = t; }      /* :69 */
    qe = q + e; qe2 = q2 + e2;
    sc_mch = mat[0]; sc_mis = mat[1];
    sc_N = mat[m * m - 1] == 0 ? (int8_t)-e2 : mat[m * m - 1];
    if (w < 0) w = tlen > qlen ? tlen : qlen;
    wl = wr = w;
    tlen_ = (tlen + 15) / 16; T = tlen_ * 16;
    n_col_ = qlen < tlen ? qlen : tlen;
    n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;
    for (t = 1, min_sc = mat[1]; t < m * m; ++t) min_sc = min_sc < mat[t] ? min_sc : mat[t];
    if (-min_sc > 2 * (q + e)) return 0;                                            /* :93 */
    long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
    if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
    long_diff = long_thres * (e - e2) - (q2 - q) - e2;

    /* one zeroed block laid out exactly as the reference's (:99-103): u v x y x2 y2 s sf qr.  The 16-wide
     * unaligned score blocks may write up to 15 bytes past s[] into sf[] and read sf[]/qr[] past their
     * ends; with this layout those accesses land where the reference's do. */
    u = (int8_t*)calloc(8 * (size_t)T + ((size_t)(qlen + 15) / 16 + 1) * 16 + 64, 1);
    v = u + T; x = v + T; y = x + T; x2 = y + T; y2 = x2 + T; s = y2 + T;
    sf = (uint8_t*)(s + T); qr = sf + T;
    memset(u, w8(-q - e), 4 * (size_t)T); memset(x2, w8(-q2 - e2), 2 * (size_t)T);
    memcpy(sf, target, tlen);
    for (t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];
    if (!approx_max) { H = (int32_t*)malloc((size_t)T * 4); for (t = 0; t < T; ++t) H[t] = NDK_NEG_INF; }
    if (with_cigar) {
        p = (uint8_t*)malloc(((size_t)(qlen + tlen - 1) * n_col_ + 1) * 16);
        off = (int*)malloc((size_t)(qlen + tlen - 1) * sizeof(int) * 2);
        off_end = off + qlen + tlen - 1;
    }

    for (r = 0, last_st = last_en = -1; r < qlen + tlen - 1; ++r) {
        int st = 0, en = tlen - 1, st0, en0;
        int8_t x1, x21, v1;
        const uint8_t *qrr = qr + (qlen - 1 - r);
        if (st < r - qlen + 1) st = r - qlen + 1;
        if (en > r) en = r;
        if (st < (r - wr + 1) >> 1) st = (r - wr + 1) >> 1;
        if (en > (r + wl) >> 1) en = (r + wl) >> 1;
        if (st > en) { ez->zdropped = 1; break; }
        st0 = st; en0 = en;
        st = st / 16 * 16; en = (en + 16) / 16 * 16 - 1;
        if (st > 0) {                                                  /* :142-152 */
            if (st - 1 >= last_st && st - 1 <= last_en) { x1 = x[st - 1]; x21 = x2[st - 1]; v1 = v[st - 1]; }
            else { x1 = w8(-q - e); x21 = w8(-q2 - e2); v1 = w8(-q - e); }
        } else {
            x1 = w8(-q - e); x21 = w8(-q2 - e2);
            v1 = r == 0 ? w8(-q - e) : r < long_thres ? w8(-e) : r == long_thres ? w8(long_diff) : w8(-e2);
        }
        if (en >= r) {                                                 /* :153-156 */
            y[r] = w8(-q - e); y2[r] = w8(-q2 - e2);
            u[r] = r == 0 ? w8(-q - e) : r < long_thres ? w8(-e) : r == long_thres ? w8(long_diff) : w8(-e2);
        }
        if (!(flag & NDK_GENERIC_SC)) {                                /* :158-176: 16-wide unaligned blocks from st0 */
            for (t = st0; t <= en0; t += 16) {
                int k;
                int8_t blk[16];
                for (k = 0; k < 16; ++k) {                             /* loads first, like _mm_loadu_si128 */
                    uint8_t sq = sf[t + k], sr = qrr[t + k];
                    blk[k] = sq == sr ? sc_mch : sc_mis;
                    if (sq == m - 1 || sr == m - 1) blk[k] = sc_N;
                }
                memcpy(s + t, blk, 16);                                /* may run up to 15 past s[T-1], into sf[] */
            }
        } else for (t = st0; t <= en0; ++t) s[t] = mat[sf[t] * m + qrr[t]];
        if (with_cigar) { off[r] = st; off_end[r] = en; }
        {                                                              /* core loop :181-313, element-wise */
            int8_t xp = x1, x2p = x21, vp = v1;                        /* previous-diagonal values at t-1 */
            uint8_t *pr = with_cigar ? p + (size_t)r * n_col_ * 16 - st : NULL;
            for (t = st; t <= en; ++t) {
                int8_t z = s[t], xt1 = xp, vt1 = vp, x2t1 = x2p, ut = u[t], a, b, a2, b2, tmp;
                uint8_t d = 0;
                xp = x[t]; vp = v[t]; x2p = x2[t];
                a = w8(xt1 + vt1); b = w8(y[t] + ut); a2 = w8(x2t1 + vt1); b2 = w8(y2[t] + ut);
                if (!with_cigar || !(flag & NDK_RIGHT)) {              /* left-aligned gaps (:220-228) */
                    d = a > z ? 1 : 0;  z = z > a ? z : a;
                    d = b > z ? 2 : d;  z = z > b ? z : b;
                    d = a2 > z ? 3 : d; z = z > a2 ? z : a2;
                    d = b2 > z ? 4 : d; z = z > b2 ? z : b2;
                } else {                                               /* right-aligned (:267-275) */
                    d = z > a ? 0 : 1;  z = z > a ? z : a;
                    d = z > b ? d : 2;  z = z > b ? z : b;
                    d = z > a2 ? d : 3; z = z > a2 ? z : a2;
                    d = z > b2 ? d : 4; z = z > b2 ? z : b2;
                }
                z = z < sc_mch ? z : sc_mch;
                u[t] = w8(z - vt1); v[t] = w8(z - ut);                 /* __dp_code_block2 */
                tmp = w8(z - q);  a = w8(a - tmp);  b = w8(b - tmp);
                tmp = w8(z - q2); a2 = w8(a2 - tmp); b2 = w8(b2 - tmp);
                if (!with_cigar || !(flag & NDK_RIGHT)) {
                    x[t]  = w8((a  > 0 ? a  : 0) - qe);  if (a  > 0) d |= 0x08;
                    y[t]  = w8((b  > 0 ? b  : 0) - qe);  if (b  > 0) d |= 0x10;
                    x2[t] = w8((a2 > 0 ? a2 : 0) - qe2); if (a2 > 0) d |= 0x20;
                    y2[t] = w8((b2 > 0 ? b2 : 0) - qe2); if (b2 > 0) d |= 0x40;
                } else {
                    x[t]  = w8((0 > a  ? 0 : a)  - qe);  if (!(0 > a))  d |= 0x08;
                    y[t]  = w8((0 > b  ? 0 : b)  - qe);  if (!(0 > b))  d |= 0x10;
                    x2[t] = w8((0 > a2 ? 0 : a2) - qe2); if (!(0 > a2)) d |= 0x20;
                    y2[t] = w8((0 > b2 ? 0 : b2) - qe2); if (!(0 > b2)) d |= 0x40;
                }
                if (with_cigar) pr[t] = d;
            }
        }
        if (!approx_max) {                                             /* :315-358 */
            int32_t max_H, max_t;
            if (r > 0) {
                int32_t HH[4], tt[4], en1 = st0 + (en0 - st0) / 4 * 4, i;
                max_H = H[en0] = en0 > 0 ? H[en0 - 1] + u[en0] : H[en0] + v[en0];
                max_t = en0;
                for (i = 0; i < 4; ++i) { HH[i] = max_H; tt[i] = max_t; }
                for (t = st0; t < en1; t += 4)
                    for (i = 0; i < 4; ++i) {
                        H[t + i] += v[t + i];
                        if (H[t + i] > HH[i]) { HH[i] = H[t + i]; tt[i] = t; }
                    }
                for (i = 0; i < 4; ++i) if (max_H < HH[i]) { max_H = HH[i]; max_t = tt[i] + i; }
                for (; t < en0; ++t) { H[t] += (int32_t)v[t]; if (H[t] > max_H) { max_H = H[t]; max_t = t; } }
            } else { H[0] = v[0] - qe0; max_H = H[0]; max_t = 0; }
            if (en0 == tlen - 1 && H[en0] > ez->mte) { ez->mte = H[en0]; ez->mte_q = r - en; }
            if (r - st0 == qlen - 1 && H[st0] > ez->mqe) { ez->mqe = H[st0]; ez->mqe_t = st0; }
            if (apply_zdrop(ez, max_H, r, max_t, zdrop, (int8_t)e2)) break;
            if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H[tlen - 1];
        } else {                                                       /* :359-375 */
            if (r > 0) {
                if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
                    int32_t d0 = v[last_H0_t], d1 = u[last_H0_t + 1];
                    if (d0 > d1) H0 += d0; else { H0 += d1; ++last_H0_t; }
                } else if (last_H0_t >= st0 && last_H0_t <= en0) H0 += v[last_H0_t];
                else { ++last_H0_t; H0 += u[last_H0_t]; }
            } else { H0 = v[0] - qe0; last_H0_t = 0; }
            if ((flag & NDK_APPROX_DROP) && apply_zdrop(ez, H0, r, last_H0_t, zdrop, (int8_t)e2)) break;
            if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H0;
        }
        last_st = st; last_en = en;
    }
    if (with_cigar) {                                                  /* :381-392 */
        int rev = !!(flag & NDK_REV_CIGAR);
        if (!ez->zdropped && !(flag & NDK_EXTZ_ONLY))
            ez->n_cigar = backtrack(rev, p, off, off_end, n_col_ * 16, tlen - 1, qlen - 1, cigar);
        else if (!ez->zdropped && (flag & NDK_EXTZ_ONLY) && ez->mqe + end_bonus > ez->max) {
            ez->reach_end = 1;
            ez->n_cigar = backtrack(rev, p, off, off_end, n_col_ * 16, ez->mqe_t, qlen - 1, cigar);
        } else if (ez->max_t >= 0 && ez->max_q >= 0)
            ez->n_cigar = backtrack(rev, p, off, off_end, n_col_ * 16, ez->max_t, ez->max_q, cigar);
    }
    free(u); free(H); free(p); free(off);
    return 0;
}
