// nd_kswrx.cuh -- register-resident generation of the batched dual-affine extension DP (ksw_extd2).
// ref: minimap2/ksw2_extd2_sse.c:1-394, ksw2.h:119-176.  Same observable behaviour as nd_ksw.cuh (16-cell band rounding with stale
// state, score blocks spilling into the target copy, SIMD tie order of the exact maximum), different execution model:
//
//   * a target column is OWNED by a lane: column t lives at ring position t mod CAP (CAP = 128*K*NW columns); 128 consecutive
//     positions form a slot, slot s belongs to warp s mod NW (register set s / NW), lane = (pos / 4) mod 32, four consecutive
//     columns per lane and slot.  (Interleaving the slots over the warps spreads a narrow band evenly.)  The state of a column
//     (u v x y x2 y2, its score s, its target base, the 32-bit H of the exact-maximum mode) never leaves that lane's registers: no
//     per-cell memory traffic at all except the one backtrack byte.  A band wider than the target fits as it is; a band narrower
//     than the target slides over the ring (a lane recycles its four columns once they are 16 cells left of the band).
//   * the four columns of a lane are packed as s16x2 pairs (c0,c2) and (c1,c3), so the "t-1" operands of (c1,c3) ARE the old
//     (c0,c2) registers, and those of (c0,c2) need one shuffle + one byte permute per array (x, v, x2) for four cells;
//   * values are kept scaled by 16 and biased (16*v + 0x2000 per half), which buys three things: plain 32-bit IADD3 works on packed
//     halves (no borrow can cross: every half stays positive), the arg-max of the five candidates rides in the low bits as a tag
//     (two VIMNMX3 instead of four compare+select pairs; ties resolve exactly as the reference's strict/non-strict comparisons
//     because the tags are ordered accordingly), and the "gap continues" flags come out of the same VIMNMX that clamps at zero
//     (a one-bit tag survives iff the candidate won);
//   * NW warps share one alignment; they meet once per anti-diagonal (one barrier), exchanging through shared memory the x/v/x2
//     of every slot's last column (the left neighbour of the next slot), their part of the exact maximum and the few H / u / v
//     values the scalar bookkeeping needs.  Small K keeps the kernel under 128 registers (4+ warps per scheduler).
// Not handled here (routed to the memory-resident kernels of nd_ksw.cuh): KSW_EZ_GENERIC_SC, bands wider than the ring.
#pragma once
#include "nd_ksw.cuh"

#define NDX_B    0x2000                       // bias of a single value: stored = 16 * v + NDX_B (+ tag)
#define NDX_CLR  0xfff0fff0u
#define NDX_HNEG (-0x40000000)

struct ndx_smem {
    unsigned bx[32], bv[32], bx2[32];         // per slot: (c1,c3) registers of x, v, x2 in its lane 31 (the left neighbour of the next slot's first column)
    int wmax[8]; unsigned wrank[8];           // exact mode: per-warp maximum of 16*H over [st0,en0] and the smallest SIMD rank attaining it
    int h_en0, h_en0m1, h_st0, h_last;        // exact mode: 16*H of columns en0, en0-1, st0, tlen-1 after this diagonal
    int d0, d1, pad0, pad1;                   // approximate mode: new v of column last_H0_t, new u of column last_H0_t + 1
};

__device__ __forceinline__ unsigned ndx_rep(int v) { return ((unsigned)v & 0xffffu) | ((unsigned)v << 16); }
__device__ __forceinline__ int ndx_lo(unsigned p) { return (int)(p & 0xffffu); }
__device__ __forceinline__ int ndx_hi(unsigned p) { return (int)(p >> 16); }
// column j (0..3) of a lane: pairs are (c0,c2) and (c1,c3)
__device__ __forceinline__ int ndx_col(unsigned p0, unsigned p1, int j) { const unsigned p = (j & 1) ? p1 : p0; return (j & 2) ? (int)(p >> 16) : (int)(p & 0xffffu); }
__device__ __forceinline__ void ndx_setcol(unsigned &p0, unsigned &p1, int j, int v)
{
    // value selects only: taking the address of one of two registers would push both to local memory
    const unsigned lo0 = (p0 & 0xffff0000u) | ((unsigned)v & 0xffffu), hi0 = (p0 & 0xffffu) | ((unsigned)v << 16);
    const unsigned lo1 = (p1 & 0xffff0000u) | ((unsigned)v & 0xffffu), hi1 = (p1 & 0xffffu) | ((unsigned)v << 16);
    p0 = j == 0 ? lo0 : j == 2 ? hi0 : p0;
    p1 = j == 1 ? lo1 : j == 3 ? hi1 : p1;
}

struct ndx_k {                                // per-call packed constants
    unsigned MCHMAX, KXA, KYB, KX2A, KY2B, CX, CY, CX2, CY2, TA, TB, TA2, TB2;
};

// one pair of cells.  nX/nV/nX2: x, v, x2 of the left neighbours on the previous diagonal; U/Y/Y2/S: own columns.  ref: __dp_code_block1/2
// and the left-/right-aligned variants of the core loop (ksw2_extd2_sse.c:31-58,214-314).
__device__ __forceinline__ void ndx_pair(const ndx_k &C, unsigned nX, unsigned nV, unsigned nX2, unsigned S,
                                         unsigned &U, unsigned &V, unsigned &X, unsigned &Y, unsigned &X2, unsigned &Y2, unsigned &D)
{
    // z = max(s, a, b, a2, b2) with a = x[t-1] + v[t-1], b = y[t] + u[t], ...: four fused add-max (VIADDMNMX); the tag in the low bits decides ties
    unsigned M = __viaddmax_s16x2(nX, nV, S);
    M = __viaddmax_s16x2(Y, U, M);
    M = __viaddmax_s16x2(nX2, nV, M);
    M = __viaddmax_s16x2(Y2, U, M);
    const unsigned Z = __vimin3_s16x2(M, C.MCHMAX, C.MCHMAX) & NDX_CLR;
    const unsigned nU = Z - nV, nVv = Z - U;
    // a - z + q - qe = x[t-1] - (z - v[t-1]) + ... = x[t-1] - u_new + ...; likewise b - z = y - v_new
    const unsigned x1 = __vimax3_s16x2(nX - nU + C.KXA, C.CX, C.CX), y1 = __vimax3_s16x2(Y - nVv + C.KYB, C.CY, C.CY);
    const unsigned x21 = __vimax3_s16x2(nX2 - nU + C.KX2A, C.CX2, C.CX2), y21 = __vimax3_s16x2(Y2 - nVv + C.KY2B, C.CY2, C.CY2);
    const unsigned F1 = (x1 | y1) & 0x00050005u, F2 = (x21 | y21) & 0x00050005u;
    D = (F2 * 2u + F1) * 8u + (M & 0x00070007u);
    U = nU; V = nVv;
    X = (x1 & NDX_CLR) | C.TA; Y = (y1 & NDX_CLR) | C.TB; X2 = (x21 & NDX_CLR) | C.TA2; Y2 = (y21 & NDX_CLR) | C.TB2;
}

// band of anti-diagonal r (ref: ksw2_extd2_sse.c:128-139).  Returns false when the band is empty.
__device__ __forceinline__ bool ndx_band(int r, int qlen, int tlen, int w, int &st0, int &en0)
{
    int st = 0, en = tlen - 1;
    if (st < r - qlen + 1) st = r - qlen + 1;
    if (en > r) en = r;
    if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
    if (en > (r + w) >> 1) en = (r + w) >> 1;
    st0 = st; en0 = en;
    return st <= en;
}

// ref: ksw2.h:119-152 (ksw_backtrack, is_rot = 1), reading this kernel's byte format: bits 0-2 the arg-max tag, bits 3-6 the
// continuation flags of x, x2, y, y2; rows are found through the 64-row chunk table, the band of a row is recomputed.
// Called by a whole warp: lane 0 walks the path (the walk is a chain of dependent one-byte loads from a matrix that has long left
// L1 and mostly L2), the other lanes run ahead of it: a step consumes one or two rows and moves at most one column to the left, so
// the byte of row r - d lies in columns [i - d, i] of that row, and lane d pulls both ends of that range into L1 before lane 0 takes
// its next 16 steps.  Returns the CIGAR length in lane 0.
__device__ int ndx_backtrack(int lane, bool is_rev, bool right, const uint8_t *p, const unsigned long long *rowtab, size_t row0, int n_col, int qlen, int tlen, int w,
                             int i0, int j0, uint32_t *cigar)
{
    int n = 0, i = i0, j = j0, state = 0;
    int run_op = -1; uint32_t run_len = 0;                    // the CIGAR run being extended lives in registers (ksw_push_cigar merges equal neighbours, ksw2.h:100-111)
    for (;;) {
        const int bi = __shfl_sync(NDK_FULL, i, 0), bj = __shfl_sync(NDK_FULL, j, 0);
        if (bi < 0 || bj < 0) break;
        {
            const int rr = bi + bj - lane;
            if (rr >= 0) {
                int st0, en0; ndx_band(rr, qlen, tlen, w, st0, en0);
                const int st = st0 / 16 * 16, en = (en0 + 16) / 16 * 16 - 1;
                const int c_hi = bi < en ? bi : en, c_lo = bi - lane > st ? bi - lane : st;
                if (c_lo <= c_hi) {
                    const uint8_t *row = p + (rowtab ? rowtab[rr >> 6] + (size_t)(rr & 63) * n_col : row0 + (size_t)rr * n_col) - st;
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(row + c_hi));      // (a plain load with a dead result is dropped by ptxas)
                    if (((size_t)(row + c_lo) ^ (size_t)(row + c_hi)) >> 5) asm volatile("prefetch.global.L1 [%0];" :: "l"(row + c_lo));   // another 32-byte sector
                }
            }
        }
        if (lane == 0) {
            for (int step = 0; step < 16 && i >= 0 && j >= 0; ++step) {
                int force = -1;
                const int r = i + j;
                int st0, en0; ndx_band(r, qlen, tlen, w, st0, en0);
                const int st = st0 / 16 * 16, en = (en0 + 16) / 16 * 16 - 1;
                if (i < st) force = 2;
                if (i > en) force = 1;
                int tmp_state = 0; unsigned fl = 0;
                if (force < 0) {
                    const unsigned b = rowtab ? p[rowtab[r >> 6] + (size_t)(r & 63) * n_col + i - st] : p[row0 + (size_t)r * n_col + i - st];
                    const int tag = (int)(b & 7u);
                    tmp_state = right ? tag : 4 - tag;
                    fl = b >> 3;
                }
                // continuation bit of state s: x -> bit 0, y -> bit 2, x2 -> bit 1, y2 -> bit 3
                if (state == 0) state = tmp_state;
                else {
                    const unsigned bit = state == 1 ? (fl & 1u) : state == 2 ? (fl >> 2 & 1u) : state == 3 ? (fl >> 1 & 1u) : (fl >> 3 & 1u);
                    if (!bit) state = 0;
                }
                if (state == 0) state = tmp_state;
                if (force >= 0) state = force;
                const int op = state == 0 ? 0 : (state == 1 || state == 3) ? 2 : 1;
                if (op == run_op) ++run_len; else { if (run_op >= 0) cigar[n++] = run_len << 4 | (uint32_t)run_op; run_op = op; run_len = 1; }
                if (op == 0) { --i; --j; } else if (op == 2) --i; else --j;
            }
        }
    }
    if (lane == 0) {
        if (i >= 0) { if (run_op == 2) run_len += (uint32_t)(i + 1); else { if (run_op >= 0) cigar[n++] = run_len << 4 | (uint32_t)run_op; run_op = 2; run_len = (uint32_t)(i + 1); } }
        if (j >= 0) { if (run_op == 1) run_len += (uint32_t)(j + 1); else { if (run_op >= 0) cigar[n++] = run_len << 4 | (uint32_t)run_op; run_op = 1; run_len = (uint32_t)(j + 1); } }
        if (run_op >= 0) cigar[n++] = run_len << 4 | (uint32_t)run_op;
        if (!is_rev) for (int k = 0; k < n >> 1; ++k) { uint32_t t = cigar[k]; cigar[k] = cigar[n - 1 - k]; cigar[n - 1 - k] = t; }
    }
    return n;
}

// Sequence layout expected here (the host re-materialises calls for this kernel): query and target start 4-byte aligned, at least 32
// zero bytes precede the query and at least 48 zero bytes follow both the query and the target.
template <int K, int NW, bool EXACT>
__global__ void __launch_bounds__(NW * 32, 512 / (NW * 32))
k_ksw_rx(const ndk_call *calls, int n_calls, const uint8_t *seqs, ndk_par P, ndk_res *res, uint8_t *pmat, uint32_t *cigars,
         unsigned long long *rowtabs, size_t tab_entries, uint32_t *work, unsigned long long *arena_cur, unsigned long long arena_bytes)
{
    constexpr int SLOTS = K * NW, CAP = 128 * SLOTS;
    static_assert(SLOTS <= 32, "ring too large for the exchange record");
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, gtid = threadIdx.x;
    __shared__ ndx_smem xch[2];
    __shared__ unsigned long long s_o;
    __shared__ int s_ci;
    unsigned long long *rowtab = arena_cur ? rowtabs + (size_t)blockIdx.x * tab_entries : nullptr;
    auto sync = [&]() { if (NW == 1) __syncwarp(); else __syncthreads(); };

    for (;;) {
        if (gtid == 0) s_ci = (int)atomicAdd(work, 1u);
        sync();
        const int ci = s_ci;
        sync();
        if (ci >= n_calls) break;
        const ndk_call c = calls[ci];
        const int qlen = c.qlen, tlen = c.tlen, flag = c.flag, zdrop = c.zdrop;
        const uint8_t *query = seqs + c.q_off, *target = seqs + c.t_off;
        const bool with_cigar = !(flag & NDK_SCORE_ONLY), right = with_cigar && (flag & NDK_RIGHT);
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
        const bool ring = T > CAP;                          // the band slides over the ring; otherwise every column has its own place
        uint8_t *p = pmat;
        const size_t row0 = rowtab ? 0 : (size_t)c.p_off;
        uint32_t *cigar = cigars + c.cigar_off;
        bool arena_full = false;
        size_t chunk_off = 0;

        if (run) {
            // ---- constants of the scaled/biased/tagged arithmetic (header comment) -----------------------------------------------
            const int tZ = right ? 0 : 4, tA = right ? 1 : 3, tB = 2, tA2 = right ? 3 : 1, tB2 = right ? 4 : 0;
            const int Lx = right ? 0 : 2, Ly = right ? 0 : 8;
            ndx_k C;
            C.MCHMAX = ndx_rep(16 * sc_mch + 2 * NDX_B + 15);
            C.KXA = ndx_rep(16 * (q - qe) + NDX_B - tA + 1);     C.KYB = ndx_rep(16 * (q - qe) + NDX_B - tB + 4);
            C.KX2A = ndx_rep(16 * (q2 - qe2) + NDX_B - tA2 + 1); C.KY2B = ndx_rep(16 * (q2 - qe2) + NDX_B - tB2 + 4);
            C.CX = ndx_rep(16 * -qe + NDX_B + Lx);  C.CY = ndx_rep(16 * -qe + NDX_B + Ly);
            C.CX2 = ndx_rep(16 * -qe2 + NDX_B + Lx); C.CY2 = ndx_rep(16 * -qe2 + NDX_B + Ly);
            C.TA = ndx_rep(tA); C.TB = ndx_rep(tB); C.TA2 = ndx_rep(tA2); C.TB2 = ndx_rep(tB2);
            const int I1 = 16 * (-q - e) + NDX_B, I2 = 16 * (-q2 - e2) + NDX_B;       // initial state values (:105-110)
            const unsigned S_MCH = ndx_rep(16 * sc_mch + 2 * NDX_B + tZ), S_D = (unsigned)(16 * (sc_mch - sc_mis)), S_N = ndx_rep(16 * sc_N + 2 * NDX_B + tZ);
            const unsigned S_0 = ndx_rep(2 * NDX_B + tZ);        // score 0: what a never-written s[] cell holds (calloc)
            // a byte equal to the ambiguity code matters only when a sequence holds one, or when a score byte that spilled into the
            // target copy happens to equal it
            bool slow_n = ((unsigned)sc_mch & 255u) == (unsigned)(m - 1) || ((unsigned)sc_mis & 255u) == (unsigned)(m - 1) || ((unsigned)sc_N & 255u) == (unsigned)(m - 1);
            slow_n = slow_n || (flag & 0x10000) != 0;           // host-set: the call contains an ambiguous base
            unsigned M1 = ndx_rep(m - 1);
            unsigned S_MCHr = S_MCH, S_Dr = S_D;
            // keep the loop constants in registers: left to itself the compiler re-derives them from the scalar options on every use
#define NDX_PIN(x) asm volatile("" : "+r"(x))
            NDX_PIN(C.MCHMAX); NDX_PIN(C.KXA); NDX_PIN(C.KYB); NDX_PIN(C.KX2A); NDX_PIN(C.KY2B); NDX_PIN(C.CX); NDX_PIN(C.CY); NDX_PIN(C.CX2); NDX_PIN(C.CY2);
            NDX_PIN(C.TA); NDX_PIN(C.TB); NDX_PIN(C.TA2); NDX_PIN(C.TB2); NDX_PIN(S_MCHr); NDX_PIN(S_Dr);
            int vb_first = -q - e, vb_lo = -e, vb_hi = -e2, lt_p = long_thres, ld_p = long_diff;     // first-column boundary values (:140-152) per diagonal
#undef NDX_PIN

            // ---- state: register set k of this warp is slot k * NW + wid -------------------------------------------------------------
            unsigned U[K][2], V[K][2], X[K][2], Y[K][2], X2[K][2], Y2[K][2], S[K][2], TG[K];
            int t0[K];
            int Hh[EXACT ? K : 1][4];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                t0[k] = (k * NW + wid) * 128 + 4 * lane;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    U[k][h] = V[k][h] = ndx_rep(I1); X[k][h] = ndx_rep(I1 + tA); Y[k][h] = ndx_rep(I1 + tB);
                    X2[k][h] = ndx_rep(I2 + tA2); Y2[k][h] = ndx_rep(I2 + tB2); S[k][h] = S_0;
                }
                TG[k] = t0[k] < T ? *reinterpret_cast<const uint32_t *>(target + t0[k]) : 0u;
                if (EXACT) { Hh[k][0] = Hh[k][1] = Hh[k][2] = Hh[k][3] = NDX_HNEG; }
            }
            if (gtid < 2) {
                ndx_smem &x = xch[gtid];
                for (int i = 0; i < 32; ++i) { x.bx[i] = ndx_rep(I1 + tA); x.bv[i] = ndx_rep(I1); x.bx2[i] = ndx_rep(I2 + tA2); }
                for (int i = 0; i < 8; ++i) { x.wmax[i] = NDX_HNEG; x.wrank[i] = 0xffffffffu; }
                x.h_en0 = x.h_en0m1 = x.h_st0 = x.h_last = NDX_HNEG; x.d0 = x.d1 = 0;
            }
            sync();

            int last_st = -1, last_en = -1, H0 = 0, last_H0_t = 0;
            int hp0 = NDX_HNEG, hp1 = NDX_HNEG, en0_prev = -1;     // 16*H of columns en0_prev and en0_prev - 1 after the previous diagonal
            const int n_diag = qlen + tlen - 1;
            for (int r = 0; r < n_diag; ++r) {
                int st0, en0;
                if (!ndx_band(r, qlen, tlen, w, st0, en0)) { ez.zdropped = 1; break; }
                const int st = st0 / 16 * 16, en = (en0 + 16) / 16 * 16 - 1;
                const int vb = r == 0 ? vb_first : r < lt_p ? vb_lo : r == lt_p ? ld_p : vb_hi;
                const int sc_hi = st0 + (en0 - st0 + 16) / 16 * 16;                  // scores are rewritten on [st0, sc_hi) (:158-176)
                const int live_hi = sc_hi < T ? sc_hi : T;                           // ... of which columns >= T alias the target copy
                ndx_smem &xo = xch[(r & 1) ^ 1];                                     // what the previous diagonal published
                ndx_smem &xn = xch[r & 1];
                if (rowtab && with_cigar && (r & 63) == 0) {                         // a new 64-row chunk of the backtrack arena
                    if (gtid == 0) { const unsigned long long o = atomicAdd(arena_cur, 64ull * (unsigned long long)n_col); rowtab[r >> 6] = o; s_o = o; }
                    sync();
                    const unsigned long long o = s_o;
                    sync();
                    if (o + 64ull * (unsigned long long)n_col > arena_bytes) { arena_full = true; break; }
                    chunk_off = (size_t)o;
                }
                uint8_t *prow = p + (rowtab ? chunk_off + (size_t)(r & 63) * n_col : row0 + (size_t)r * n_col) - st;

                // left-boundary operands of cell st (:140-152), packed with the tags of x / x2
                int bX, bV, bX2; bool patch;
                if (st > 0) { patch = !(st - 1 >= last_st && st - 1 <= last_en); bX = I1 + tA; bV = I1; bX2 = I2 + tA2; }
                else { patch = true; bX = I1 + tA; bV = 16 * vb + NDX_B; bX2 = I2 + tA2; }

                if (ring && st != last_st) {                                         // columns that fell 16 cells behind the band for good: recycle (st moves every 32nd diagonal)
#pragma unroll
                    for (int k = 0; k < K; ++k)
                        if (t0[k] + 3 < st - 16 && t0[k] + CAP < T) {
                            t0[k] += CAP;
                            U[k][0] = U[k][1] = V[k][0] = V[k][1] = ndx_rep(I1); X[k][0] = X[k][1] = ndx_rep(I1 + tA); Y[k][0] = Y[k][1] = ndx_rep(I1 + tB);
                            X2[k][0] = X2[k][1] = ndx_rep(I2 + tA2); Y2[k][0] = Y2[k][1] = ndx_rep(I2 + tB2); S[k][0] = S[k][1] = S_0;
                            TG[k] = *reinterpret_cast<const uint32_t *>(target + t0[k]);
                            if (EXACT) { Hh[k][0] = Hh[k][1] = Hh[k][2] = Hh[k][3] = NDX_HNEG; }
                        }
                }
                // first row: y, y2, u of column r (:153-156).  The slot is warp-uniform (a uniform branch in the slot loop); the lane and the
                // half-word enter as select masks
                int fr_sl = -1; unsigned fr_m0 = 0, fr_m1 = 0;
                if (en >= r) {
                    const int pos = r % CAP;
                    fr_sl = pos >> 7;
                    const unsigned hm = ((pos >> 2) & 31) == lane ? ((pos & 2) ? 0xffff0000u : 0x0000ffffu) : 0u;
                    fr_m0 = (pos & 1) ? 0u : hm; fr_m1 = (pos & 1) ? hm : 0u;
                }
                int lmax = NDX_HNEG, lk = 0; bool ltie = false;                      // exact mode: this lane's best 16*H over [st0, en0], its register set, seen twice?
                const int en1 = st0 + (en0 - st0) / 4 * 4;
                const int qidx_base = r - 3;
                const unsigned band_w = (unsigned)(en - st), live_w = (unsigned)(live_hi + 3 - st0);
                const unsigned pm_on = patch ? 0x0000ffffu : 0u;
                const int lo_c = st < st0 - 3 ? st : st0 - 3, hi_c = en > live_hi - 1 ? en : live_hi - 1;   // columns this diagonal touches (hull)
                // NW == 1: all slots of the call live in this warp, so lane 0 of a slot gets its left neighbour -- lane 31 of the slot before --
                // by the same shuffle as the other lanes: lane 31 feeds the previous slot's register into a rotate-by-one shuffle, and the
                // slots run in DESCENDING order (slot k-1 still holds the previous diagonal's values when slot k reads them).  Slot 0 wraps
                // around to the last slot (ring), whose old values are saved here.
                unsigned wX = 0, wV = 0, wX2 = 0;
                if (NW == 1) { wX = X[K - 1][1]; wV = V[K - 1][1]; wX2 = X2[K - 1][1]; }
                // approximate mode: v of column last_H0_t and u of column last_H0_t + 1 for the bookkeeping (:361-371) -- the array values the
                // reference would read, whether or not the column was computed on this diagonal.  Slot and half-word are warp-uniform: the
                // slot loop keeps a copy of the register pair, the owning lane broadcasts it afterwards
                const int p0 = last_H0_t % CAP, p1 = (last_H0_t + 1) % CAP;
                const unsigned m0 = 0u - (unsigned)(p0 & 1), m1 = 0u - (unsigned)(p1 & 1);   // bit masks, not `c ? a[1] : a[0]`: that would turn the register arrays into indexed local memory

#pragma unroll
                for (int kk = 0; kk < K; ++kk) {
                    const int k = NW == 1 ? K - 1 - kk : kk;
                    const int tt = t0[k];
                    const bool act = (unsigned)(tt - st) <= band_w;                  // st, en are multiples of 16 (-1): all four columns or none
                    const bool live = (unsigned)(tt + 3 - st0) < live_w;             // some column gets a fresh score: tt+3 >= st0 && tt < live_hi
                    const int sl = k * NW + wid;                                     // this register set's slot
                    // is the slot touched at all?  Without a sliding ring its columns are [128 sl, 128 sl + 127]: a scalar test; on a ring the
                    // lanes of a slot can be one generation apart, so they vote
                    bool touched;
                    if (ring) touched = __any_sync(NDK_FULL, act || live); else touched = sl * 128 <= hi_c && sl * 128 + 127 >= lo_c;
                    if (touched) {
                    if (fr_sl == sl) {                                               // warp-uniform; column r is inside [st, en], so its slot is a touched one
                        const unsigned vy = ndx_rep(I1 + tB), vy2 = ndx_rep(I2 + tB2), vu = ndx_rep(16 * vb + NDX_B);
                        Y[k][0] = (Y[k][0] & ~fr_m0) | (vy & fr_m0);    Y[k][1] = (Y[k][1] & ~fr_m1) | (vy & fr_m1);
                        Y2[k][0] = (Y2[k][0] & ~fr_m0) | (vy2 & fr_m0); Y2[k][1] = (Y2[k][1] & ~fr_m1) | (vy2 & fr_m1);
                        U[k][0] = (U[k][0] & ~fr_m0) | (vu & fr_m0);    U[k][1] = (U[k][1] & ~fr_m1) | (vu & fr_m1);
                    }

                    // ---- scores of this diagonal (:157-176), stale elsewhere ----------------------------------------------------------------
                    if (live) {
                        const int qi = qidx_base - tt;                               // query[qi .. qi+3] face columns c3 .. c0
                        const int qa = qi & ~3;
                        const uint32_t w0 = *reinterpret_cast<const uint32_t *>(query + qa), w1 = *reinterpret_cast<const uint32_t *>(query + qa + 4);
                        const uint32_t w4 = __funnelshift_r(w0, w1, 8 * (qi & 3));    // bytes b0..b3 = query[qi..qi+3]
                        const unsigned QP0 = __byte_perm(w4, 0u, 0x4143), QP1 = __byte_perm(w4, 0u, 0x4042);   // (c0,c2) = (b3,b1), (c1,c3) = (b2,b0)
                        const unsigned TP0 = __byte_perm(TG[k], 0u, 0x4240), TP1 = __byte_perm(TG[k], 0u, 0x4341);
                        unsigned s0 = S_MCHr - __vminu2(TP0 ^ QP0, 0x00010001u) * S_Dr;
                        unsigned s1 = S_MCHr - __vminu2(TP1 ^ QP1, 0x00010001u) * S_Dr;
                        if (slow_n) {
                            const unsigned e0 = __vminu2(TP0 ^ M1, 0x00010001u) & __vminu2(QP0 ^ M1, 0x00010001u);   // 0 where an operand is the ambiguity code
                            const unsigned e1 = __vminu2(TP1 ^ M1, 0x00010001u) & __vminu2(QP1 ^ M1, 0x00010001u);
                            const unsigned k0 = (e0 ^ 0x00010001u) * 0xffffu, k1 = (e1 ^ 0x00010001u) * 0xffffu;      // 0xffff where ambiguous
                            s0 = (s0 & ~k0) | (S_N & k0); s1 = (s1 & ~k1) | (S_N & k1);
                        }
                        if (tt >= st0 && tt + 3 < live_hi) { S[k][0] = s0; S[k][1] = s1; }
                        else {
                            const unsigned span = (unsigned)(live_hi - st0);
                            const unsigned m0 = ((unsigned)(tt - st0) < span ? 0xffffu : 0u) | ((unsigned)(tt + 2 - st0) < span ? 0xffff0000u : 0u);
                            const unsigned m1 = ((unsigned)(tt + 1 - st0) < span ? 0xffffu : 0u) | ((unsigned)(tt + 3 - st0) < span ? 0xffff0000u : 0u);
                            S[k][0] = (s0 & m0) | (S[k][0] & ~m0); S[k][1] = (s1 & m1) | (S[k][1] & ~m1);
                        }
                    }

                    // ---- left neighbours of (c0,c2): previous lane's (c1,c3); lane 0 takes what the previous slot's lane 31 published -------------------
                    unsigned gX, gV, gX2;
                    if (NW == 1) {
                        const int kp = k > 0 ? k - 1 : 0;
                        const bool l31 = lane == 31;
                        const unsigned eX = l31 ? (k == 0 ? wX : X[kp][1]) : X[k][1], eV = l31 ? (k == 0 ? wV : V[kp][1]) : V[k][1], eX2 = l31 ? (k == 0 ? wX2 : X2[kp][1]) : X2[k][1];
                        const int src = (lane + 31) & 31;
                        gX = __shfl_sync(NDK_FULL, eX, src); gV = __shfl_sync(NDK_FULL, eV, src); gX2 = __shfl_sync(NDK_FULL, eX2, src);
                    } else {
                        const int ps = sl == 0 ? SLOTS - 1 : sl - 1;
                        gX = __shfl_up_sync(NDK_FULL, X[k][1], 1); gV = __shfl_up_sync(NDK_FULL, V[k][1], 1); gX2 = __shfl_up_sync(NDK_FULL, X2[k][1], 1);
                        if (lane == 0) { gX = xo.bx[ps]; gV = xo.bv[ps]; gX2 = xo.bx2[ps]; }
                    }
                    if (act) {
                        unsigned nX = __byte_perm(gX, X[k][1], 0x5432), nV = __byte_perm(gV, V[k][1], 0x5432), nX2 = __byte_perm(gX2, X2[k][1], 0x5432);
                        {   // cell st takes the boundary values when column st - 1 was outside the previous band
                            const unsigned pm = tt == st ? pm_on : 0u;
                            nX = (nX & ~pm) | ((unsigned)bX & pm); nV = (nV & ~pm) | ((unsigned)bV & pm); nX2 = (nX2 & ~pm) | ((unsigned)bX2 & pm);
                        }
                        // ---- the recurrence: (c1,c3) reads the old (c0,c2) as its neighbours ---------------------------------------------------
                        const unsigned oX = X[k][0], oV = V[k][0], oX2 = X2[k][0];
                        unsigned D0, D1;
                        ndx_pair(C, nX, nV, nX2, S[k][0], U[k][0], V[k][0], X[k][0], Y[k][0], X2[k][0], Y2[k][0], D0);
                        ndx_pair(C, oX, oV, oX2, S[k][1], U[k][1], V[k][1], X[k][1], Y[k][1], X2[k][1], Y2[k][1], D1);
                        if (with_cigar) *reinterpret_cast<uint32_t *>(prow + tt) = __byte_perm(D0, D1, 0x6240);   // bytes c0 c1 c2 c3
                        if (NW > 1 && lane == 31) { xn.bx[sl] = X[k][1]; xn.bv[sl] = V[k][1]; xn.bx2[sl] = X2[k][1]; }

                        // ---- exact maximum: H[t] += v[t] over [st0, en0), H[en0] from its left neighbour (:316-346).  A lane keeps its best
                        // 16*H and where it saw it; the SIMD-search rank of the winner is worked out after the loop
                        if (EXACT) {
                            if (r > 0) {
                                if (tt > st0 && tt + 3 < en0 - 1) {                  // four interior columns (st0, en0 - 1 and en0 take the other branch)
                                    Hh[k][0] += ndx_lo(V[k][0]) - NDX_B; Hh[k][1] += ndx_lo(V[k][1]) - NDX_B; Hh[k][2] += ndx_hi(V[k][0]) - NDX_B; Hh[k][3] += ndx_hi(V[k][1]) - NDX_B;
                                    const int m4 = max(max(Hh[k][0], Hh[k][1]), max(Hh[k][2], Hh[k][3]));
                                    if (m4 >= lmax) { ltie = ltie || m4 == lmax; if (m4 > lmax) ltie = false; lmax = m4; lk = k; }
                                } else if (tt + 3 >= st0 && tt <= en0) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const int t = tt + j;
                                        if (t < st0 || t > en0) continue;
                                        int h;
                                        if (t < en0) h = Hh[k][j] + ndx_col(V[k][0], V[k][1], j) - NDX_B;
                                        else {                                        // t == en0: H[en0] = H[en0-1] + u[en0] (old H of the left neighbour), :323
                                            const int hl = en0 == en0_prev ? hp1 : hp0;  // column en0 - 1 after the previous diagonal
                                            h = en0 > 0 ? hl + ndx_col(U[k][0], U[k][1], j) - NDX_B : Hh[k][j] + ndx_col(V[k][0], V[k][1], j) - NDX_B;
                                            xn.h_en0 = h;
                                        }
                                        Hh[k][j] = h;
                                        if (t == en0 - 1) xn.h_en0m1 = h;
                                        if (t == st0) xn.h_st0 = h;
                                        if (h >= lmax) { ltie = ltie || h == lmax; if (h > lmax) ltie = false; lmax = h; lk = k; }
                                    }
                                }
                            } else if (tt == 0) {                                    // r == 0: H[0] = v[0] - qe (:345)
                                Hh[k][0] = ndx_lo(V[k][0]) - NDX_B - 16 * qe0;
                                lmax = Hh[k][0]; lk = k; xn.h_en0 = lmax; xn.h_st0 = lmax;
                            }
                        }
                    }
                    // approximate mode: the two values the bookkeeping reads (:361-371) are columns of [st0, en0] whenever it uses them, i.e. of a
                    // touched slot
                    if (!EXACT) {                                                    // warp-uniform branches, one lane stores
                        if ((p0 >> 7) == sl) { const unsigned pr = V[k][0] ^ ((V[k][0] ^ V[k][1]) & m0); if (lane == ((p0 >> 2) & 31)) xn.d0 = ((int)((p0 & 2) ? pr >> 16 : pr & 0xffffu) - NDX_B) >> 4; }
                        if ((p1 >> 7) == sl) { const unsigned pr = U[k][0] ^ ((U[k][0] ^ U[k][1]) & m1); if (lane == ((p1 >> 2) & 31)) xn.d1 = ((int)((p1 & 2) ? pr >> 16 : pr & 0xffffu) - NDX_B) >> 4; }
                    }
                    }   // slot touched on this diagonal
                }

                // ---- hand the bookkeeping its operands (one lane each) -----------------------------------------------------------------------------
                if (EXACT) {
                    const int wm = __reduce_max_sync(NDK_FULL, lmax);
                    // smallest SIMD-search rank among this warp's columns of [st0, en0] that attain the maximum (:320-343: en0 first, then the
                    // 4-lane blocks lane-major, then the scalar tail).  Usually one lane holds the maximum once: only its register set lk is scanned.
                    unsigned lrank = 0xffffffffu;
                    const bool mine = lmax == wm && wm != NDX_HNEG;
                    const unsigned cand = __ballot_sync(NDK_FULL, mine);
                    if (mine) {
                        const bool full = ltie || (cand & (cand - 1)) != 0;         // a tie inside this lane or between lanes: scan every register set
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            if (!full && k != lk) continue;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int t = t0[k] + j;
                                if (t >= st0 && t <= en0 && Hh[k][j] == wm) {
                                    const unsigned rank = r == 0 || t == en0 ? 0u : t < en1 ? 1u + ((unsigned)((t - st0) & 3) << 20) + (unsigned)((t - st0) >> 2) : 1u + (4u << 20) + (unsigned)(t - en1);
                                    lrank = rank < lrank ? rank : lrank;
                                }
                            }
                        }
                    }
                    const unsigned wr = __reduce_min_sync(NDK_FULL, lrank);
                    if (lane == 0) { xn.wmax[wid] = wm; xn.wrank[wid] = wr; }
                }
                sync();

                // ---- scores that spilled past the target copy (:103,:175): target bytes of columns 0..15 are overwritten ---------------------------------
                if (sc_hi > T && st0 < 16 && wid == 0 && t0[0] < 16 && t0[0] < sc_hi - T) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int col = t0[0] + j;
                        if (col < sc_hi - T) {
                            const int sq = (int)query[qlen - 1 - col], sr = (int)query[r - T - col];     // qr[col] read as a target base, qrr[T + col]
                            int z = sq == sr ? sc_mch : sc_mis; if (sq == m - 1 || sr == m - 1) z = sc_N;
                            TG[0] = (TG[0] & ~(0xffu << (8 * j))) | (((unsigned)z & 0xffu) << (8 * j));
                        }
                    }
                }

                // ---- scalar bookkeeping, replicated in every thread (:316-376) ----------------------------------------------------------------------
                if (EXACT) {
                    int best; unsigned brank;
                    if (NW == 1) { best = xn.wmax[0]; brank = xn.wrank[0]; }
                    else {                                                           // lane w reads warp w's part; two warp reductions
                        const int hm = lane < NW ? xn.wmax[lane & 7] : NDX_HNEG; const unsigned rk = lane < NW ? xn.wrank[lane & 7] : 0xffffffffu;
                        best = __reduce_max_sync(NDK_FULL, hm);
                        brank = __reduce_min_sync(NDK_FULL, hm == best ? rk : 0xffffffffu);
                    }
                    const int max_H = best >> 4;
                    int max_t;
                    if (r == 0) max_t = 0;
                    else if (brank == 0) max_t = en0;
                    else if (brank - 1 < (4u << 20)) max_t = st0 + (int)((brank - 1) & 0xfffff) * 4 + (int)((brank - 1) >> 20);
                    else max_t = en1 + (int)(brank - 1 - (4u << 20));
                    {   // H of columns en0 and en0 - 1 as the next diagonal will find them; a column left of the band kept its old value
                        const int hl = en0 == en0_prev ? hp1 : hp0;
                        hp0 = xn.h_en0; hp1 = en0 - 1 >= st0 ? xn.h_en0m1 : hl; en0_prev = en0;
                    }
                    if (en0 == tlen - 1) { const int h = hp0 >> 4; if (h > ez.mte) { ez.mte = h; ez.mte_q = r - en; } }
                    if (r - st0 == qlen - 1) { const int h = xn.h_st0 >> 4; if (h > ez.mqe) { ez.mqe = h; ez.mqe_t = st0; } }
                    if (ndk_zdrop(ez, max_H, r, max_t, zdrop, (int)(int8_t)e2)) break;
                    if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = hp0 >> 4;     // H[tlen-1] = H[en0]
                } else {
                    const int dd0 = xn.d0, dd1 = xn.d1;
                    if (r > 0) {
                        if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
                            if (dd0 > dd1) H0 += dd0; else { H0 += dd1; ++last_H0_t; }
                        } else if (last_H0_t >= st0 && last_H0_t <= en0) H0 += dd0;
                        else { ++last_H0_t; H0 += dd1; }
                    } else { H0 = dd0 - qe0; last_H0_t = 0; }
                    if ((flag & NDK_APPROX_DROP) && ndk_zdrop(ez, H0, r, last_H0_t, zdrop, (int)(int8_t)e2)) break;
                    if (r == qlen + tlen - 2 && en0 == tlen - 1) ez.score = H0;
                }
                last_st = st; last_en = en;
            }
            sync();
            if (arena_full) ez.pad = 1;
            else if (with_cigar && wid == 0) {                                          // warp 0: lane 0 walks, the others prefetch (ez is replicated)
                const bool rev = (flag & NDK_REV_CIGAR) != 0;
                if (!ez.zdropped && !(flag & NDK_EXTZ_ONLY))
                    ez.n_cigar = ndx_backtrack(lane, rev, right, p, rowtab, row0, n_col, qlen, tlen, w, tlen - 1, qlen - 1, cigar);
                else if (!ez.zdropped && (flag & NDK_EXTZ_ONLY) && ez.mqe + c.end_bonus > ez.max) {
                    ez.reach_end = 1;
                    ez.n_cigar = ndx_backtrack(lane, rev, right, p, rowtab, row0, n_col, qlen, tlen, w, ez.mqe_t, qlen - 1, cigar);
                } else if (ez.max_t >= 0 && ez.max_q >= 0)
                    ez.n_cigar = ndx_backtrack(lane, rev, right, p, rowtab, row0, n_col, qlen, tlen, w, ez.max_t, ez.max_q, cigar);
            }
        }
        if (gtid == 0) res[ci] = ez;
        sync();
    }
}

