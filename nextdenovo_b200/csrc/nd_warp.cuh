// nd_warp.cuh -- warp-cooperative generation of the two heavy kernels: the chaining DP forward pass
// (one warp per independent anchor segment) and the exact flag-sort emulation (breadth-first over the
// radix levels, one warp per bucket job).  sm_100a; uses redux.sync, ballot, fns.
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include "nd_core.cuh"
#include "nd_kernels.cuh"

#define NDW_FULL 0xffffffffu

// ------------------------------------------------------------------------------------------------
// chaining DP, forward pass.  ref: minimap2/chain.c:45-86.
//
// For anchor i the reference scans j = i-1 .. st sequentially with three pieces of sequential state:
// the strict running maximum (max_f, max_j), the saturating skip counter n_skip with its early
// exit, and the per-iteration marks t[p[j]] = i.  One warp evaluates 32 predecessors per step:
//   * marks only flow towards smaller j (p[j] < j) and an early exit only discards smaller j, so
//     every candidate lane may publish its mark speculatively before the exit position is known;
//   * the lanes that raise the running maximum are the left-to-right records of the score vector:
//     found right-to-left with redux.max + ballot (usually 1-2 rounds);
//   * with the record mask N and the "marked, not a record" mask M known to all lanes, n_skip is
//     replayed on bit masks (popc between records, fns for the exit lane) without communication.
// The most recent NDW_RING anchors (x, q, f, p, v, t) live in shared memory; older ones are read
// from global memory.
// ------------------------------------------------------------------------------------------------
#define NDW_RING 64
#define NDW_CHAIN_WARPS 4
#define NDW_GAP_TAB 2048

struct ndw_ring {
    uint32_t x[NDW_RING];
    int32_t q[NDW_RING], f[NDW_RING], p[NDW_RING], v[NDW_RING], t[NDW_RING];
};

__global__ void __launch_bounds__(NDW_CHAIN_WARPS * 32)
k_chain_fill_warp(nd_seg_view S, nd_reads_dev Q, const float *avg_qspan, nd_chain_par par, int max_span, int uniform_span,
                  nd_chain_bufs B, uint32_t *work)
{
    __shared__ ndw_ring rings[NDW_CHAIN_WARPS];
    __shared__ int16_t s_gap[NDW_GAP_TAB];
    const int lane = threadIdx.x & 31;
    ndw_ring &R = rings[threadIdx.x >> 5];
    const nd_u64 mdx = (nd_u64)par.max_dist_x;
    // Without HPC every anchor spans exactly k bases, so the average span (chain.c:41-42) is the same float for
    // every read and the gap cost (chain.c:72) depends on the diagonal distance only: tabulate it once per block.
    const bool use_tab = uniform_span > 0 && par.bw < NDW_GAP_TAB;
    if (use_tab) {
        const float avg_u = __fdiv_rn(__ull2float_rn((nd_u64)uniform_span), __ll2float_rn(1LL));
        for (int d = threadIdx.x; d <= par.bw; d += blockDim.x) s_gap[d] = (int16_t)nd_gap_cost(d, avg_u);
    }
    __syncthreads();

    // Most segments are a stray anchor or two and cannot reach min_cnt / min_chain_score: a warp claims 32 segments with one atomic,
    // every lane sizes one of them, and only the viable ones are chained (one queue round trip per segment used to be the largest
    // stall of this kernel).
    const uint32_t n_batches = (S.n_seg + 31u) >> 5;
    for (;;) {
        uint32_t s0 = 0;
        if (lane == 0) s0 = atomicAdd(work, 32u);
        s0 = __shfl_sync(NDW_FULL, s0, 0);
        if (s0 >= S.n_seg) break;
        nd_u64 b_l = 0; int n_l = 0; bool ok_l = false;
        // lane l of batch b sizes segment l * n_batches + b: the viable segments of a read lie next to each other (the group filter
        // of the sort puts the stray anchors behind them), and a warp that drew 32 of them in a row would chain them one after another
        // while its neighbours idle (measured: 0.33 -> 0.57 ms with consecutive claims)
        const uint32_t s_l = (uint32_t)lane * n_batches + (s0 >> 5);
        if (s_l < S.n_seg) ok_l = nd_seg_bounds(S, s_l, par, max_span, &b_l, &n_l);
        unsigned viable = __ballot_sync(NDW_FULL, ok_l);
      while (viable) {
        const int src = __ffs(viable) - 1; viable &= viable - 1;
        const nd_u64 b = (nd_u64)__shfl_sync(NDW_FULL, (uint32_t)(b_l >> 32), src) << 32 | __shfl_sync(NDW_FULL, (uint32_t)b_l, src);
        const int n = __shfl_sync(NDW_FULL, n_l, src);
        const int r = nd_find_owner(S.a_off, Q.n_reads, b);
        const float avg = avg_qspan[r];
        const bool tab_ok = use_tab && avg == (float)uniform_span;   // (float)sum/n can differ from k once sum >= 2^24
        const nd_u128 *a = S.a + b;
        int32_t *f = B.f + b, *p = B.p + b, *v = B.v + b, *t = B.t + b;

        for (int i = lane; i < n; i += 32) t[i] = 0;
        for (int i = lane; i < NDW_RING; i += 32) R.t[i] = 0;
        __syncwarp();

        int st = 0;
        nd_u64 x_st = a[0].x;
        nd_u128 ai = a[0];
        for (int i = 0; i < n; ++i) {
            const nd_u128 a_next = i + 1 < n ? a[i + 1] : ai;      // prefetch
            const uint32_t xi = (uint32_t)ai.x;
            const int32_t qi = (int32_t)ai.y, q_span = (int32_t)(ai.y >> 32 & 0xff);
            while (st < i && ai.x > x_st + mdx) {                    // chain.c:50
                int idx = st + lane;
                bool adv = idx < i && ai.x > a[idx].x + mdx;
                unsigned m = __ballot_sync(NDW_FULL, adv);
                int c = m == NDW_FULL ? 32 : __ffs(~m) - 1;
                st += c; x_st = a[st].x;
                if (c < 32) break;
            }
            if (i - st > par.max_iter) { st = i - par.max_iter; x_st = a[st].x; }   // chain.c:51

            int32_t max_f = q_span, max_j = -1, n_skip = 0;
            bool stop = false;
            for (int jb = i - 1; jb >= st && !stop; jb -= 32) {
                const int j = jb - lane;
                const bool in = j >= st;
                const bool near = i - j <= NDW_RING;
                uint32_t xj = 0; int32_t qj = 0, fj = 0, pj = -1;
                if (in) {
                    if (near) { int k = j & (NDW_RING - 1); xj = R.x[k]; qj = R.q[k]; fj = R.f[k]; pj = R.p[k]; }
                    else { nd_u128 aj = a[j]; xj = (uint32_t)aj.x; qj = (int32_t)aj.y; fj = f[j]; pj = p[j]; }
                }
                bool cand = false; int32_t sc = INT_MIN;
                if (in) {
                    const uint32_t dr = xi - xj;
                    const int32_t dq = qi - qj;
                    if (dr != 0 && dq > 0 && dq <= par.max_dist_y && dq <= par.max_dist_x) {
                        nd_i64 d = (nd_i64)dr - (nd_i64)dq;
                        if (d < 0) d = -d;
                        if (d <= (nd_i64)par.bw) {
                            int32_t m = (nd_i64)dq < (nd_i64)dr ? dq : (int32_t)dr;
                            if (m > q_span) m = q_span;
                            sc = m - (tab_ok ? (int)s_gap[(int)d] : nd_gap_cost((int)d, avg)) + fj;
                            cand = true;
                        }
                    }
                }
                // speculative marks (chain.c:82), then read this lane's own mark (chain.c:78)
                if (cand && pj >= 0) { if (i - pj <= NDW_RING) R.t[pj & (NDW_RING - 1)] = i; else t[pj] = i; }
                __syncwarp();
                bool mark = false;
                if (cand) mark = (near ? R.t[j & (NDW_RING - 1)] : t[j]) == i;
                const unsigned candm = __ballot_sync(NDW_FULL, cand);
                if (candm == 0) continue;
                // records of sc against the running maximum, right to left
                unsigned Nm = 0;
                {
                    int cut = 32;
                    for (;;) {
                        int val = (cand && lane < cut) ? sc : INT_MIN;
                        int mx = __reduce_max_sync(NDW_FULL, val);
                        if (mx <= max_f) break;
                        unsigned eq = __ballot_sync(NDW_FULL, val == mx);
                        int first = __ffs(eq) - 1;
                        Nm |= 1u << first; cut = first;
                        if (first == 0) break;
                    }
                }
                const unsigned Mm = __ballot_sync(NDW_FULL, cand && mark) & ~Nm;
                // replay n_skip over the lanes (chain.c:75-80)
                int brk = -1;
                {
                    unsigned ev = Nm; int lo = 0;
                    for (;;) {
                        int nx = ev ? __ffs(ev) - 1 : 32;
                        unsigned range = (nx >= 32 ? NDW_FULL : ((1u << nx) - 1)) & ~((lo >= 32) ? NDW_FULL : ((1u << lo) - 1));
                        unsigned seg = Mm & range;
                        int c = __popc(seg);
                        if (n_skip + c > par.max_skip) { brk = (int)__fns(seg, 0, par.max_skip - n_skip + 1); break; }
                        n_skip += c;
                        if (nx >= 32) break;
                        if (n_skip > 0) --n_skip;
                        ev &= ev - 1; lo = nx + 1;
                    }
                }
                const unsigned vis = brk >= 0 ? ((1u << brk) - 1) : NDW_FULL;
                const unsigned Nv = Nm & vis;
                if (Nv) {
                    int last = 31 - __clz((int)Nv);
                    max_f = __shfl_sync(NDW_FULL, sc, last);
                    max_j = jb - last;
                }
                if (brk >= 0) stop = true;
            }
            // publish anchor i
            if (lane == 0) {
                int32_t vi = max_f;
                if (max_j >= 0) {
                    int32_t vm = (i - max_j <= NDW_RING) ? R.v[max_j & (NDW_RING - 1)] : v[max_j];
                    if (vm > max_f) vi = vm;
                }
                int k = i & (NDW_RING - 1);
                R.x[k] = xi; R.q[k] = qi; R.f[k] = max_f; R.p[k] = max_j; R.v[k] = vi;
                f[i] = max_f; p[i] = max_j; v[i] = vi;
            }
            __syncwarp();
            ai = a_next;
        }
        __syncwarp();
      }
    }
}

// ------------------------------------------------------------------------------------------------
// exact flag-sort emulation, breadth first.  ref: minimap2/ksort.h:98-151.
//
// A job is one call of rs_sort(): a bucket [beg, beg+n) and a byte position.  The reference's
// cycle-leader permutation is sequential by nature (which region is popped next depends on the digit
// just popped), and its order of moves decides the final order of equal keys, so it has to be
// replayed -- but only its *control*: one warp
//   1. extracts the digit of every element into a byte array and histograms it (all lanes);
//   2. derives src_of[dst] (which original slot ends in slot dst):
//        - one non-empty bucket: identity (the pass is a no-op);
//        - two non-empty buckets: closed form from two rank scans (all lanes), see below;
//        - otherwise lane 0 replays the pops/pushes on the 1-byte digits only (L1-resident stream
//          per bucket, bucket heads in shared memory), writing src_of with fire-and-forget stores;
//   3. gathers the 16-byte anchors into the other buffer (all lanes, no dependent chain);
//   4. re-queues child buckets > 64 for the next byte, insertion-sorts the rest in place (one lane
//      per child bucket) and lands finished ranges in buffer A.
//
// Two-bucket closed form.  Buckets d0 < d1 with regions R0 = [0, n0), R1 = [n0, n).  m_i = i-th
// element of R0 with digit d1, z_i = i-th element of R1 with digit d0 (K of each).  Replaying the
// reference: each m_i is dropped at R1's head and pushes the natives in front of it forward until a
// z pops out, which fills m_i's hole.  Hence R0 keeps its natives and hole i receives z_i; in R1,
// slot n0 receives m_0, a slot following z_r (r < K-1) receives m_{r+1}, every other slot up to the
// last z receives its left neighbour, slots after the last z are untouched.
// ------------------------------------------------------------------------------------------------
#define NDW_SORT_WARPS 4

struct ndw_sort_bufs {
    nd_u128 *A, *B;            // A holds the input and the final result
    uint8_t *dig;              // digit per element
    uint32_t *src_of, *tmp;    // permutation, rank lists
};

// Fast path (SURVEY H1 option ii): the tie order of the reference's unstable sort is observable only among anchors with EQUAL x
// (same target position hit from several query positions).  A read whose anchors all have distinct x has exactly one sorted order,
// so any correct sort will do: one block sorts the read's (x, index) pairs in shared memory (bitonic network), checks neighbours
// for equal x, and only a read that does have duplicates is handed to the exact replay below, untouched.
#define NDW_FAST_MAX 4096
#define NDW_FAST_T 256

// Group filter in front of the block sort.  An anchor can only end in a chain if its (strand, target read) group -- the high 32 bits
// of x -- holds at least min_cnt anchors whose spans could add up to min_chain_score (the same test nd_seg_bounds applies per segment;
// a group is a superset of its segments).  Most anchors of a read against a large index are stray k-mer hits in groups of one or two
// (their number grows with the index: +0.8 M per step for every doubling at k = 17), and the sort is the stage they cost most in.
// Dropping them is exact when the read's remaining anchors have distinct x: their sorted order is then unique, groups never share a
// segment (different high bits are more than max_gap apart), and the predecessor window of chain.c:50 never crosses a group.  If the
// kept anchors do hold equal x the read goes to the exact replay with ALL its anchors, untouched.  One block per read counts the groups
// in a shared-memory hash table; dropped anchors are replaced, behind the sorted ones, by placeholders that are each a segment of one
// (x above every real target, more than max_gap apart), which the chain kernels skip.
#define NDW_FILT_HS 4096
#define NDW_FILT_MAXN 16384
#define NDW_FILT_EMPTY 0xffffffffu
__global__ void __launch_bounds__(NDW_FAST_T)
k_group_filter(const nd_u128 *a, const nd_u64 *a_off, int n_reads, int min_cnt, int min_sc, int max_span, uint32_t max_gap, uint8_t *keep, uint32_t *n_kept)
{
    __shared__ uint32_t hk[NDW_FILT_HS], hc[NDW_FILT_HS];
    __shared__ int s_fail;
    __shared__ uint32_t s_distinct, s_kept;
    const int r = blockIdx.x, tid = threadIdx.x;
    const nd_u64 b = a_off[r], n64 = a_off[r + 1] - b;
    if (n64 <= 1) { if (tid == 0) { n_kept[r] = (uint32_t)n64; if (n64) keep[b] = 1; } return; }
    const bool can = n64 <= NDW_FILT_MAXN && n64 * ((nd_u64)max_gap + 1) < (1ULL << 32);
    const uint32_t n = can ? (uint32_t)n64 : 0;
    for (int i = tid; i < NDW_FILT_HS; i += NDW_FAST_T) { hk[i] = NDW_FILT_EMPTY; hc[i] = 0; }
    if (tid == 0) { s_fail = can ? 0 : 1; s_distinct = 0; s_kept = 0; }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += NDW_FAST_T) {
        const uint32_t key = (uint32_t)(a[b + i].x >> 32);
        if (key == NDW_FILT_EMPTY) { s_fail = 1; continue; }
        uint32_t h = (key * 2654435761u) >> 20;
        for (int tries = 0;; ++tries) {
            const uint32_t old = atomicCAS(&hk[h], NDW_FILT_EMPTY, key);
            if (old == NDW_FILT_EMPTY) { if (atomicAdd(&s_distinct, 1u) >= NDW_FILT_HS * 3 / 4) s_fail = 1; }
            if (old == NDW_FILT_EMPTY || old == key) { atomicAdd(&hc[h], 1u); break; }
            h = (h + 1) & (NDW_FILT_HS - 1);
            if (tries >= NDW_FILT_HS) { s_fail = 1; break; }
        }
    }
    __syncthreads();
    if (s_fail) {                                                       // no filtering for this read: everything is kept
        for (nd_u64 i = tid; i < n64; i += NDW_FAST_T) keep[b + i] = 1;
        if (tid == 0) n_kept[r] = n64 > 0xfffffff0ull ? 0xfffffff0u : (uint32_t)n64;
        return;
    }
    uint32_t mine = 0;
    for (uint32_t i = tid; i < n; i += NDW_FAST_T) {
        const uint32_t key = (uint32_t)(a[b + i].x >> 32);
        uint32_t h = (key * 2654435761u) >> 20;
        while (hk[h] != key) h = (h + 1) & (NDW_FILT_HS - 1);
        const uint32_t c = hc[h];
        const bool k = (int)c >= min_cnt && (nd_i64)c * max_span >= (nd_i64)min_sc;
        keep[b + i] = k ? 1 : 0;
        mine += k;
    }
    for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((tid & 31) == 0 && mine) atomicAdd(&s_kept, mine);
    __syncthreads();
    if (tid == 0) n_kept[r] = s_kept;
}

// Fast path (SURVEY H1 option ii): the tie order of the reference's unstable sort is observable only among anchors with EQUAL x
// (same target position hit from several query positions).  A read whose (kept) anchors all have distinct x has exactly one sorted
// order, so any correct sort will do: one block sorts the read's (x, index) pairs in shared memory (bitonic network), checks neighbours
// for equal x, and only a read that does have duplicates is handed to the exact replay below, untouched.
template <int CAPF>
__global__ void __launch_bounds__(NDW_FAST_T)
k_sort_fast(nd_u128 *a, const nd_u64 *a_off, int n_reads, int lo_excl, const uint8_t *keep, const uint32_t *n_kept, uint32_t max_gap, nd_rs_job *jobs, uint32_t *n_jobs)
{
    extern __shared__ unsigned char ndw_fast_smem[];
    nd_u64 *sx = reinterpret_cast<nd_u64 *>(ndw_fast_smem);          // CAPF keys
    nd_u64 *sy = sx + CAPF;                                           // CAPF payloads (y)
    uint16_t *si = reinterpret_cast<uint16_t *>(sy + CAPF);           // CAPF indices
    __shared__ int s_dup;
    __shared__ uint32_t s_w[NDW_FAST_T / 32];
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const nd_u64 b = a_off[r], n64 = a_off[r + 1] - b;
    const long long nk64 = keep ? (long long)n_kept[r] : (long long)n64;
    if (nk64 <= (long long)lo_excl || nk64 > (long long)CAPF) return;
    if (n64 < 2) return;
    const int n = (int)n64, nk = (int)nk64;
    int m = 1; while (m < nk) m <<= 1;                                // pad to a power of two with the maximum key
    if (tid == 0) s_dup = 0;
    if (!keep || nk == n) {
        for (int i = tid; i < n; i += NDW_FAST_T) {
            const uint4 v = *reinterpret_cast<const uint4 *>(a + b + i);
            sx[i] = (nd_u64)v.y << 32 | v.x; sy[i] = (nd_u64)v.w << 32 | v.z; si[i] = (uint16_t)i;
        }
        __syncthreads();
    } else {                                                          // load the kept anchors only (block-wide compaction, 256 at a time)
        int base = 0;
        for (int i0 = 0; i0 < n; i0 += NDW_FAST_T) {
            const int i = i0 + tid;
            const bool f = i < n && keep[b + i];
            uint4 v = make_uint4(0, 0, 0, 0);
            if (f) v = *reinterpret_cast<const uint4 *>(a + b + i);
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (lane == 0) s_w[wid] = (uint32_t)__popc(bal);
            __syncthreads();
            int before = 0, total = 0;
#pragma unroll
            for (int q = 0; q < NDW_FAST_T / 32; ++q) { const int cq = (int)s_w[q]; before += q < wid ? cq : 0; total += cq; }
            if (f) { const int pos = base + before + __popc(bal & ((1u << lane) - 1)); sx[pos] = (nd_u64)v.y << 32 | v.x; sy[pos] = (nd_u64)v.w << 32 | v.z; si[pos] = (uint16_t)pos; }
            base += total;
            __syncthreads();
        }
    }
    for (int i = nk + tid; i < m; i += NDW_FAST_T) { sx[i] = ND_U64_MAX; si[i] = (uint16_t)i; }
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (m >> 1); t += NDW_FAST_T) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i | j;       // the pair (i, i + j) of this stage
                const bool up = (i & k) == 0;
                const nd_u64 xi = sx[i], xp = sx[p];
                if ((xi > xp) == up) { sx[i] = xp; sx[p] = xi; const uint16_t ti = si[i]; si[i] = si[p]; si[p] = ti; }
            }
            __syncthreads();
        }
    int dup = 0;
    for (int i = tid; i + 1 < nk; i += NDW_FAST_T) dup |= sx[i] == sx[i + 1];
    if (dup) s_dup = 1;
    __syncthreads();
    if (s_dup) {                                                      // equal keys: the exact emulation decides their order -- over ALL anchors of the read
        if (n <= ND_RS_SMALL) { if (tid == 0) nd_ins_sort_128x(a + b, a + b + n); return; }   // the reference insertion-sorts these (ksort.h:147)
        if (tid == 0) { const uint32_t q = atomicAdd(n_jobs, 1u); jobs[q].beg = (uint32_t)b; jobs[q].n = (uint32_t)n; jobs[q].shift = 56; jobs[q].buf = 0; }
        return;
    }
    for (int i = tid; i < nk; i += NDW_FAST_T) {
        const nd_u64 x = sx[i], y = sy[si[i]];
        *reinterpret_cast<uint4 *>(a + b + i) = make_uint4((uint32_t)x, (uint32_t)(x >> 32), (uint32_t)y, (uint32_t)(y >> 32));
    }
    for (int i = nk + tid; i < n; i += NDW_FAST_T) {                  // placeholders for the dropped anchors: one-anchor segments above every real target
        const uint32_t lowx = (uint32_t)(i - nk) * (max_gap + 1u);
        *reinterpret_cast<uint4 *>(a + b + i) = make_uint4(lowx, 0xffffffffu, 0u, 0u);
    }
}

__global__ void k_rs_init(nd_u128 *a, const nd_u64 *a_off, int n_reads, nd_rs_job *jobs, uint32_t *n_jobs, int fast_max, const uint32_t *n_kept)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    nd_u64 b = a_off[r], n = a_off[r + 1] - b;
    if (fast_max > 0 && (n_kept ? (nd_u64)n_kept[r] : n) <= (nd_u64)fast_max) return;   // k_sort_fast takes it: block sort; equal keys are queued there or insertion-sorted
    if (n > ND_RS_SMALL) {
        uint32_t q = atomicAdd(n_jobs, 1u);
        jobs[q].beg = (uint32_t)b; jobs[q].n = (uint32_t)n; jobs[q].shift = 56; jobs[q].buf = 0;
    } else if (n > 1) nd_ins_sort_128x(a + b, a + b + n);
}

__global__ void __launch_bounds__(NDW_SORT_WARPS * 32)
k_rs_level(ndw_sort_bufs W, const nd_rs_job *jobs_in, const uint32_t *n_in, nd_rs_job *jobs_out, uint32_t *n_out, uint32_t *work)
{
    __shared__ uint32_t s_head[NDW_SORT_WARPS][256], s_tail[NDW_SORT_WARPS][256];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t *head = s_head[wib], *tail = s_tail[wib];
    const uint32_t nj = *n_in;
    const unsigned lt = (1u << lane) - 1;
    for (;;) {
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(work, 1u);
        q = __shfl_sync(NDW_FULL, q, 0);
        if (q >= nj) break;
        const nd_rs_job job = jobs_in[q];
        const nd_u128 *in = (job.buf ? W.B : W.A) + job.beg;
        nd_u128 *out = (job.buf ? W.A : W.B) + job.beg;
        const bool out_is_A = job.buf != 0;
        uint8_t *dig = W.dig + job.beg;
        uint32_t *src_of = W.src_of + job.beg, *tmp = W.tmp + job.beg;
        const int s = job.shift;
        const uint32_t n = job.n;

        // 1. digits + histogram
        for (int d = lane; d < 256; d += 32) head[d] = 0;
        __syncwarp();
        for (uint32_t i0 = lane; i0 < n; i0 += 128) {               // 4 independent loads in flight per lane
            nd_u64 xv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { const uint32_t i = i0 + 32 * k; xv[k] = i < n ? in[i].x : 0; }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t i = i0 + 32 * k;
                if (i < n) { const uint32_t d = (uint32_t)(xv[k] >> s & 255); dig[i] = (uint8_t)d; src_of[i] = i; atomicAdd(&head[d], 1u); }
            }
        }
        __syncwarp();
        uint32_t c[8], sum = 0, nonempty = 0, dmin = 256, dmax = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            c[k] = head[lane * 8 + k]; sum += c[k];
            if (c[k]) { ++nonempty; uint32_t d = lane * 8 + k; dmin = d < dmin ? d : dmin; dmax = d > dmax ? d : dmax; }
        }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(NDW_FULL, incl, o); if (lane >= o) incl += y; }
        uint32_t run = incl - sum;
        nonempty = __reduce_add_sync(NDW_FULL, nonempty);
        dmin = __reduce_min_sync(NDW_FULL, dmin); dmax = __reduce_max_sync(NDW_FULL, dmax);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) { head[lane * 8 + k] = run; run += c[k]; tail[lane * 8 + k] = run; }
        __syncwarp();

        if (nonempty == 1) {                       // no-op pass
            if (s) {
                if (lane == 0) { uint32_t o = atomicAdd(n_out, 1u); jobs_out[o] = job; jobs_out[o].shift = s > 8 ? s - 8 : 0; }
            } else if (job.buf) {
                for (uint32_t i = lane; i < n; i += 32) W.A[job.beg + i] = in[i];
            }
            __syncwarp();
            continue;
        }

        // 2. src_of
        if (nonempty == 2) {
            const uint32_t d0 = dmin, d1 = dmax, n0 = tail[d0];
            uint32_t K = 0;
            for (uint32_t b0 = 0; b0 < n0; b0 += 128) {               // m list
                uint32_t dg[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { const uint32_t i = b0 + 32 * k + lane; dg[k] = i < n0 ? dig[i] : 256u; }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t i = b0 + 32 * k + lane; const bool fl = dg[k] == d1;
                    const unsigned m = __ballot_sync(NDW_FULL, fl);
                    if (fl) tmp[K + __popc(m & lt)] = i;
                    K += __popc(m);
                }
            }
            uint32_t K1 = 0;
            for (uint32_t b0 = n0; b0 < n; b0 += 128) {               // z list
                uint32_t dg[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { const uint32_t i = b0 + 32 * k + lane; dg[k] = i < n ? dig[i] : 256u; }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t i = b0 + 32 * k + lane; const bool fl = dg[k] == d0;
                    const unsigned m = __ballot_sync(NDW_FULL, fl);
                    if (fl) tmp[n0 + K1 + __popc(m & lt)] = i;
                    K1 += __popc(m);
                }
            }
            __syncwarp();
            if (K) {
                const uint32_t *mpos = tmp, *zpos = tmp + n0;
                const uint32_t zl = zpos[K - 1];
                uint32_t rk = 0;
                for (uint32_t b0 = 0; b0 < n0; b0 += 32) {            // holes of R0 <- z_i
                    uint32_t i = b0 + lane; bool fl = i < n0 && dig[i] == d1;
                    unsigned m = __ballot_sync(NDW_FULL, fl);
                    if (fl) src_of[i] = zpos[rk + __popc(m & lt)];
                    rk += __popc(m);
                }
                rk = 0;
                for (uint32_t b0 = n0; b0 <= zl; b0 += 32) {           // R1 up to the last z
                    uint32_t j = b0 + lane; bool inr = j < zl;         // slot j+1 <= zl gets a new occupant
                    bool fl = j <= zl && dig[j] == d0;
                    unsigned m = __ballot_sync(NDW_FULL, fl);
                    if (inr) src_of[j + 1] = fl ? mpos[rk + __popc(m & lt) + 1] : j;
                    rk += __popc(m);
                }
                if (lane == 0) src_of[n0] = mpos[0];
            }
        } else if (lane == 0) {
            for (uint32_t d = 0; d < 256; ++d) {
                uint32_t hd = head[d];
                const uint32_t td = tail[d];
                while (hd < td) {
                    uint32_t t = dig[hd];
                    if (t == d) { ++hd; continue; }
                    uint32_t carry = hd;
                    do {
                        const uint32_t ht = head[t];
                        const uint32_t x = dig[ht];
                        src_of[ht] = carry; carry = ht; head[t] = ht + 1;
                        t = x;
                    } while (t != d);
                    src_of[hd] = carry; ++hd;
                }
                head[d] = hd;
            }
        }
        __syncwarp();
        __threadfence_block();

        // 3. gather into the other buffer
        for (uint32_t i0 = lane; i0 < n; i0 += 128) {
            uint32_t sidx[4]; uint4 vv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { const uint32_t i = i0 + 32 * k; sidx[k] = i < n ? src_of[i] : 0; }
#pragma unroll
            for (int k = 0; k < 4; ++k) vv[k] = *reinterpret_cast<const uint4 *>(in + sidx[k]);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const uint32_t i = i0 + 32 * k; if (i < n) *reinterpret_cast<uint4 *>(out + i) = vv[k]; }
        }
        __syncwarp();

        // 4. children
        if (s) {
            const int ns = s > 8 ? s - 8 : 0;
            for (int d = lane; d < 256; d += 32) {
                const uint32_t bb = d ? tail[d - 1] : 0, ee = tail[d];
                if (ee - bb > ND_RS_SMALL) {
                    uint32_t o = atomicAdd(n_out, 1u);
                    jobs_out[o].beg = job.beg + bb; jobs_out[o].n = ee - bb; jobs_out[o].shift = ns; jobs_out[o].buf = job.buf ^ 1;
                } else if (ee > bb) {
                    if (ee - bb > 1) nd_ins_sort_128x(out + bb, out + ee);
                    if (!out_is_A) for (uint32_t i = bb; i < ee; ++i) W.A[job.beg + i] = out[i];
                }
            }
        } else if (!out_is_A) {
            for (uint32_t i = lane; i < n; i += 32) W.A[job.beg + i] = out[i];
        }
        __syncwarp();
    }
}

// host-side sequencing ------------------------------------------------------------------------
struct ndw_launch_hooks { void (*begin)(void *ctx, const char *name); void (*end)(void *ctx); void *ctx; };

// jobs: device buffer for 2 * (NA/65 + n_reads + 1) jobs; counters: 24 u32 on the device; keep / n_kept: scratch of the group filter
// (n_a bytes, n_reads u32) or null to sort every anchor
struct ndw_filter_par { uint8_t *keep; uint32_t *n_kept; int min_cnt, min_sc, max_span; uint32_t max_gap; };
static inline void ndw_sort_launch(cudaStream_t st, ndw_sort_bufs W, const nd_u64 *a_off, int n_reads, nd_u64 n_a,
                                   nd_rs_job *jobs, uint32_t *counters, int n_sm, ndw_launch_hooks hk, bool fast_path, ndw_filter_par F)
{
    const nd_u64 half = n_a / (ND_RS_SMALL + 1) + (nd_u64)n_reads + 1;
    nd_rs_job *buf[2] = { jobs, jobs + half };
    cudaMemsetAsync(counters, 0, 24 * sizeof(uint32_t), st);
    const bool filt = fast_path && F.keep && F.n_kept;
    if (filt) {
        hk.begin(hk.ctx, "sort_filter");
        k_group_filter<<<n_reads, NDW_FAST_T, 0, st>>>(W.A, a_off, n_reads, F.min_cnt, F.min_sc, F.max_span, F.max_gap, F.keep, F.n_kept);
        hk.end(hk.ctx);
    }
    hk.begin(hk.ctx, "sort_init");
    k_rs_init<<<(n_reads + 127) / 128, 128, 0, st>>>(W.A, a_off, n_reads, buf[0], counters, fast_path ? NDW_FAST_MAX : 0, filt ? F.n_kept : nullptr);
    hk.end(hk.ctx);
    if (fast_path) {                                   // distinct-key reads: block sort in shared memory, four size classes
        cudaFuncSetAttribute(k_sort_fast<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 18);   // per device: cheap, idempotent
        const uint8_t *kp = filt ? F.keep : nullptr; const uint32_t *nk = filt ? F.n_kept : nullptr;
        hk.begin(hk.ctx, "sort_fast");
        k_sort_fast<512><<<n_reads, NDW_FAST_T, 512 * 18, st>>>(W.A, a_off, n_reads, -1, kp, nk, F.max_gap, buf[0], counters);
        k_sort_fast<1024><<<n_reads, NDW_FAST_T, 1024 * 18, st>>>(W.A, a_off, n_reads, 512, kp, nk, F.max_gap, buf[0], counters);
        k_sort_fast<2048><<<n_reads, NDW_FAST_T, 2048 * 18, st>>>(W.A, a_off, n_reads, 1024, kp, nk, F.max_gap, buf[0], counters);
        k_sort_fast<4096><<<n_reads, NDW_FAST_T, 4096 * 18, st>>>(W.A, a_off, n_reads, 2048, kp, nk, F.max_gap, buf[0], counters);
        hk.end(hk.ctx);
    }
    for (int lvl = 0; lvl < 8; ++lvl) {
        hk.begin(hk.ctx, "sort_level");
        k_rs_level<<<n_sm * 8, NDW_SORT_WARPS * 32, 0, st>>>(W, buf[lvl & 1], counters + lvl, buf[(lvl + 1) & 1], counters + lvl + 1, counters + 12 + lvl);
        hk.end(hk.ctx);
    }
}

static inline void ndw_chain_fill_launch(cudaStream_t st, nd_seg_view S, nd_reads_dev Q, const float *avg_qspan, nd_chain_par par,
                                         int max_span, int uniform_span, nd_chain_bufs B, uint32_t *work_counter, int n_sm)
{
    cudaMemsetAsync(work_counter, 0, sizeof(uint32_t), st);
    unsigned blocks = (unsigned)(n_sm * 16);           // 64 warps per SM: the kernel is latency-bound (serial dependence over a segment's anchors)
    unsigned need = (S.n_seg + NDW_CHAIN_WARPS * 32 - 1) / (NDW_CHAIN_WARPS * 32);   // a warp claims 32 segments at a time
    if (need < blocks) blocks = need ? need : 1;
    k_chain_fill_warp<<<blocks, NDW_CHAIN_WARPS * 32, 0, st>>>(S, Q, avg_qspan, par, max_span, uniform_span, B, work_counter);
}
