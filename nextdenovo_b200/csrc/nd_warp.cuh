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
#define NDW_RING 256
#define NDW_CHAIN_WARPS 4

struct ndw_ring {
    uint32_t x[NDW_RING];
    int32_t q[NDW_RING], f[NDW_RING], p[NDW_RING], v[NDW_RING], t[NDW_RING];
};

__global__ void __launch_bounds__(NDW_CHAIN_WARPS * 32)
k_chain_fill_warp(nd_seg_view S, nd_reads_dev Q, const float *avg_qspan, nd_chain_par par, int max_span,
                  nd_chain_bufs B, uint32_t *work)
{
    __shared__ ndw_ring rings[NDW_CHAIN_WARPS];
    const int lane = threadIdx.x & 31;
    ndw_ring &R = rings[threadIdx.x >> 5];
    const nd_u64 mdx = (nd_u64)par.max_dist_x;

    for (;;) {
        uint32_t s = 0;
        if (lane == 0) s = atomicAdd(work, 1u);
        s = __shfl_sync(NDW_FULL, s, 0);
        if (s >= S.n_seg) break;
        nd_u64 b; int n;
        if (!nd_seg_bounds(S, s, par, max_span, &b, &n)) continue;
        const int r = nd_find_owner(S.a_off, Q.n_reads, b);
        const float avg = avg_qspan[r];
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
                            sc = m - nd_gap_cost((int)d, avg) + fj;
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

// ------------------------------------------------------------------------------------------------
// exact flag-sort emulation, breadth first.  ref: minimap2/ksort.h:98-151.
// A job is one call of rs_sort(): a bucket [beg, beg+n) and a byte position.  A pass over a bucket
// whose elements all share the digit is a no-op and only re-queues the bucket; otherwise the
// cycle-leader permutation is replayed by lane 0 (its order of moves is what makes the reference's
// tie order what it is), while histogramming, re-queueing and the <=64 insertion sorts of the child
// buckets run across the lanes.
// ------------------------------------------------------------------------------------------------
#define NDW_SORT_WARPS 4

__global__ void k_rs_init(nd_u128 *a, const nd_u64 *a_off, int n_reads, nd_rs_job *jobs, uint32_t *n_jobs)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    nd_u64 b = a_off[r], n = a_off[r + 1] - b;
    if (n > ND_RS_SMALL) {
        uint32_t q = atomicAdd(n_jobs, 1u);
        jobs[q].beg = (uint32_t)b; jobs[q].n = (uint32_t)n; jobs[q].shift = 56;
    } else if (n > 1) nd_ins_sort_128x(a + b, a + b + n);
}

__global__ void __launch_bounds__(NDW_SORT_WARPS * 32)
k_rs_level(nd_u128 *a, const nd_rs_job *jobs_in, const uint32_t *n_in, nd_rs_job *jobs_out, uint32_t *n_out)
{
    __shared__ uint32_t s_head[NDW_SORT_WARPS][256], s_tail[NDW_SORT_WARPS][256];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t *head = s_head[wib], *tail = s_tail[wib];
    const uint32_t nj = *n_in;
    const uint32_t wid = blockIdx.x * NDW_SORT_WARPS + wib, nw = gridDim.x * NDW_SORT_WARPS;
    for (uint32_t q = wid; q < nj; q += nw) {
        const nd_rs_job job = jobs_in[q];
        nd_u128 *base = a + job.beg;
        const int s = job.shift;
        for (int d = lane; d < 256; d += 32) head[d] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < job.n; i += 32) atomicAdd(&head[base[i].x >> s & 255], 1u);
        __syncwarp();
        // exclusive prefix over 256 counters: 8 per lane + warp scan
        uint32_t c[8], sum = 0, nonempty = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { c[k] = head[lane * 8 + k]; sum += c[k]; nonempty += c[k] != 0; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(NDW_FULL, incl, o); if (lane >= o) incl += y; }
        uint32_t run = incl - sum;
        nonempty = __reduce_add_sync(NDW_FULL, nonempty);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) { head[lane * 8 + k] = run; run += c[k]; tail[lane * 8 + k] = run; }
        __syncwarp();
        if (nonempty > 1 && lane == 0) {
            for (uint32_t d = 0; d < 256;) {
                uint32_t hd = head[d];
                if (hd == tail[d]) { ++d; continue; }
                uint32_t t = (uint32_t)(base[hd].x >> s & 255);
                if (t == d) { head[d] = hd + 1; continue; }
                nd_u128 carry = base[hd];
                do {
                    uint32_t ht = head[t];
                    nd_u128 evicted = base[ht];
                    base[ht] = carry; head[t] = ht + 1;
                    carry = evicted;
                    t = (uint32_t)(carry.x >> s & 255);
                } while (t != d);
                base[hd] = carry; head[d] = hd + 1;
            }
        }
        __syncwarp();
        if (s) {
            const int ns = s > 8 ? s - 8 : 0;
            for (int d = lane; d < 256; d += 32) {
                uint32_t bb = d ? tail[d - 1] : 0, ee = tail[d];
                if (ee - bb > ND_RS_SMALL) {
                    uint32_t o = atomicAdd(n_out, 1u);
                    jobs_out[o].beg = job.beg + bb; jobs_out[o].n = ee - bb; jobs_out[o].shift = ns;
                } else if (ee - bb > 1) nd_ins_sort_128x(base + bb, base + ee);
            }
        }
        __syncwarp();
    }
}

// host-side sequencing ------------------------------------------------------------------------
struct ndw_launch_hooks { void (*begin)(void *ctx, const char *name); void (*end)(void *ctx); void *ctx; };

// jobs: device buffer for 2 * (NA/65 + n_reads + 1) jobs; counters: 10 u32 on the device
static inline void ndw_sort_launch(cudaStream_t st, nd_u128 *a, const nd_u64 *a_off, int n_reads, nd_u64 n_a,
                                   nd_rs_job *jobs, uint32_t *counters, int n_sm, ndw_launch_hooks hk)
{
    const nd_u64 half = n_a / (ND_RS_SMALL + 1) + (nd_u64)n_reads + 1;
    nd_rs_job *buf[2] = { jobs, jobs + half };
    cudaMemsetAsync(counters, 0, 10 * sizeof(uint32_t), st);
    hk.begin(hk.ctx, "sort_init");
    k_rs_init<<<(n_reads + 127) / 128, 128, 0, st>>>(a, a_off, n_reads, buf[0], counters);
    hk.end(hk.ctx);
    for (int lvl = 0; lvl < 8; ++lvl) {
        hk.begin(hk.ctx, "sort_level");
        k_rs_level<<<n_sm * 8, NDW_SORT_WARPS * 32, 0, st>>>(a, buf[lvl & 1], counters + lvl, buf[(lvl + 1) & 1], counters + lvl + 1);
        hk.end(hk.ctx);
    }
}

static inline void ndw_chain_fill_launch(cudaStream_t st, nd_seg_view S, nd_reads_dev Q, const float *avg_qspan, nd_chain_par par,
                                         int max_span, nd_chain_bufs B, uint32_t *work_counter, int n_sm)
{
    cudaMemsetAsync(work_counter, 0, sizeof(uint32_t), st);
    unsigned blocks = (unsigned)(n_sm * 8);
    unsigned need = (S.n_seg + NDW_CHAIN_WARPS - 1) / NDW_CHAIN_WARPS;
    if (need < blocks) blocks = need ? need : 1;
    k_chain_fill_warp<<<blocks, NDW_CHAIN_WARPS * 32, 0, st>>>(S, Q, avg_qspan, par, max_span, B, work_counter);
}
