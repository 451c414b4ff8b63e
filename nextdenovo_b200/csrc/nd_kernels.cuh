// nd_kernels.cuh -- sm_100a kernels of the overlap hot path (baseline generation: one thread per
// chunk / minimizer / read / segment; the warp-cooperative generation lives in nd_warp.cuh).
#pragma once
#include <cuda_runtime.h>
#include "nd_core.cuh"

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------

// last index r in [0,n) with off[r] <= v   (off is non-decreasing, off[0] <= v)
__device__ __forceinline__ int nd_find_owner(const nd_u64 *off, int n, nd_u64 v)
{
    int lo = 0, hi = n;            // invariant: off[lo] <= v, (hi == n or off[hi] > v)
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------
// index table.  Replaces the khash buckets of minimap2/index.c:191-243 (and mm_idx_get, :81-98).
//
// Every key (minimizer hash, group tag above it for a tagged index) is first mixed bijectively on key_bits bits (nd_idx_mix:
// minimizer hashes themselves crowd towards 0, being window minima), so mixed keys are spread uniformly over [0, 2^key_bits).
// The table is ORDERED by mixed key: slot(key) = ideal(key) + displacement, with ideal() monotone and keys stored in increasing
// order ("sorted array with gaps", load ~0.7).  Consequences:
//   * it is built from the sorted keys without atomics: slot_i = i + max_{j<=i}(ideal_j - j), one prefix-max scan;
//   * a lookup reads consecutive 16-byte slots from ideal(key) until it meets a key >= its own (empty slots hold the maximum
//     key), so hits and misses both cost ~2 slots of one or two adjacent 32-byte sectors;
//   * it is range-partitioned: slice r holds the keys with mulhi(key * key_mul, n_slices) == r in its own contiguous slot range
//     (slice_cap ideal slots + a pad for the displaced tail), so N GPUs build one slice each and all-gather equal-sized pieces.
// Slot value: a minimizer that occurs once carries its position inline (bit 63 set; rid < 2^31 keeps that bit free, as in the
// reference's singleton trick, index.c:91-93,223-226); otherwise off << 24 | min(n, 2^24-1) into the slice's position list.
// ------------------------------------------------------------------------------------------------
struct nd_ht_slot { nd_u64 key, val; };
#define ND_HT_EMPTY ND_U64_MAX
#define ND_HT_INLINE (1ULL << 63)
#define ND_HT_PAD 4096ULL

struct nd_idx_view {
    const nd_ht_slot *tab;         // n_slices * slice_slots slots
    const nd_u64 *pos;             // position lists (sorted by key, ascending position inside a key: index.c:230)
    const nd_u64 *pos_base;        // n_slices entries: where slice r's list starts in pos[]
    nd_u64 slice_cap, slice_slots; // ideal slots per slice; slice_slots = slice_cap + ND_HT_PAD
    nd_u64 key_mul;                // floor((2^64 - 1) / 2^key_bits): key * key_mul is a 64-bit fixed-point fraction in [0, 1)
    uint32_t n_slices; int key_bits;
};

// returns n (0 = absent); *v = inline position (n == 1 && *inl) or offset into pos[]
__device__ __forceinline__ uint32_t nd_idx_lookup(const nd_idx_view &I, nd_u64 raw_key, nd_u64 *v, bool *inl)
{
    const nd_u64 key = nd_idx_mix(raw_key, I.key_bits);
    const nd_u64 t = key * I.key_mul;
    const uint32_t sl = (uint32_t)__umul64hi(t, (nd_u64)I.n_slices);
    const nd_ht_slot *p = I.tab + (nd_u64)sl * I.slice_slots + __umul64hi(t * (nd_u64)I.n_slices, I.slice_cap);
    for (;;) {
        const ulonglong2 e = __ldg(reinterpret_cast<const ulonglong2 *>(p));
        if (e.x >= key) {
            if (e.x != key) return 0;
            if (e.y & ND_HT_INLINE) { *v = e.y & ~ND_HT_INLINE; *inl = true; return 1; }
            *v = I.pos_base[sl] + (e.y >> 24); *inl = false;
            return (uint32_t)(e.y & 0xffffff);
        }
        ++p;
    }
}

// ------------------------------------------------------------------------------------------------
// sketch
// ------------------------------------------------------------------------------------------------

struct nd_reads_dev {
    const uint32_t *words;
    const nd_u64 *word_off;
    const uint32_t *len;
    const uint32_t *nmask;
    const nd_u64 *name_key;
    const uint32_t *name_hash;
    const uint32_t *name_id;
    const uint32_t *tag;           // optional group tag per read
    int n_reads;
};

// one thread per chunk of C bases; chunk_off[r] = index of read r's first chunk
template <int RW, bool LONGK>
__global__ void __launch_bounds__(128)
k_sketch_chunks(nd_reads_dev R, const nd_u64 *chunk_off, nd_u64 n_chunks, int C, int w, int k, int hpc,
                uint32_t rid0, int use_rid, nd_u128 *staging, int S, uint32_t *cnt, int *overflow, const uint32_t *slow)
{
    nd_u64 g = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_chunks) return;
    int r = nd_find_owner(chunk_off, R.n_reads, g);
    if (slow && !slow[r]) return;                       // this read was sketched by k_sketch_fast
    int len = (int)R.len[r];
    int cs = (int)(g - chunk_off[r]) * C;
    int ce = cs + C < len ? cs + C : len;
    nd_seqview sv;
    sv.words = R.words + R.word_off[r];
    sv.nmask = R.nmask ? R.nmask + (R.word_off[r] >> 1) : nullptr;
    sv.len = len;
    int n = nd_sketch_chunk<RW, LONGK>(sv, w, k, hpc, use_rid ? rid0 + (uint32_t)r : 0u, cs, ce, staging + g * (nd_u64)S, S);
    if (n > S) { atomicExch(overflow, 1); n = S; }
    cnt[g] = (uint32_t)n;
}

// Fast path: one block per 512-base chunk, closed form of the window state machine (nd_core.cuh).
// Hashes of the chunk plus a (w-1) halo are computed once into shared memory (one k-mer per thread
// step, k-mers cut out of the packed words by funnel shifts), every position then decides locally
// whether it is pushed, and the block compacts the survivors in position order.  Reads that break a
// precondition (symmetric k-mer, too short for one window) are flagged in slow[] and redone by
// k_sketch_chunks.
#define NDS_C 512
#define NDS_T 128
#define NDS_MAXW 64
__global__ void __launch_bounds__(NDS_T)
k_sketch_fast(nd_reads_dev R, const nd_u64 *chunk_off, nd_u64 n_chunks, int w, int k, uint32_t rid0, int use_rid,
              nd_u128 *staging, int S, uint32_t *cnt, uint32_t *slow)
{
    __shared__ nd_u64 sx[NDS_C + 2 * NDS_MAXW];
    __shared__ uint8_t sz[NDS_C + 2 * NDS_MAXW];
    __shared__ uint32_t swords[(NDS_C + 2 * NDS_MAXW) / 16 + 8];
    __shared__ int s_r;
    __shared__ uint32_t s_wsum[NDS_T / 32];
    const nd_u64 g = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_r = nd_find_owner(chunk_off, R.n_reads, g);
    __syncthreads();
    const int r = s_r, len = (int)R.len[r];
    const int cs = (int)(g - chunk_off[r]) * NDS_C, ce = cs + NDS_C < len ? cs + NDS_C : len;
    if (len < k + w - 1) { if (tid == 0) { slow[r] = 1; cnt[g] = 0; } return; }
    const int lo = cs - (w - 1) > 0 ? cs - (w - 1) : 0, hi = ce + (w - 1) < len ? ce + (w - 1) : len;
    const uint32_t *words = R.words + R.word_off[r];
    const int w0 = (lo >> 4) - 2 > 0 ? (lo >> 4) - 2 : 0, w1 = (hi - 1) >> 4;
    for (int i = tid; i <= w1 - w0; i += NDS_T) swords[i] = words[w0 + i];
    __syncthreads();
    int any_sym = 0;
    for (int idx = tid; idx < hi - lo; idx += NDS_T) {
        int z; bool sym;
        sx[idx] = nd_kmer_x(swords - w0, lo + idx, k, &z, &sym);
        sz[idx] = (uint8_t)z;
        any_sym |= sym;
    }
    if (__syncthreads_or(any_sym)) { if (tid == 0) { slow[r] = 1; cnt[g] = 0; } return; }
    const nd_u64 ridbits = (nd_u64)(use_rid ? rid0 + (uint32_t)r : 0u) << 32;
    nd_u128 *dst = staging + g * (nd_u64)S;
    uint32_t base = 0;
    for (int j = 0; j < NDS_C / NDS_T; ++j) {
        const int t = cs + j * NDS_T + tid;
        bool flag = false;
        if (t < ce && t >= k - 1) {
            const int fix = (cs == 0 && t < k + w - 2) ? nd_first_window_fix(sx, lo, t, w, k) : -1;
            flag = fix >= 0 ? fix == 1 : nd_window_emit(sx, lo, t, w, k, len);
        }
        const unsigned m = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) s_wsum[wid] = __popc(m);
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int q = 0; q < NDS_T / 32; ++q) { const uint32_t c = s_wsum[q]; before += q < wid ? c : 0; total += c; }
        if (flag) {
            nd_u128 o;
            o.x = sx[t - lo];
            o.y = ridbits | (nd_u64)(uint32_t)t << 1 | (nd_u64)sz[t - lo];
            dst[base + before + __popc(m & ((1u << lane) - 1))] = o;
        }
        base += total;
        __syncthreads();
    }
    if (tid == 0) cnt[g] = base;
}

// Second generation of the fast path, for the window lengths the pipeline uses (W = 5: ava-ont / ava-pb, W = 10: the -w 10 runs).
// Same closed form, half the instructions:
//   * a thread owns FOUR consecutive positions: the first k-mer is cut out of the packed words, the next three are rolled exactly as
//     the reference rolls them (fw = fw << 2 | c, rv = rv >> 2 | (3 - c) << 2(k-1), ref: minimap2/sketch.c:100-101);
//   * the window test reads the 4 + 2(W-1) hashes around its positions once into registers and counts, without branches, the runs of
//     neighbours that are >= each of them (a position is pushed iff left run + right run >= W-1; invalid positions -- before the first
//     k-mer, beyond the read -- hold 0, which is smaller than every hash and so ends a run);
//   * survivors are compacted in position order with one block-wide prefix sum instead of a ballot per 128 positions.
// which read a chunk belongs to, once per chunk (a 12-step binary search by one thread at the head of every sketch block kept the
// other 127 waiting at its first barrier)
__global__ void k_chunk_owner(const nd_u64 *chunk_off, int n_reads, nd_u64 n_chunks, uint32_t *owner)
{
    nd_u64 g = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_chunks) owner[g] = (uint32_t)nd_find_owner(chunk_off, n_reads, g);
}

template <int W>
__global__ void __launch_bounds__(NDS_T, 8)
k_sketch_fast4(nd_reads_dev R, const nd_u64 *chunk_off, const uint32_t *owner, nd_u64 n_chunks, int k, uint32_t rid0, int use_rid,
               nd_u128 *staging, int S, uint32_t *cnt, uint32_t *slow)
{
    constexpr int PAD = 16;                               // >= W-1 zero entries on both sides of the hashed range
    __shared__ nd_u64 sx[NDS_C + 2 * NDS_MAXW + 2 * PAD];
    __shared__ uint8_t sz[NDS_C + 2 * NDS_MAXW + 2 * PAD];
    __shared__ uint32_t swords[(NDS_C + 2 * NDS_MAXW) / 16 + 8];
    __shared__ uint32_t s_wsum[NDS_T / 32];
    const nd_u64 g = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int r = (int)owner[g], len = (int)R.len[r];
    const int cs = (int)(g - chunk_off[r]) * NDS_C, ce = cs + NDS_C < len ? cs + NDS_C : len;
    if (len < k + W - 1) { if (tid == 0) { slow[r] = 1; cnt[g] = 0; } return; }
    const int lo = cs - (W - 1) > 0 ? cs - (W - 1) : 0, hi = ce + (W - 1) < len ? ce + (W - 1) : len;
    const uint32_t *words = R.words + R.word_off[r];
    const int w0 = (lo >> 4) - 2 > 0 ? (lo >> 4) - 2 : 0, w1 = (hi - 1) >> 4;
    for (int i = tid; i <= w1 - w0; i += NDS_T) swords[i] = words[w0 + i];
    for (int i = tid; i < PAD; i += NDS_T) { sx[i] = 0; sx[PAD + (hi - lo) + i] = 0; }
    __syncthreads();
    const nd_u64 mask = (1ULL << 2 * k) - 1;
    const int rc_shift = 2 * (k - 1);
    int any_sym = 0;
    // the chunk plus its halo is hi - lo <= 512 + 2(W-1) positions: every thread hashes one run of four, the first few threads a run of
    // five (a second round for the halo alone kept 123 of 128 threads waiting at the barrier for as long as the first round took)
    {
        const int extra = hi - lo > 4 * NDS_T ? hi - lo - 4 * NDS_T : 0;      // <= 2(W-1) < NDS_T
        const int base = 4 * tid + (tid < extra ? tid : extra), run = 4 + (tid < extra ? 1 : 0);
        int p = lo + base;
        nd_u64 fw = 0, rv = 0;
        if (base < hi - lo) {
            fw = nd_bases32(swords - w0, p) & mask;
            rv = (~nd_rev2(fw)) >> (64 - 2 * k);
            if (p < k - 1) rv &= ~((1ULL << 2 * (k - 1 - p)) - 1);
        }
#pragma unroll
        for (int j = 0; j < 5; ++j, ++p) {
            if (j >= run || base + j >= hi - lo) break;
            if (j) {
                const nd_u64 c = (swords - w0)[p >> 4] >> (30 - 2 * (p & 15)) & 3;
                fw = (fw << 2 | c) & mask;
                rv = (rv >> 2) | (3ULL ^ c) << rc_shift;
            }
            const int z = fw < rv ? 0 : 1;
            any_sym |= fw == rv;
            sx[PAD + base + j] = p >= k - 1 ? (nd_hash64_masked(z ? rv : fw, mask) << 8 | (nd_u64)k) : 0;
            sz[PAD + base + j] = (uint8_t)z;
        }
    }
    if (__syncthreads_or(any_sym)) { if (tid == 0) { slow[r] = 1; cnt[g] = 0; } return; }
    // ---- window test for positions t0 .. t0+3 from a register window of hashes ----------------------------------------------------
    const int t0 = cs + 4 * tid;
    nd_u64 xw[4 + 2 * (W - 1)];
    {
        const nd_u64 *src = sx + PAD + (t0 - lo) - (W - 1);                  // index PAD + (t - lo) holds position t
#pragma unroll
        for (int i = 0; i < 4 + 2 * (W - 1); ++i) xw[i] = src[i];             // always inside sx[]; entries beyond the hashed range only feed positions >= ce
    }
    unsigned flags = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int t = t0 + j;
        const nd_u64 xt = xw[W - 1 + j];
        int a = 0, b = 0; bool okl = true, okr = true;
#pragma unroll
        for (int d = 1; d < W; ++d) {
            okl = okl && xw[W - 1 + j - d] >= xt; a += okl;
            okr = okr && xw[W - 1 + j + d] >= xt; b += okr;
        }
        bool f = t < ce && t >= k - 1 && a + b >= W - 1;
        if (cs == 0 && t < k + W - 2 && t < ce && t >= k - 1) {             // first-window irregularity (sketch.c:114-121); needs the
            const int fix = nd_first_window_fix(sx + PAD, lo, t, W, k);     // invalid entries to read as "no candidate", not as 0
            if (fix >= 0) f = fix == 1;
        }
        flags |= (unsigned)f << j;
    }
    // ---- compaction in position order ---------------------------------------------------------------------------------------------------
    const uint32_t mine = __popc(flags);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) s_wsum[wid] = incl;
    __syncthreads();
    uint32_t before = incl - mine, total = 0;
#pragma unroll
    for (int q = 0; q < NDS_T / 32; ++q) { const uint32_t c = s_wsum[q]; before += q < wid ? c : 0; total += c; }
    const nd_u64 ridbits = (nd_u64)(use_rid ? rid0 + (uint32_t)r : 0u) << 32;
    nd_u128 *dst = staging + g * (nd_u64)S + before;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (flags >> j & 1) {
            const int t = t0 + j;
            nd_u128 o; o.x = xw[W - 1 + j]; o.y = ridbits | (nd_u64)(uint32_t)t << 1 | (nd_u64)sz[PAD + t - lo];
            *dst++ = o;
        }
    if (tid == 0) cnt[g] = total;
}

// one warp per chunk: staging -> packed minimizer array
__global__ void k_sketch_compact(const nd_u128 *staging, int S, const uint32_t *cnt, const nd_u64 *pos,
                                 nd_u64 n_chunks, nd_u128 *mz)
{
    nd_u64 g = ((nd_u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (g >= n_chunks) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(staging + g * (nd_u64)S);
    uint4 *dst = reinterpret_cast<uint4 *>(mz + pos[g]);
    int n = (int)cnt[g];
    for (int i = lane; i < n; i += 32) dst[i] = src[i];
}

// off_out[r] = pos[chunk_off[r]] (pos has n_chunks+1 entries)
__global__ void k_gather_off(const nd_u64 *pos, const nd_u64 *chunk_off, int n_plus1, nd_u64 *off_out)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_plus1) off_out[r] = pos[chunk_off[r]];
}

// ------------------------------------------------------------------------------------------------
// index build helpers
// ------------------------------------------------------------------------------------------------

// index key = minimizer hash, optionally with the group tag of the read it came from above the 2k hash bits
__global__ void k_split_xy(const nd_u128 *mz, nd_u64 n, nd_u64 *key, nd_u64 *val, const uint32_t *t_tag, int tag_shift, int key_bits)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    nd_u128 m = mz[i];
    nd_u64 k = m.x >> 8;
    if (t_tag) k |= (nd_u64)t_tag[m.y >> 32] << tag_shift;
    key[i] = nd_idx_mix(k, key_bits); val[i] = m.y;
}

// per distinct key i (sorted): d[i] = ideal(key_i) - i, the quantity whose running maximum gives the displacement
__global__ void k_index_ideal(const nd_u64 *ukey, nd_u64 n_distinct, nd_u64 key_mul, uint32_t n_slices, nd_u64 slice_cap, nd_i64 *d)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_distinct) d[i] = (nd_i64)nd_idx_ideal(ukey[i], key_mul, n_slices, slice_cap) - (nd_i64)i;
}

// place the distinct keys of one slice: slot_i = i + running_max_i; value = inline position or (run start, count); histogram of
// run lengths (small ones per block in shared memory: almost every run has length 1).  overflow: a key displaced beyond the pad.
__global__ void __launch_bounds__(256)
k_index_place(const nd_u64 *ukey, const uint32_t *cnt, const nd_u64 *start, const nd_i64 *dmax, nd_u64 n_distinct, const nd_u64 *vals,
              nd_ht_slot *slice, nd_u64 slice_slots, uint32_t *hist, int hist_bins, int *overflow)
{
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_distinct) {
        const nd_u64 slot = (nd_u64)((nd_i64)i + dmax[i]);
        const uint32_t c = cnt[i];
        if (slot >= slice_slots - 1) atomicExch(overflow, 1);     // the last slot stays empty: it ends every scan
        else {
            nd_ht_slot e; e.key = ukey[i];
            const nd_u64 st = start[i];
            e.val = c == 1 ? (ND_HT_INLINE | vals[st]) : (st << 24 | (nd_u64)(c < 0xffffffu ? c : 0xffffffu));
            *reinterpret_cast<ulonglong2 *>(slice + slot) = make_ulonglong2(e.key, e.val);
        }
        if (c < 256) atomicAdd(&sh[c], 1u);
        else atomicAdd(&hist[c < (uint32_t)hist_bins - 1 ? c : (uint32_t)hist_bins - 1], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// which slice each sorted key falls into: bounds[r] = first index whose slice is >= r (n_slices + 1 entries)
__global__ void k_index_slice_bounds(const nd_u64 *key_sorted, nd_u64 n, nd_u64 key_mul, uint32_t n_slices, nd_u64 *bounds)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const uint32_t cur = i < n ? nd_idx_slice(key_sorted[i], key_mul, n_slices) : n_slices;
    const uint32_t prev = i > 0 ? nd_idx_slice(key_sorted[i - 1], key_mul, n_slices) : 0;
    if (i == 0) { for (uint32_t r = 0; r <= cur; ++r) bounds[r] = 0; }
    else for (uint32_t r = prev + 1; r <= cur; ++r) bounds[r] = i;
}

__global__ void k_index_get(nd_idx_view I, nd_u64 key, nd_u64 *out)
{
    nd_u64 v = 0; bool inl = false;
    const uint32_t n = nd_idx_lookup(I, key, &v, &inl);
    out[0] = v; out[1] = n; out[2] = inl ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// seeds: probe + count, then fill.  ref: minimap2/map.c:91-148,216-245
// ------------------------------------------------------------------------------------------------

#define ND_MN_REP 0x40000000u           // m_n of a minimizer dropped by the occurrence filter (t >= max_occ, ref: map.c:105); no anchors

struct nd_seed_par {
    nd_idx_view idx;
    const nd_u64 *t_key; const uint32_t *t_len;
    int mid_occ, no_diag, no_dual;
    int tag_shift;                     // > 0: probe key = hash | tag(read) << tag_shift
};

__device__ __forceinline__ bool nd_skip_seed(const nd_seed_par &P, nd_u64 r, uint32_t q_pos, nd_u64 qkey, uint32_t qlen, int *is_self)
{
    *is_self = 0;
    if (P.no_diag | P.no_dual) {
        uint32_t rid = (uint32_t)(r >> 32);
        nd_u64 tkey = __ldg(P.t_key + rid);
        if (P.no_diag && tkey == qkey && __ldg(P.t_len + rid) == qlen) {
            if ((uint32_t)r >> 1 == q_pos >> 1) return true;
            if ((r & 1) == (q_pos & 1)) *is_self = 1;
        }
        if (P.no_dual && qkey > tkey) return true;
    }
    return false;
}

// one thread per query minimizer
__global__ void __launch_bounds__(256)
k_probe_count(nd_seed_par P, nd_reads_dev Q, const nd_u128 *mz, const nd_u64 *mz_off, nd_u64 n_mz,
              nd_u64 *m_off, uint32_t *m_n, uint32_t *m_keep)
{
    nd_u64 m = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t keep = 0;
    if (m < n_mz) {
        nd_u128 z = mz[m];
        nd_u64 off = 0; uint32_t n = 0; bool inl = false;
        nd_u64 key = z.x >> 8; int r = -1;
        if (P.tag_shift > 0) { r = nd_find_owner(mz_off, Q.n_reads, m); key |= (nd_u64)Q.tag[r] << P.tag_shift; }
        n = nd_idx_lookup(P.idx, key, &off, &inl);
        if (n && n < (uint32_t)P.mid_occ) {
            if (r < 0) r = nd_find_owner(mz_off, Q.n_reads, m);
            nd_u64 qkey = Q.name_key[r]; uint32_t qlen = Q.len[r], q_pos = (uint32_t)z.y;
            int self;
            if (inl) keep = !nd_skip_seed(P, off, q_pos, qkey, qlen, &self);
            else for (uint32_t kk = 0; kk < n; ++kk) keep += !nd_skip_seed(P, __ldg(P.idx.pos + off + kk), q_pos, qkey, qlen, &self);
        } else n = n ? ND_MN_REP : 0;                                            // repetitive (n >= mid_occ) vs absent from the index
        m_off[m] = off; m_n[m] = inl ? (n | 0x80000000u) : n; m_keep[m] = keep; // top bit: m_off is the position itself
    }
}

__global__ void __launch_bounds__(256)
k_fill_anchors(nd_seed_par P, nd_reads_dev Q, const nd_u128 *mz, const nd_u64 *mz_off, nd_u64 n_mz,
               const nd_u64 *m_off, const uint32_t *m_n, const uint32_t *m_keep, const nd_u64 *a_pos, nd_u128 *a)
{
    // one minimizer in five yields anchors: the block first packs the indices of those among its 256 minimizers into shared memory, so
    // that the fill loop below runs in the first few warps with all lanes busy instead of in every warp with a handful
    __shared__ uint16_t s_idx[256];
    __shared__ uint32_t s_wcnt[8];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const nd_u64 m0 = (nd_u64)blockIdx.x * blockDim.x;
    const bool has = m0 + tid < n_mz && m_keep[m0 + tid] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, has);
    if (lane == 0) s_wcnt[wid] = (uint32_t)__popc(bal);
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { const uint32_t cq = s_wcnt[q]; before += q < wid ? cq : 0; total += cq; }
    if (has) s_idx[before + __popc(bal & ((1u << lane) - 1))] = (uint16_t)tid;
    __syncthreads();
    if ((uint32_t)tid >= total) return;
    const nd_u64 m = m0 + s_idx[tid];
    nd_u128 z = mz[m];
    int r = nd_find_owner(mz_off, Q.n_reads, m);
    nd_u64 qkey = Q.name_key[r]; uint32_t qlen = Q.len[r], q_pos = (uint32_t)z.y, q_span = (uint32_t)(z.x & 0xff);
    bool tandem = false;                                   // ref: map.c:113-115
    if (m > mz_off[r] && (mz[m - 1].x >> 8) == (z.x >> 8)) tandem = true;
    if (m + 1 < mz_off[r + 1] && (mz[m + 1].x >> 8) == (z.x >> 8)) tandem = true;
    nd_u64 off = m_off[m]; uint32_t n = m_n[m];
    const bool inl = (n & 0x80000000u) != 0; n &= 0x7fffffffu;
    nd_u128 *dst = a + a_pos[m];
    for (uint32_t kk = 0; kk < n; ++kk) {
        nd_u64 rr = inl ? off : __ldg(P.idx.pos + off + kk);
        int self;
        if (nd_skip_seed(P, rr, q_pos, qkey, qlen, &self)) continue;
        nd_u128 o;
        uint32_t rpos = (uint32_t)rr >> 1;
        if ((rr & 1) == (q_pos & 1)) {
            o.x = (rr & 0xffffffff00000000ULL) | rpos;
            o.y = (nd_u64)q_span << 32 | (q_pos >> 1);
        } else {
            o.x = 1ULL << 63 | (rr & 0xffffffff00000000ULL) | rpos;
            o.y = (nd_u64)q_span << 32 | (uint32_t)((int32_t)qlen - (int32_t)((q_pos >> 1) + 1 - q_span) - 1);
        }
        if (tandem) o.y |= ND_SEED_TANDEM;
        if (self) o.y |= ND_SEED_SELF;
        *dst++ = o;
    }
}

// per read (one warp each): anchor offsets and the float average query span (ref: chain.c:41-42)
__global__ void k_read_stats(const nd_u128 *mz, const nd_u64 *mz_off, const uint32_t *m_keep, const nd_u64 *a_pos,
                             int n_reads, nd_u64 *a_off, float *avg_qspan)
{
    int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r > n_reads) return;
    nd_u64 b = mz_off[r];
    if (lane == 0) a_off[r] = a_pos[b];
    if (r == n_reads) return;
    nd_u64 e = mz_off[r + 1], sum = 0, n = a_pos[e] - a_pos[b];
    for (nd_u64 m = b + lane; m < e; m += 32) sum += (nd_u64)m_keep[m] * (mz[m].x & 0xff);
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) avg_qspan[r] = n ? __fdiv_rn(__ull2float_rn(sum), __ll2float_rn((long long)n)) : 0.f;
}

// ------------------------------------------------------------------------------------------------
// anchor sort, baseline: one thread per read runs the exact flag-sort emulation
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
k_sort_seq(nd_u128 *a, const nd_u64 *a_off, int n_reads, nd_rs_job *jobs)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    nd_u64 b = a_off[r], n = a_off[r + 1] - b;
    if (n < 2) return;
    uint32_t head[256], tail[256];
    nd_rs_sort_128x_seq(a + b, (uint32_t)n, jobs + 2 * (b / 65) + 4 * (nd_u64)r, head, tail);
}

// ------------------------------------------------------------------------------------------------
// segments
// ------------------------------------------------------------------------------------------------
__global__ void k_seg_flags(const nd_u128 *a, nd_u64 n, nd_u64 max_dist_x, uint8_t *flag)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (i == 0 || a[i].x > a[i - 1].x + max_dist_x) ? 1 : 0;
}
__global__ void k_seg_read_starts(const nd_u64 *a_off, int n_reads, uint8_t *flag)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_reads && a_off[r] < a_off[r + 1]) flag[a_off[r]] = 1;
}

// ------------------------------------------------------------------------------------------------
// chain
// ------------------------------------------------------------------------------------------------
struct nd_chain_bufs { int32_t *f, *p, *v, *t, *e; };

// query hash used to randomise the order of a read's hits.  ref: minimap2/map.c:525-527, khash.h:400-409
__device__ __forceinline__ uint32_t nd_wang32(uint32_t key)
{
    key += ~(key << 15); key ^= (key >> 10); key += (key << 3);
    key ^= (key >> 6);   key += ~(key << 11); key ^= (key >> 16);
    return key;
}
__device__ __forceinline__ uint32_t nd_query_hash(uint32_t name_hash, uint32_t qlen, int seed)
{
    uint32_t h = name_hash;
    h ^= nd_wang32(qlen) + nd_wang32((uint32_t)seed);
    return nd_wang32(h);
}

struct nd_seg_view {
    const nd_u128 *a; nd_u64 n_a;
    const uint32_t *seg_start; uint32_t n_seg;
    const nd_u64 *a_off;
};

// segment s -> [b, e); returns false when the segment cannot yield a chain (too few anchors / too
// little total span to reach min_cnt / min_chain_score)
__device__ __forceinline__ bool nd_seg_bounds(const nd_seg_view &S, uint32_t s, const nd_chain_par &par, int max_span, nd_u64 *b, int *n)
{
    nd_u64 lo = S.seg_start[s], hi = s + 1 < S.n_seg ? S.seg_start[s + 1] : S.n_a;
    *b = lo; *n = (int)(hi - lo);
    return *n >= par.min_cnt && (nd_i64)(*n) * max_span >= par.min_sc;
}

// forward DP, baseline: one thread per segment
__global__ void __launch_bounds__(64)
k_chain_fill_seq(nd_seg_view S, nd_reads_dev Q, const float *avg_qspan, nd_chain_par par, int max_span, nd_chain_bufs B)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S.n_seg) return;
    nd_u64 b; int n;
    if (!nd_seg_bounds(S, s, par, max_span, &b, &n)) return;
    int r = nd_find_owner(S.a_off, Q.n_reads, b);
    nd_chain_fill_seq(S.a + b, n, par, avg_qspan[r], B.f + b, B.p + b, B.v + b, B.t + b);
}

// chain ends -> backtrack -> hits: one thread per segment (pointer chasing; latency is hidden by
// the number of segments in flight)
__global__ void __launch_bounds__(64)
k_chain_backtrack(nd_seg_view S, nd_reads_dev Q, nd_chain_par par, int max_span, int seed_h,
                  nd_chain_bufs B, nd_hit *hits, nd_u64 hit_cap, nd_u64 *n_hits)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S.n_seg) return;
    nd_u64 b; int n;
    if (!nd_seg_bounds(S, s, par, max_span, &b, &n)) return;
    int r = nd_find_owner(S.a_off, Q.n_reads, b);
    uint32_t h = nd_query_hash(Q.name_hash[r], Q.len[r], seed_h);
    nd_chain_backtrack_seq(S.a + b, n, par, B.f + b, B.p + b, B.v + b, B.t + b, B.e + b, h, (int)Q.len[r], r,
                           (int)(b - S.a_off[r]),
                           [&](const nd_hit &hit) {
                               nd_u64 slot = atomicAdd(n_hits, 1ULL);
                               if (slot < hit_cap) hits[slot] = hit;
                           });
}

// regime B hand-over: the anchors of every kept chain in chain order (what mm_chain_dp returns in b[],
// ref: minimap2/chain.c:135-141) and the chain's peak score f[j_end] (its backtracking priority, :104).
// One thread per hit; p[] is relative to the start of the segment the chain lives in.
__global__ void k_hit_cnt(const nd_hit *hits, nd_u64 n, uint32_t *cnt)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cnt[i] = (uint32_t)hits[i].cnt;
    if (i == n) cnt[i] = 0;
}
__global__ void k_chain_collect(const nd_hit *hits, nd_u64 n_hits, int qidx_base, const nd_u64 *a_off, const uint32_t *seg_start, uint32_t n_seg,
                                const nd_u128 *a, const int32_t *p, const int32_t *f, const nd_u64 *dst_off, nd_u128 *out, int32_t *fend)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_hits) return;
    const nd_hit h = hits[i];
    const nd_u64 g = a_off[h.qidx - qidx_base] + (nd_u64)h.as;
    uint32_t lo = 0, hi = n_seg;
    while (hi - lo > 1) { uint32_t mid = lo + (hi - lo) / 2; if ((nd_u64)seg_start[mid] <= g) lo = mid; else hi = mid; }
    const nd_u64 b = seg_start[lo];
    fend[i] = f[g];
    int j = (int)(g - b);
    nd_u128 *dst = out + dst_off[i];
    for (int k = h.cnt - 1; k >= 0; --k) { dst[k] = a[b + j]; j = p[b + j]; }
}

// ------------------------------------------------------------------------------------------------
// hit ordering helpers
// ------------------------------------------------------------------------------------------------
__global__ void k_hit_keys(const nd_hit *hits, nd_u64 n, nd_u64 *zkey, uint32_t *idx)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { zkey[i] = hits[i].zkey; idx[i] = (uint32_t)i; }
}
// one composite key when it fits 64 bits: query index above the complement of zkey (zkey < 2^shift because a chain score never
// exceeds the query length): ascending order = by query, then zkey descending
__global__ void k_hit_keys1(const nd_hit *hits, nd_u64 n, int shift, nd_u64 *key, uint32_t *idx)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { key[i] = (nd_u64)(uint32_t)hits[i].qidx << shift | (((1ULL << shift) - 1) - hits[i].zkey); idx[i] = (uint32_t)i; }
}
__global__ void k_hit_qidx(const nd_hit *hits, const uint32_t *idx, nd_u64 n, uint32_t *qidx)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) qidx[i] = (uint32_t)hits[idx[i]].qidx;
}
__global__ void k_hit_gather(const nd_hit *hits, const uint32_t *idx, nd_u64 n, int qidx_base, nd_hit *out)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { nd_hit h = hits[idx[i]]; h.qidx += qidx_base; out[i] = h; }
}

// ------------------------------------------------------------------------------------------------
// PAF-only extras (none of this reaches .ovl): the query minimizers that pass the occurrence filter (`mini_pos`,
// ref: minimap2/map.c:96-117), the length of the query covered by the filtered ones (`rep_len`, :106-111,119) and the raw counts of
// mm_est_err (ref: minimap2/esterr.c:30-64).  The float arithmetic on top (avg_k, logf) is the host writer's.
// ------------------------------------------------------------------------------------------------
__global__ void k_paf_flags(const uint32_t *m_n, nd_u64 n_mz, uint32_t *flag)
{
    nd_u64 m = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < n_mz) flag[m] = (m_n[m] & ND_MN_REP) ? 0u : 1u;
    if (m == n_mz) flag[m] = 0;
}
__global__ void k_paf_minipos(const nd_u128 *mz, const uint32_t *m_n, const nd_u64 *rank, nd_u64 n_mz, nd_u64 *mini_pos)
{
    nd_u64 m = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_mz || (m_n[m] & ND_MN_REP)) return;
    const nd_u128 z = mz[m];
    mini_pos[rank[m]] = (nd_u64)(z.x & 0xff) << 32 | (uint32_t)z.y >> 1;
}
struct nd_paf_query { int32_t rep_len, n_mini; nd_u64 sum_span; };
// one thread per read: the interval union of collect_matches() is sequential over the read's minimizers
__global__ void k_paf_query(const nd_u128 *mz, const nd_u64 *mz_off, const uint32_t *m_n, const nd_u64 *rank, int n_reads, nd_paf_query *out)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    int rep_st = 0, rep_en = 0, rep_len = 0; nd_u64 sum = 0;
    for (nd_u64 m = mz_off[r]; m < mz_off[r + 1]; ++m) {
        const nd_u128 z = mz[m];
        const int q_span = (int)(z.x & 0xff);
        if (m_n[m] & ND_MN_REP) {
            const int en = (int)((uint32_t)z.y >> 1) + 1, st = en - q_span;
            if (st > rep_en) { rep_len += rep_en - rep_st; rep_st = st; rep_en = en; } else rep_en = en;
        } else sum += (nd_u64)q_span;
    }
    rep_len += rep_en - rep_st;
    nd_paf_query o; o.rep_len = rep_len; o.n_mini = (int32_t)(rank[mz_off[r + 1]] - rank[mz_off[r]]); o.sum_span = sum;
    out[r] = o;
}
__device__ __forceinline__ int32_t nd_for_qpos(int32_t qlen, const nd_u128 &a)        // ref: esterr.c:7-14
{
    int32_t x = (int32_t)a.y; const int32_t q_span = (int32_t)(a.y >> 32 & 0xff);
    if (a.x >> 63) x = qlen - 1 - (x + 1 - q_span);
    return x;
}
// one thread per hit (sorted order): st / en / n_match of mm_est_err, literally; out[2i] = n_match, out[2i+1] = en - st + 1 (or -1)
__global__ void k_paf_esterr(const nd_hit *hits, nd_u64 n_hits, int qidx_base, const uint32_t *q_len, const nd_u64 *mz_off, const nd_u64 *rank,
                             const nd_u64 *mini_pos, const nd_u64 *ch_off, const nd_u128 *ch_a, int32_t *out)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_hits) return;
    const nd_hit h = hits[i];
    const int r = h.qidx - qidx_base;
    const int32_t qlen = (int32_t)q_len[r];
    const nd_u64 *mp = mini_pos + rank[mz_off[r]];
    const int32_t n = (int32_t)(rank[mz_off[r + 1]] - rank[mz_off[r]]);
    const nd_u128 *a = ch_a + ch_off[i];
    const int cnt = h.cnt;
    int32_t n_match = 0, span = -1;
    if (n > 0 && cnt > 0) {
        const int32_t x0 = nd_for_qpos(qlen, h.rev ? a[cnt - 1] : a[0]);
        int32_t L = 0, R = n - 1, st = -1;
        while (L <= R) { const int32_t m = (int32_t)(((nd_u64)L + (nd_u64)R) >> 1); const int32_t y = (int32_t)mp[m]; if (y < x0) L = m + 1; else if (y > x0) R = m - 1; else { st = m; break; } }
        if (st >= 0) {
            int32_t en = st, k = 1; n_match = 1;
            for (int32_t j = st + 1; j < n && k < cnt; ++j) {
                const int32_t x = nd_for_qpos(qlen, h.rev ? a[cnt - 1 - k] : a[k]);
                if (x == (int32_t)mp[j]) { ++k; en = j; ++n_match; }
            }
            span = en - st + 1;
        }
    }
    out[2 * i] = n_match; out[2 * i + 1] = span;
}

// ------------------------------------------------------------------------------------------------
// step-1 output stage on the device.  ref: minimap2/map.c:1297-1304 (filter), :804-821 (--dvt test),
// lib/ovl.c:109-150 (delta + big-endian base-128 varints; flag bits 1/2/3 for decreasing names and
// tlen > qlen)
// ------------------------------------------------------------------------------------------------
struct nd_emit_par {
    const uint32_t *q_id, *q_len; const nd_u64 *q_key;     // indexed by hit.qidx - qidx_base
    const uint32_t *t_id, *t_len; const nd_u64 *t_key;     // indexed by hit.rid
    int qidx_base, minlen, dvt, maxhan1, maxhan2;
};

__device__ __forceinline__ int nd_check_dovetail(int rev, uint32_t qs, uint32_t qe, uint32_t qlen, uint32_t ts, uint32_t te, uint32_t tlen,
                                                 int maxhan1, int maxhan2)
{
    if (rev) {
        if (qs <= (uint32_t)maxhan1 && ts <= (uint32_t)maxhan1) return 1;
        else if (qlen - qe <= (uint32_t)maxhan1 && tlen - te <= (uint32_t)maxhan1) return 2;
    } else {
        if (qlen - qe <= (uint32_t)maxhan1 && ts <= (uint32_t)maxhan1) return 4;
        else if (qs <= (uint32_t)maxhan1 && tlen - te <= (uint32_t)maxhan1) return 7;
    }
    if (maxhan2 > 0) {
        if (qs <= (uint32_t)maxhan2 && qe + maxhan2 >= qlen) return 8;
        if (ts <= (uint32_t)maxhan2 && te + maxhan2 >= tlen) return 9;
    }
    return 0;
}

__global__ void k_emit_flags(const nd_hit *hits, nd_u64 n, nd_emit_par P, uint8_t *keep)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const nd_hit h = hits[i];
    const int q = h.qidx - P.qidx_base;
    bool k = P.q_key[q] != P.t_key[h.rid] && h.qe - h.qs >= P.minlen;
    if (k && P.dvt) k = nd_check_dovetail(h.rev, h.qs, h.qe, P.q_len[q], h.rs, h.re, P.t_len[h.rid], P.maxhan1, P.maxhan2) != 0;
    keep[i] = k;
}

__device__ __forceinline__ int nd_varint_len(uint32_t v) { return v < 0x80 ? 1 : v < (1u << 14) ? 2 : v < (1u << 21) ? 3 : v < (1u << 28) ? 4 : 5; }
__device__ __forceinline__ int nd_varint_put(uint8_t *dst, uint32_t v)
{
    const int ng = nd_varint_len(v);
    for (int g = 0; g < ng; ++g) dst[g] = (uint8_t)((v >> (7 * (ng - 1 - g)) & 0x7f) | (g + 1 < ng ? 0x80 : 0));
    return ng;
}

// fields of record j given its predecessor's names (the previous kept record, or the carried state)
__device__ __forceinline__ void nd_emit_fields(const nd_hit *kept, const uint32_t *kidx, nd_u64 j, const nd_emit_par &P, const uint32_t *state, uint32_t f[8])
{
    const nd_hit h = kept[kidx[j]];
    uint32_t pq, pt;
    if (j == 0) { pq = state[0]; pt = state[1]; }
    else { const nd_hit g = kept[kidx[j - 1]]; pq = P.q_id[g.qidx - P.qidx_base]; pt = P.t_id[g.rid]; }
    const uint32_t qn = P.q_id[h.qidx - P.qidx_base], tn = P.t_id[h.rid];
    const uint32_t ql = (uint32_t)(h.qe - h.qs), tl = (uint32_t)(h.re - h.rs);
    uint32_t flags = (uint32_t)h.rev & 0xff;
    if (qn >= pq) f[0] = qn - pq; else { flags |= 2; f[0] = pq - qn; }
    if (tn >= pt) f[4] = tn - pt; else { flags |= 4; f[4] = pt - tn; }
    if (ql >= tl) f[6] = ql - tl; else { flags |= 8; f[6] = tl - ql; }
    f[1] = flags; f[2] = (uint32_t)h.qs; f[3] = ql; f[5] = (uint32_t)h.rs; f[7] = (uint32_t)h.mlen;
}

__global__ void k_emit_len(const nd_hit *hits, const uint32_t *kidx, nd_u64 n_kept, nd_emit_par P, const uint32_t *state, uint32_t *len, nd_u64 *sum_qspan)
{
    nd_u64 j = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    nd_u64 span = 0;
    if (j < n_kept) {
        uint32_t f[8]; nd_emit_fields(hits, kidx, j, P, state, f);
        int nb = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) nb += nd_varint_len(f[i]);
        len[j] = (uint32_t)nb;
        span = f[3];
    }
    for (int o = 16; o; o >>= 1) span += __shfl_xor_sync(0xffffffffu, span, o);
    if ((threadIdx.x & 31) == 0 && span) atomicAdd(sum_qspan, span);
}

__global__ void k_emit_write(const nd_hit *hits, const uint32_t *kidx, nd_u64 n_kept, nd_emit_par P, const uint32_t *state, const nd_u64 *off,
                             uint8_t *out, uint32_t *new_state)
{
    nd_u64 j = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_kept) return;
    uint32_t f[8]; nd_emit_fields(hits, kidx, j, P, state, f);
    uint8_t *dst = out + off[j];
#pragma unroll
    for (int i = 0; i < 8; ++i) dst += nd_varint_put(dst, f[i]);
    if (j == n_kept - 1) { const nd_hit h = hits[kidx[j]]; new_state[0] = P.q_id[h.qidx - P.qidx_base]; new_state[1] = P.t_id[h.rid]; }
}
