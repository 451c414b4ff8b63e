// nd_core.cuh -- per-thread ("one lane") building blocks of the overlap hot path.
//
// Everything here is __host__ __device__ so the exact same code that the sm_100a kernels execute
// can be unit-tested on a CPU-only box by tests/hd_harness (test infrastructure; the product only
// ever launches these through kernels in nd_kernels.cu).
//
// Citations "ref:" are relative to the reference tree (/root/reference/).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ND_HD __host__ __device__ __forceinline__
#define ND_HD_NOINLINE __host__ __device__
#else
#define ND_HD inline
#define ND_HD_NOINLINE
#endif

typedef unsigned long long nd_u64;
typedef long long nd_i64;

struct nd_u128 { nd_u64 x, y; };                 // ref: minimap2/minimap.h:53 (mm128_t)

#define ND_U64_MAX 0xffffffffffffffffULL
#define ND_SEED_TANDEM (1ULL << 42)              // ref: minimap2/mmpriv.h:20
#define ND_SEED_SELF   (1ULL << 43)              // ref: minimap2/mmpriv.h:21
#define ND_RS_SMALL 64                           // ref: minimap2/ksort.h:95 (RS_MIN_SIZE)

// ------------------------------------------------------------------------------------------------
// hashes
// ------------------------------------------------------------------------------------------------

// ref: minimap2/sketch.c:28-38 -- invertible hash on the low 2k bits
ND_HD nd_u64 nd_hash64_masked(nd_u64 key, nd_u64 mask)
{
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

// ref: minimap2/hit.c:40-50 -- same mixer without the mask, used to randomise hit order
ND_HD nd_u64 nd_mix64(nd_u64 key)
{
    key = (~key + (key << 21));
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8));
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4));
    key = key ^ key >> 28;
    key = (key + (key << 31));
    return key;
}

// ------------------------------------------------------------------------------------------------
// ordered index table: where a key belongs (layout and rationale: nd_kernels.cuh).  key * key_mul is a 64-bit fixed-point
// fraction in [0, 1), monotone in the key; its integer part after scaling by n_slices is the slice, the rest the ideal slot.
// ------------------------------------------------------------------------------------------------
// Minimizer keys are NOT uniform (a minimizer is the minimum hash of its window, so keys crowd towards 0): the table is ordered
// by a second bijective mix of the key on key_bits bits (odd multiplier mod 2^key_bits), which spreads any key set evenly.
ND_HD nd_u64 nd_idx_mix(nd_u64 key, int key_bits) { return (key * 0x9E3779B97F4A7C15ULL) & ((1ULL << key_bits) - 1); }

ND_HD nd_u64 nd_mulhi64(nd_u64 a, nd_u64 b)
{
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (nd_u64)(((unsigned __int128)a * b) >> 64);
#endif
}
ND_HD uint32_t nd_idx_slice(nd_u64 key, nd_u64 key_mul, uint32_t n_slices) { return (uint32_t)nd_mulhi64(key * key_mul, (nd_u64)n_slices); }
ND_HD nd_u64 nd_idx_ideal(nd_u64 key, nd_u64 key_mul, uint32_t n_slices, nd_u64 slice_cap)   // slice-local ideal slot
{
    const nd_u64 t = key * key_mul;
    return nd_mulhi64(t * (nd_u64)n_slices, slice_cap);        // t * n_slices (mod 2^64) = position inside the slice as a fraction
}

// ------------------------------------------------------------------------------------------------
// 2-bit packed reads: 16 bases per u32, first base in the top two bits (ref: lib/bseq.c:114-139);
// optional ambiguity mask, one bit per base (bit (i&31) of nmask[i>>5]), for FASTA input with N.
// ------------------------------------------------------------------------------------------------
struct nd_seqview {
    const uint32_t *words;
    const uint32_t *nmask;   // may be null
    int len;
};

ND_HD int nd_base(const nd_seqview &s, int i)
{
    if (s.nmask && (s.nmask[i >> 5] >> (i & 31) & 1)) return 4;
    return (int)(s.words[i >> 4] >> (30 - 2 * (i & 15)) & 3);
}

// ------------------------------------------------------------------------------------------------
// minimizer sketch, one chunk of one read per call
// ------------------------------------------------------------------------------------------------

// Is position p the start of an iteration of the reference loop (minimap2/sketch.c:89)?  HPC folds
// a homopolymer run into one iteration; ambiguous bases are never folded.
ND_HD bool nd_iter_start(const nd_seqview &s, int p, int hpc)
{
    if (p == 0 || !hpc) return true;
    int c = nd_base(s, p), d = nd_base(s, p - 1);
    return c == 4 || d == 4 || c != d;
}

// Emit the minimizers pushed by loop iterations of the reference's sliding-window state machine
// (ref: minimap2/sketch.c:77-143) whose start index lies in [cs, ce), plus the final flush when
// ce >= len.  Output is in the reference's push order, so concatenating the chunks of a read in
// order reproduces mm_sketch() exactly.  Returns the number of pushes; entries beyond `cap` are
// not written (the caller treats n > cap as a fatal overflow).  RW = ring capacity (> w).
//
// Exactness of a mid-read start: the state at cs depends only on (a) the last k non-ambiguous
// bases / k runs (k-mers, HPC span queue) and (b) the last w slots plus whether l has reached w+k.
// We therefore (1) consume k non-ambiguous "units" with the window logic disabled, then (2) run the
// full state machine silently until >= w+k slots have been written, and only then reach cs.  If
// the look-back was too short for that (ambiguous bases, symmetric k-mers, long homopolymers) it
// is doubled and the warm-up restarted; a warm-up that starts at position 0 is exact by
// construction.
// Long k-mers (k > 28; the ava-hifi preset uses k = 51): ref: minimap2/sketch.c:215-356.  The k-mer is 2k bits in four words (word 0 =
// newest bases), the strands are compared as 256-bit integers, the hash is hash64 of the top (partial) word under its mask plus the
// unmasked hash of every lower word that is not zero.  Two quirks of the reference are part of its output and kept:
//   * for k a multiple of 32 the top-word mask is computed as (1 << 64) - 1, which x86 evaluates to 0;
//   * the homopolymer run queue has 32 entries whatever k is, so for k > 31 a run is overwritten before it is popped and `span` is
//     A + (sum of the last k mod 32 runs) with A = the sum of the first k - (k mod 32) runs after the last reset (read start or
//     ambiguous base) -- history a chunk does not see.  A chunk that starts mid-read therefore computes A for the true reset point
//     and for its own warm-up start (two short forward scans) and shifts `span` by the difference once its queue is primed.
struct nd_k256 { nd_u64 u[4]; };
ND_HD nd_u64 nd_hash64_nomask(nd_u64 key) { return nd_mix64(key); }
ND_HD void nd_k256_push(nd_k256 &F, nd_k256 &R, int c, int k_idx, nd_u64 mask, int rc_shift)
{
    F.u[3] = F.u[3] << 2 | F.u[2] >> 62; F.u[2] = F.u[2] << 2 | F.u[1] >> 62; F.u[1] = F.u[1] << 2 | F.u[0] >> 62; F.u[0] = F.u[0] << 2 | (nd_u64)c;
    R.u[0] = R.u[0] >> 2 | R.u[1] << 62; R.u[1] = R.u[1] >> 2 | R.u[2] << 62; R.u[2] = R.u[2] >> 2 | R.u[3] << 62; R.u[3] >>= 2;
    if (k_idx == 0) { F.u[0] &= mask; R.u[0] |= (nd_u64)(3 ^ c) << rc_shift; }
    else if (k_idx == 1) { F.u[1] &= mask; R.u[1] |= (nd_u64)(3 ^ c) << rc_shift; }
    else if (k_idx == 2) { F.u[2] &= mask; R.u[2] |= (nd_u64)(3 ^ c) << rc_shift; }
    else { F.u[3] &= mask; R.u[3] |= (nd_u64)(3 ^ c) << rc_shift; }
}
ND_HD int nd_k256_cmp(const nd_k256 &a, const nd_k256 &b)
{
    if (a.u[3] != b.u[3]) return a.u[3] < b.u[3] ? -1 : 1;
    if (a.u[2] != b.u[2]) return a.u[2] < b.u[2] ? -1 : 1;
    if (a.u[1] != b.u[1]) return a.u[1] < b.u[1] ? -1 : 1;
    if (a.u[0] != b.u[0]) return a.u[0] < b.u[0] ? -1 : 1;
    return 0;
}
ND_HD nd_u64 nd_k256_hash(const nd_k256 &K, int k_idx, nd_u64 mask)
{
    const nd_u64 top = k_idx == 0 ? K.u[0] : k_idx == 1 ? K.u[1] : k_idx == 2 ? K.u[2] : K.u[3];
    nd_u64 h = nd_hash64_masked(top, mask);
    if (k_idx > 2 && K.u[2]) h += nd_hash64_nomask(K.u[2]);
    if (k_idx > 1 && K.u[1]) h += nd_hash64_nomask(K.u[1]);
    if (k_idx > 0 && K.u[0]) h += nd_hash64_nomask(K.u[0]);
    return h;
}
// sum of the first `cnt` homopolymer runs starting at iteration boundary p (stops early at an ambiguous base or the read end)
ND_HD int nd_first_runs(const nd_seqview &s, int p, int cnt)
{
    int sum = 0;
    while (cnt > 0 && p < s.len) {
        const int c = nd_base(s, p);
        if (c == 4) break;
        int run = 1;
        while (p + run < s.len && nd_base(s, p + run) == c) ++run;
        sum += run; p += run; --cnt;
    }
    return sum;
}

template <int RW, bool LONGK>
ND_HD_NOINLINE int nd_sketch_chunk(const nd_seqview s, int w, int k, int hpc, uint32_t rid,
                                   int cs, int ce, nd_u128 *out, int cap)
{
    const int k_idx = (k - 1) / 32, top_bits = (((k - 1) & 31) + 1) << 1;
    const nd_u64 mask = LONGK ? (1ULL << (top_bits & 63)) - 1 : (1ULL << 2 * k) - 1;
    const int rc_shift = LONGK ? ((k - 1) & 31) << 1 : 2 * (k - 1);
    const nd_u128 none = { ND_U64_MAX, ND_U64_MAX };
    nd_u128 ring[RW], cur;
    nd_u64 fw, rv;
    nd_k256 F, R;
    int runq[32], rq_head, rq_n;               // HPC: run lengths, a 32-entry ring as the reference's tiny_queue_t (ref: sketch.c:40-58)
    int slot, cur_slot, l, span;
    int lookback = 2 * k + w + 8;
    int start, i, j, n;

    for (;;) {
        int units = 0, slots = 0, adj = 0;
        bool primed, verified, retry = false;
        start = cs;
        if (cs > 0) {                          // back up to an iteration boundary `lookback` iterations earlier
            int need = lookback;
            while (start > 0 && need > 0) { --start; if (nd_iter_start(s, start, hpc)) --need; }
            while (start > 0 && !nd_iter_start(s, start, hpc)) --start;
        }
        primed = verified = (start == 0);
        if (LONGK && hpc && k > 31 && start > 0) {   // span offset of the wrapped run queue (header comment)
            int rp = start;                    // true reset point: after the last ambiguous base before `start`, or the read start
            if (s.nmask) { while (rp > 0 && nd_base(s, rp - 1) != 4) --rp; } else rp = 0;
            const int first = k - (k & 31);
            adj = nd_first_runs(s, rp, first) - nd_first_runs(s, start, first);
        }
        fw = rv = 0; cur = none; rq_head = rq_n = 0;
        F.u[0] = F.u[1] = F.u[2] = F.u[3] = 0; R = F;
        slot = cur_slot = l = span = 0; n = 0;
        for (j = 0; j < w; ++j) ring[j] = none;

        for (i = start; i < s.len; ++i) {
            const bool live = i >= cs;         // pushes of this iteration belong to this chunk
            if (live) {
                if (i >= ce) break;
                if (!verified) {
                    if (slots < w + k) { retry = true; break; }
                    verified = true;
                }
            }
            int c = nd_base(s, i);
            nd_u128 cand = none;
            if (c < 4) {
                int z;
                if (hpc) {
                    int run = 1;
                    if (i + 1 < s.len && nd_base(s, i + 1) == c) {
                        for (run = 2; i + run < s.len; ++run)
                            if (nd_base(s, i + run) != c) break;
                        i += run - 1;
                    }
                    runq[(rq_head + rq_n++) & 31] = run;
                    span += run;
                    if (rq_n > k) { span -= runq[rq_head]; rq_head = (rq_head + 1) & 31; --rq_n; }
                } else span = l + 1 < k ? l + 1 : k;
                bool sym;
                if (!LONGK) {
                    fw = (fw << 2 | (nd_u64)c) & mask;
                    rv = (rv >> 2) | (nd_u64)(3 ^ c) << rc_shift;
                    sym = fw == rv; z = fw < rv ? 0 : 1;
                } else {
                    nd_k256_push(F, R, c, k_idx, mask, rc_shift);
                    const int cmp = nd_k256_cmp(F, R);
                    sym = cmp == 0; z = cmp < 0 ? 0 : 1;
                }
                if (!primed && units < k) {                        // phase 1: history only
                    if (++units == k) { span += adj; adj = 0; }    // queue primed: from here on `span` is the reference's
                    continue;
                }
                if (sym) continue;             // strand-symmetric k-mer: no slot
                ++l;
                if (l >= k && span < 256) {
                    const nd_u64 h = LONGK ? nd_k256_hash(z ? R : F, k_idx, mask) : nd_hash64_masked(z ? rv : fw, mask);
                    cand.x = h << 8 | (nd_u64)span;
                    cand.y = (nd_u64)rid << 32 | (nd_u64)(uint32_t)i << 1 | (nd_u64)z;
                }
            } else {
                l = 0; rq_n = rq_head = 0; span = 0; adj = 0;     // a real reset: the emulated queue is exact from here on
                if (!primed && units < k) continue;
            }
            ++slots;
            ring[slot] = cand;
#define ND_PUSH(v) do { if (live) { if (n < cap) out[n] = (v); ++n; } } while (0)
            if (l == w + k - 1 && cur.x != ND_U64_MAX) {
                for (j = slot + 1; j < w; ++j) if (cur.x == ring[j].x && ring[j].y != cur.y) ND_PUSH(ring[j]);
                for (j = 0; j < slot; ++j)     if (cur.x == ring[j].x && ring[j].y != cur.y) ND_PUSH(ring[j]);
            }
            if (cand.x <= cur.x) {
                if (l >= w + k && cur.x != ND_U64_MAX) ND_PUSH(cur);
                cur = cand; cur_slot = slot;
            } else if (slot == cur_slot) {
                if (l >= w + k - 1 && cur.x != ND_U64_MAX) ND_PUSH(cur);
                cur.x = ND_U64_MAX;
                for (j = slot + 1; j < w; ++j) if (cur.x >= ring[j].x) { cur = ring[j]; cur_slot = j; }
                for (j = 0; j <= slot; ++j)    if (cur.x >= ring[j].x) { cur = ring[j]; cur_slot = j; }
                if (l >= w + k - 1 && cur.x != ND_U64_MAX) {
                    for (j = slot + 1; j < w; ++j) if (cur.x == ring[j].x && cur.y != ring[j].y) ND_PUSH(ring[j]);
                    for (j = 0; j <= slot; ++j)    if (cur.x == ring[j].x && cur.y != ring[j].y) ND_PUSH(ring[j]);
                }
            }
#undef ND_PUSH
            if (++slot == w) slot = 0;
        }
        if (!retry && ce >= s.len && !verified) {   // read ended inside the warm-up: flush needs an exact state too
            if (slots < w + k) retry = true;
        }
        if (retry) { lookback *= 2; continue; }
        break;
    }
    if (ce >= s.len && cur.x != ND_U64_MAX) { if (n < cap) out[n] = cur; ++n; }
    return n;
}

// ------------------------------------------------------------------------------------------------
// closed form of the sketch for the common case (no ambiguous bases, no HPC, no strand-symmetric
// k-mer anywhere in the read, at least one full window).  Then every position is a slot, l = p+1, a
// k-mer is valid iff p >= k-1, and replaying minimap2/sketch.c:89-142 shows that position t is pushed
// iff some full window [s, s+w-1] inside [k-1, len-1] contains t with x[t] equal to the window
// minimum (ties are all pushed; the right-most one is the running "min"), each exactly once and in
// position order.  Equivalently: with a/b the runs of neighbours (left/right, inside the valid range,
// at most w-1) whose x is >= x[t], t is pushed iff a + b >= w-1.
// One start-up irregularity of the reference is kept (nd_first_window_fix): at the step that completes
// the first window (position P = k+w-2, l == w+k-1) the ties of the running minimum are pushed
// unconditionally (sketch.c:114-119) while the running minimum t* itself can only be pushed later,
// i.e. iff x[P] > x[t*] (a displacement at that step needs l >= w+k, sketch.c:121).
// Reads that violate a precondition are routed to the state machine above.
// ------------------------------------------------------------------------------------------------

// 32 bases ending at p (base p in the two lowest bits, zeros before the read start), from the packed words
ND_HD nd_u64 nd_bases32(const uint32_t *words, int p)
{
    const int j = p >> 4, s = 2 * (15 - (p & 15));
    const nd_u64 w2 = words[j], w1 = j >= 1 ? words[j - 1] : 0, w0 = j >= 2 ? words[j - 2] : 0;
    const nd_u64 hi = w0 << 32 | w1;
    return (hi << (32 - s)) | (w2 >> s);
}

ND_HD nd_u64 nd_rev2(nd_u64 v)     // reverse the order of the 32 two-bit groups
{
#if defined(__CUDA_ARCH__)
    v = __brevll(v);
#else
    v = ((v >> 1) & 0x5555555555555555ULL) | ((v & 0x5555555555555555ULL) << 1);
    v = ((v >> 2) & 0x3333333333333333ULL) | ((v & 0x3333333333333333ULL) << 2);
    v = ((v >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((v & 0x0F0F0F0F0F0F0F0FULL) << 4);
    v = ((v >> 8) & 0x00FF00FF00FF00FFULL) | ((v & 0x00FF00FF00FF00FFULL) << 8);
    v = ((v >> 16) & 0x0000FFFF0000FFFFULL) | ((v & 0x0000FFFF0000FFFFULL) << 16);
    v = (v >> 32) | (v << 32);
#endif
    return ((v & 0xAAAAAAAAAAAAAAAAULL) >> 1) | ((v & 0x5555555555555555ULL) << 1);
}

// x (hash<<8|k, or ND_U64_MAX when p < k-1) and strand of position p; *sym set when forward == reverse
ND_HD nd_u64 nd_kmer_x(const uint32_t *words, int p, int k, int *z, bool *sym)
{
    const nd_u64 mask = (1ULL << 2 * k) - 1;
    const nd_u64 fw = nd_bases32(words, p) & mask;
    nd_u64 rv = (~nd_rev2(fw)) >> (64 - 2 * k);
    if (p < k - 1) rv &= ~((1ULL << 2 * (k - 1 - p)) - 1);    // bases before the read start contribute 0 (sketch.c:79)
    *sym = fw == rv;
    *z = fw < rv ? 0 : 1;
    if (p < k - 1) return ND_U64_MAX;
    return nd_hash64_masked(*z ? rv : fw, mask) << 8 | (nd_u64)k;
}

// First-window irregularity: for t in [k-1, P-1], P = k+w-2.  xs[i] = x of position lo+i, lo <= k-1.
// Returns -1 when the general rule applies to t, else 0/1 = pushed.
ND_HD int nd_first_window_fix(const nd_u64 *xs, int lo, int t, int w, int k)
{
    const int P = k + w - 2;
    if (t >= P) return -1;
    nd_u64 m1 = ND_U64_MAX; int tstar = -1;
    for (int j = k - 1; j < P; ++j) if (xs[j - lo] <= m1) { m1 = xs[j - lo]; tstar = j; }
    if (xs[t - lo] != m1) return -1;
    if (t != tstar) return 1;
    return xs[P - lo] > m1 ? 1 : 0;
}

// xs[i] = x of position lo+i.  Is position t pushed?
ND_HD bool nd_window_emit(const nd_u64 *xs, int lo, int t, int w, int k, int len)
{
    const nd_u64 xt = xs[t - lo];
    int a = 0, b = 0;
    while (a < w - 1 && t - 1 - a >= k - 1 && xs[t - 1 - a - lo] >= xt) ++a;
    const int need = w - 1 - a;
    while (b < need && t + 1 + b <= len - 1 && xs[t + 1 + b - lo] >= xt) ++b;
    return a + b >= w - 1;
}

// ------------------------------------------------------------------------------------------------
// exact emulation of the reference's in-place MSD radix sort (tie order is observable)
// ref: minimap2/ksort.h:98-151 (rs_insertsort / rs_sort / radix_sort), minimap2/misc.c:155-156
// ------------------------------------------------------------------------------------------------

ND_HD void nd_ins_sort_128x(nd_u128 *beg, nd_u128 *end)
{
    for (nd_u128 *i = beg + 1; i < end; ++i)
        if (i->x < (i - 1)->x) {
            nd_u128 tmp = *i, *j;
            for (j = i; j > beg && tmp.x < (j - 1)->x; --j) *j = *(j - 1);
            *j = tmp;
        }
}

struct nd_rs_job { uint32_t beg, n; int shift; int buf; };   // buf: which of the two anchor buffers holds the bucket (warp sort)

// One pass of the flag sort over a[beg, beg+n) on byte `shift`; appends child ranges that need
// another pass to `next` (at most n/65 entries), insertion-sorts the small ones in place.
ND_HD_NOINLINE uint32_t nd_rs_pass(nd_u128 *a, nd_rs_job job, nd_rs_job *next, uint32_t n_next,
                                   uint32_t *head, uint32_t *tail)
{
    const int s = job.shift;
    nd_u128 *base = a + job.beg;
    uint32_t i, d, acc = 0, nonempty = 0;
    for (d = 0; d < 256; ++d) head[d] = 0;
    for (i = 0; i < job.n; ++i) ++head[base[i].x >> s & 255];
    for (d = 0; d < 256; ++d) { uint32_t c = head[d]; nonempty += c != 0; head[d] = acc; acc += c; tail[d] = acc; }
    if (nonempty > 1) {                               // a single non-empty bucket is a no-op pass
        for (d = 0; d < 256;) {
            if (head[d] == tail[d]) { ++d; continue; }
            uint32_t t = (uint32_t)(base[head[d]].x >> s & 255);
            if (t == d) { ++head[d]; continue; }
            nd_u128 carry = base[head[d]];
            do {
                nd_u128 evicted = base[head[t]];
                base[head[t]++] = carry;
                carry = evicted;
                t = (uint32_t)(carry.x >> s & 255);
            } while (t != d);
            base[head[d]++] = carry;
        }
    }
    if (s) {
        int ns = s > 8 ? s - 8 : 0;
        for (d = 0; d < 256; ++d) {
            uint32_t b = d ? tail[d - 1] : 0, e = tail[d];
            if (e - b > ND_RS_SMALL) { next[n_next].beg = job.beg + b; next[n_next].n = e - b; next[n_next].shift = ns; ++n_next; }
            else if (e - b > 1) nd_ins_sort_128x(base + b, base + e);
        }
    }
    return n_next;
}

// Whole sort by one thread.  `jobs` is scratch for 2*(n/65+1) entries.
ND_HD_NOINLINE void nd_rs_sort_128x_seq(nd_u128 *a, uint32_t n, nd_rs_job *jobs, uint32_t *head, uint32_t *tail)
{
    if (n <= ND_RS_SMALL) { nd_ins_sort_128x(a, a + n); return; }
    uint32_t half = n / (ND_RS_SMALL + 1) + 1;
    nd_rs_job *cur = jobs, *nxt = jobs + half;
    uint32_t n_cur = 1, n_nxt;
    cur[0].beg = 0; cur[0].n = n; cur[0].shift = 56;
    while (n_cur) {
        n_nxt = 0;
        for (uint32_t q = 0; q < n_cur; ++q) n_nxt = nd_rs_pass(a, cur[q], nxt, n_nxt, head, tail);
        nd_rs_job *t = cur; cur = nxt; nxt = t; n_cur = n_nxt;
    }
}

// ------------------------------------------------------------------------------------------------
// chaining: forward DP over one independent segment (all anchors share strand+target and no gap
// between consecutive anchors exceeds max_dist_x, so no anchor outside the segment is ever inside
// the reference's [st, i) window).  ref: minimap2/chain.c:45-86
// ------------------------------------------------------------------------------------------------

struct nd_chain_par {
    int max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc;
};

ND_HD int nd_ilog2_u32(uint32_t v)   // ref: minimap2/chain.c:8-20 ; v > 0
{
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    int r = 0; while (v >>= 1) ++r; return r;
#endif
}

// gap cost term of chain.c:72 -- two separately rounded double products, then truncation
ND_HD int nd_gap_cost(int dd, float avg_qspan)
{
#if defined(__CUDA_ARCH__)
    double lin = __dmul_rn(__dmul_rn((double)dd, .01), (double)avg_qspan);
#else
    volatile double t = (double)dd * .01;
    double lin = t * (double)avg_qspan;
#endif
    int lg = dd ? nd_ilog2_u32((uint32_t)dd) : 0;
    return (int)lin + (lg >> 1);
}

// a: the segment's anchors; f,p,v,t: per-anchor scratch (same indexing).  Indices are segment
// relative; p = -1 for "no predecessor".
ND_HD_NOINLINE void nd_chain_fill_seq(const nd_u128 *a, int n, const nd_chain_par par, float avg_qspan,
                                      int32_t *f, int32_t *p, int32_t *v, int32_t *t)
{
    int st = 0;
    for (int i = 0; i < n; ++i) t[i] = 0;
    for (int i = 0; i < n; ++i) {
        nd_u64 ri = a[i].x;
        int32_t qi = (int32_t)a[i].y, q_span = (int32_t)(a[i].y >> 32 & 0xff);
        int32_t max_f = q_span, n_skip = 0, max_j = -1;
        while (st < i && ri > a[st].x + (nd_u64)par.max_dist_x) ++st;
        if (i - st > par.max_iter) st = i - par.max_iter;
        for (int j = i - 1; j >= st; --j) {
            nd_i64 dr = (nd_i64)(ri - a[j].x);
            int32_t dq = qi - (int32_t)a[j].y, dd, sc;
            if (dr == 0 || dq <= 0) continue;
            if (dq > par.max_dist_y || dq > par.max_dist_x) continue;
            dd = (int32_t)(dr > dq ? dr - dq : dq - dr);
            if (dd > par.bw) continue;
            sc = dq < dr ? dq : (int32_t)dr;
            if (sc > q_span) sc = q_span;
            sc -= nd_gap_cost(dd, avg_qspan);
            sc += f[j];
            if (sc > max_f) {
                max_f = sc; max_j = j;
                if (n_skip > 0) --n_skip;
            } else if (t[j] == i) {
                if (++n_skip > par.max_skip) break;
            }
            if (p[j] >= 0) t[p[j]] = i;
        }
        f[i] = max_f; p[i] = max_j;
        v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f;
    }
}

// ------------------------------------------------------------------------------------------------
// chain ends -> peaks -> backtrack -> hits, one segment.  ref: minimap2/chain.c:88-130 and
// minimap2/hit.c:8-38,52-88 (coordinates, fuzzy lengths, randomised sort key).
// ------------------------------------------------------------------------------------------------

struct nd_hit {                // what leaves the hot path per chain (subset of mm_reg1_t, minimap.h:84-99)
    nd_u64 zkey;               // (score<<32 | cnt) ^ h : the reference orders a read's hits by this, descending
    int32_t qidx, rid;
    int32_t qs, qe, rs, re;
    int32_t mlen, blen;
    int32_t score, cnt;
    int32_t rev;
    int32_t as;                // index (read-relative) of the chain's last anchor, for later stages
};

ND_HD bool nd_end_less(const int32_t *f, int32_t ja, int32_t jb)   // order of (f[j]<<32 | j)
{
    return f[ja] < f[jb] || (f[ja] == f[jb] && ja < jb);
}

// heap sort of ends[0..m) ascending by (f[j], j)
ND_HD void nd_sort_ends(int32_t *e, int m, const int32_t *f)
{
    for (int start = m / 2 - 1; start >= 0; --start) {
        int root = start;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= m) break;
            if (child + 1 < m && nd_end_less(f, e[child], e[child + 1])) ++child;
            if (nd_end_less(f, e[root], e[child])) { int32_t tmp = e[root]; e[root] = e[child]; e[child] = tmp; root = child; }
            else break;
        }
    }
    for (int end = m - 1; end > 0; --end) {
        int32_t tmp = e[0]; e[0] = e[end]; e[end] = tmp;
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && nd_end_less(f, e[child], e[child + 1])) ++child;
            if (nd_end_less(f, e[root], e[child])) { int32_t t2 = e[root]; e[root] = e[child]; e[child] = t2; root = child; }
            else break;
        }
    }
}

// Returns the number of hits written through `emit(hit)`.  t and e are scratch (n entries each).
// seg_off = index of the segment's first anchor within the read's anchor array.
template <class Emit>
ND_HD_NOINLINE int nd_chain_backtrack_seq(const nd_u128 *a, int n, const nd_chain_par par,
                                          const int32_t *f, const int32_t *p, const int32_t *v,
                                          int32_t *t, int32_t *e, uint32_t qhash, int qlen, int qidx, int seg_off,
                                          Emit emit)
{
    int m = 0, n_hits = 0;
    for (int i = 0; i < n; ++i) t[i] = 0;
    for (int i = 0; i < n; ++i) if (p[i] >= 0) t[p[i]] = 1;
    for (int i = 0; i < n; ++i)
        if (t[i] == 0 && v[i] >= par.min_sc) {
            int j = i;
            while (j >= 0 && f[j] < v[j]) j = p[j];
            if (j < 0) j = i;
            e[m++] = j;
        }
    if (m == 0) return 0;
    nd_sort_ends(e, m, f);
    for (int i = 0; i < n; ++i) t[i] = 0;
    for (int q = m - 1; q >= 0; --q) {               // highest (score, index) first
        int j = e[q], last = j, first = j, cnt = 0;
        int32_t sc = f[j], mlen = 0, blen = 0;
        int prev = -1;
        do {
            if (prev >= 0) {                          // pair (j, prev): prev is the later anchor
                int span = (int)(a[prev].y >> 32 & 0xff);
                int tl = (int32_t)a[prev].x - (int32_t)a[j].x;
                int ql = (int32_t)a[prev].y - (int32_t)a[j].y;
                blen += tl > ql ? tl : ql;
                mlen += tl > span && ql > span ? span : tl < ql ? tl : ql;
            }
            ++cnt; t[j] = 1; first = j; prev = j; j = p[j];
        } while (j >= 0 && t[j] == 0);
        bool keep = false;
        if (j < 0) keep = cnt >= par.min_cnt;
        else if (sc - f[j] >= par.min_sc) { keep = cnt >= par.min_cnt; sc -= f[j]; }
        if (!keep) continue;
        {
            nd_hit h;
            int32_t span0 = (int32_t)(a[first].y >> 32 & 0xff);
            uint32_t hh = (uint32_t)nd_mix64((nd_mix64(a[first].x) + nd_mix64(a[first].y)) ^ (nd_u64)qhash);
            h.zkey = ((nd_u64)(uint32_t)sc << 32 | (nd_u64)(uint32_t)cnt) ^ (nd_u64)hh;
            h.qidx = qidx;
            h.rev = (int32_t)(a[first].x >> 63);
            h.rid = (int32_t)(a[first].x << 1 >> 33);
            h.rs = (int32_t)a[first].x + 1 > span0 ? (int32_t)a[first].x + 1 - span0 : 0;
            h.re = (int32_t)a[last].x + 1;
            if (!h.rev) { h.qs = (int32_t)a[first].y + 1 - span0; h.qe = (int32_t)a[last].y + 1; }
            else { h.qs = qlen - ((int32_t)a[last].y + 1); h.qe = qlen - ((int32_t)a[first].y + 1 - span0); }
            h.mlen = mlen + span0; h.blen = blen + span0;
            h.score = sc; h.cnt = cnt; h.as = seg_off + last;
            emit(h);
            ++n_hits;
        }
    }
    return n_hits;
}
