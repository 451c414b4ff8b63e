// nd_ovlsort_gpu.cuh -- kernels of the GPU `ovl_sort` (ndg_ovl_sort): the records of the 8-field .ovl streams are decoded, mirrored,
// sorted and filtered per query read on the device and re-encoded; only the input bytes go up and the output bytes come down.
// ref: util/ovl_sort.c (reader :977-1046, order :246-262, merge with the per-read pseudo record :930-975, filters :446-741),
// lib/ovl.c:150-204 (decode_ovl), :109-150 (encode_ovl).  The per-read stage is nd_ovlsort.cuh (one thread per read).
//
// Decoding a delta/varint stream in parallel: a byte < 128 ends a value, so the value index of every byte is a prefix count of
// terminators, a value is assembled by its terminator's thread from the <= 4 continuation bytes before it, and the running names
// (each record stores +-delta to the previous one) are prefix sums of the signed deltas mod 2^32.
#pragma once
#include <cuda_runtime.h>
#include "nd_ovlsort.cuh"
#include "nd_kernels.cuh"

__global__ void k_os_term(const uint8_t *b, nd_u64 n, uint32_t *flag)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = b[i] < 128 ? 1u : 0u;
    if (i == n) flag[i] = 0;
}
__global__ void k_os_values(const uint8_t *b, nd_u64 n, const nd_u64 *vidx, uint32_t *vals)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || b[i] >= 128) return;
    nd_u64 v = b[i]; int sh = 7;
    for (nd_u64 j = i; j > 0 && b[j - 1] >= 128 && sh < 64; --j, sh += 7) v |= (nd_u64)(b[j - 1] & 127) << sh;
    vals[vidx[i]] = (uint32_t)v;
}
// signed name deltas of record r (fields 0 and 4, signs in bits 1 and 2 of field 1), as wrapping u32
__global__ void k_os_deltas(const uint32_t *vals, nd_u64 n_rec, uint32_t *dq, uint32_t *dt)
{
    nd_u64 r = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint32_t *v = vals + 8 * r;
    dq[r] = (v[1] & 2) ? 0u - v[0] : v[0];
    dt[r] = (v[1] & 4) ? 0u - v[4] : v[4];
}
// seeds: sorted names + lengths (binary search); returns the length or 0xffffffff
__device__ __forceinline__ uint32_t nd_os_seed_len(const uint32_t *names, const uint32_t *lens, nd_u64 n, uint32_t name)
{
    nd_u64 lo = 0, hi = n;
    while (lo < hi) { const nd_u64 mid = (lo + hi) >> 1; if (names[mid] < name) lo = mid + 1; else hi = mid; }
    return lo < n && names[lo] == name ? lens[lo] : 0xffffffffu;
}
// reader filter of one stream (ref: util/ovl_sort.c:1003-1042): which of the two orientations of record r qualify, and whether a
// seed test failed (the reader gives up on an orientation after 5 failures in a file)
__global__ void k_os_cand(const uint32_t *vals, const uint32_t *qn, const uint32_t *tn, nd_u64 n_rec, const uint32_t *s_name, const uint32_t *s_len, nd_u64 n_seed,
                          uint8_t *ok, uint32_t *fail1, uint32_t *fail5)
{
    nd_u64 r = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rec) return;
    if (r == n_rec) { fail1[r] = fail5[r] = 0; return; }
    const uint32_t *v = vals + 8 * r;
    const uint32_t q = qn[r], t = tn[r], qs = v[2], qe = v[2] + v[3], ts = v[5];
    const uint32_t te = (v[1] & 8) ? v[5] + v[3] + v[6] : v[5] + v[3] - v[6];
    uint8_t o = 0; uint32_t f1 = 0, f5 = 0;
    if (q != t && !(qe - qs < 500 || te - ts < 500)) {
        o = 4;                                                      // reached the seed tests
        const uint32_t lq = nd_os_seed_len(s_name, s_len, n_seed, q), lt = nd_os_seed_len(s_name, s_len, n_seed, t);
        if (lq != 0xffffffffu && lq >= qe) o |= 1; else f1 = 1;
        if (lt != 0xffffffffu && lt >= te) o |= 2; else f5 = 1;
    }
    ok[r] = o; fail1[r] = f1; fail5[r] = f5;
}
// final say per orientation: the seed test must pass and fewer than 5 failures may precede it in this stream
__global__ void k_os_keep(const uint8_t *ok, const nd_u64 *nf1, const nd_u64 *nf5, nd_u64 n_rec, uint32_t *keep)
{
    nd_u64 r = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_rec) return;
    if (r == n_rec) { keep[2 * r] = 0; return; }
    keep[2 * r] = (ok[r] & 1) && nf1[r] < 5 ? 1u : 0u;
    keep[2 * r + 1] = (ok[r] & 2) && nf5[r] < 5 ? 1u : 0u;
}
__global__ void k_os_scatter(const uint32_t *vals, const uint32_t *qn, const uint32_t *tn, nd_u64 n_rec, const uint32_t *keep, const nd_u64 *pos, nd_u64 base, ndos_rec *out)
{
    nd_u64 r = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const uint32_t *v = vals + 8 * r;
    const uint32_t qs = v[2], qe = v[2] + v[3], ts = v[5], te = (v[1] & 8) ? v[5] + v[3] + v[6] : v[5] + v[3] - v[6];
    if (keep[2 * r]) { ndos_rec o = { qn[r], v[1] & 1, qs, qe - 1, tn[r], ts, te - 1, v[7] }; out[base + pos[2 * r]] = o; }
    if (keep[2 * r + 1]) { ndos_rec o = { tn[r], v[1] & 1, ts, te - 1, qn[r], qs, qe - 1, v[7] }; out[base + pos[2 * r + 1]] = o; }
}
// sort keys: pass 1 by query span ascending, pass 2 (stable) by qname ascending then match descending
__global__ void k_os_key_span(const ndos_rec *r, nd_u64 n, uint32_t *key, uint32_t *idx)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { key[i] = (uint32_t)((int)(r[i].qe - r[i].qs)) ^ 0x80000000u; idx[i] = (uint32_t)i; }   // signed order
}
__global__ void k_os_key_main(const ndos_rec *r, const uint32_t *idx, nd_u64 n, nd_u64 *key)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const ndos_rec &o = r[idx[i]]; key[i] = (nd_u64)o.qname << 32 | (uint32_t)~o.match; }
}
__global__ void k_os_starts(const ndos_rec *r, const uint32_t *idx, nd_u64 n, uint32_t *start)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) start[i] = i == 0 || r[idx[i]].qname != r[idx[i - 1]].qname ? 1u : 0u;
    if (i == n) start[i] = 0;
}
// records in sorted order with the read's own record in front of each read (ref: :945-953): element i of the sorted list goes to
// i + (reads started up to and including i); a read's first element also writes the pseudo record and the read table
__global__ void k_os_layout(const ndos_rec *r, const uint32_t *idx, const uint32_t *start, const nd_u64 *nstart_excl, nd_u64 n, const uint32_t *s_name, const uint32_t *s_len, nd_u64 n_seed,
                            ndos_rec *out, nd_u64 *q_off, uint32_t *q_len)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ndos_rec o = r[idx[i]];
    const nd_u64 ord = nstart_excl[i] + start[i];              // 1-based ordinal of the read this record belongs to
    out[i + ord] = o;
    if (start[i]) {
        const uint32_t ql = nd_os_seed_len(s_name, s_len, n_seed, o.qname);
        ndos_rec p = { o.qname, 0, 0, ql - 1, o.qname, 0, ql - 1, 0 };
        out[i + ord - 1] = p;
        q_off[ord - 1] = i + ord - 1; q_len[ord - 1] = ql;
    }
}
__global__ void k_os_binlen(const uint32_t *q_len, nd_u64 nq, uint32_t *bl)
{
    nd_u64 q = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) bl[q] = (q_len[q] >> NDOS_BIN_OFFSET) + 4;
    if (q == nq) bl[q] = 0;
}
// the per-read stage: one thread per read (sequential by nature: every admission depends on the histogram the earlier ones left)
__global__ void __launch_bounds__(64)
k_os_reads(ndos_par P, const ndos_rec *recs, const nd_u64 *q_off, const uint32_t *q_len, nd_u64 nq, nd_u64 n_tot, const nd_u64 *bin_off, uint16_t *bins, uint32_t *acc, uint8_t *dead, char *label)
{
    nd_u64 q = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const nd_u64 b = q_off[q], e = q + 1 < nq ? q_off[q + 1] : n_tot;
    uint32_t n_out;
    label[q] = ndos_one_read(P, q_len[q], recs + b, (uint32_t)(e - b), bins + bin_off[q], acc + b, dead + b, &n_out);
}
__global__ void k_os_live(const ndos_rec *recs, const uint8_t *dead, nd_u64 n, uint8_t *live)
{
    nd_u64 i = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) live[i] = !dead[i] && recs[i].qe ? 1 : 0;
}
__device__ __forceinline__ void nd_os_fields(const ndos_rec *recs, const uint32_t *kidx, nd_u64 j, uint32_t f[8])
{
    const ndos_rec o = recs[kidx[j]];
    uint32_t pq = 0, pt = 0;
    if (j > 0) { const ndos_rec g = recs[kidx[j - 1]]; pq = g.qname; pt = g.tname; }
    const uint32_t ql = o.qe - o.qs, tl = o.te - o.ts;
    uint32_t flags = o.rev;
    if (o.qname >= pq) f[0] = o.qname - pq; else { flags |= 2; f[0] = pq - o.qname; }
    if (o.tname >= pt) f[4] = o.tname - pt; else { flags |= 4; f[4] = pt - o.tname; }
    if (ql >= tl) f[6] = ql - tl; else { flags |= 8; f[6] = tl - ql; }
    f[1] = flags; f[2] = o.qs; f[3] = ql; f[5] = o.ts; f[7] = o.match;
}
__global__ void k_os_len(const ndos_rec *recs, const uint32_t *kidx, nd_u64 nk, uint32_t *len)
{
    nd_u64 j = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nk) { uint32_t f[8]; nd_os_fields(recs, kidx, j, f); int nb = 0; for (int i = 0; i < 8; ++i) nb += nd_varint_len(f[i]); len[j] = (uint32_t)nb; }
    if (j == nk) len[j] = 0;
}
__global__ void k_os_write(const ndos_rec *recs, const uint32_t *kidx, nd_u64 nk, const nd_u64 *off, uint8_t *out)
{
    nd_u64 j = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nk) return;
    uint32_t f[8]; nd_os_fields(recs, kidx, j, f);
    uint8_t *dst = out + off[j];
    for (int i = 0; i < 8; ++i) dst += nd_varint_put(dst, f[i]);
}
__global__ void k_os_qnames(const ndos_rec *recs, const nd_u64 *q_off, nd_u64 nq, uint32_t *qname)
{
    nd_u64 q = (nd_u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) qname[q] = recs[q_off[q]].qname;
}
