// ndgpu.cu -- C-ABI implementation (include/ndgpu.h): device memory, streams, kernel sequencing.
// Compute happens only in the sm_100a kernels of nd_kernels.cuh / nd_warp.cuh; there is no CPU path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <string>
#include <vector>
#include <memory>
#include <algorithm>
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include "../../include/ndgpu.h"
#include "nd_kernels.cuh"
#include "nd_warp.cuh"
#include "nd_ksw.cuh"
#include "nd_kswrx.cuh"
#include "nd_ovlsort_gpu.cuh"
#include "nd_align.hpp"
#include "nd_step2.hpp"
#include <omp.h>
#include <parallel/algorithm>
#include <omp.h>

static_assert(sizeof(nd_hit) == sizeof(ndg_hit_t), "hit layout");
static_assert(sizeof(nd_hit) == 56, "hit size");

// ------------------------------------------------------------------------------------------------
namespace {

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 2 + 256;      // generous slack: regrowth (cudaFree + cudaMalloc) stalls the stream for ~100 ms
        if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); p = nullptr; return -1; } want = bytes; }
        cap = want;
        return 0;
    }
    int grow_keep(size_t bytes, size_t used, cudaStream_t st) {   // preserves the first `used` bytes
        if (bytes <= cap) return 0;
        void *np = nullptr; size_t want = std::max(bytes + bytes / 2, (size_t)1 << 20);
        if (cudaMalloc(&np, want) != cudaSuccess) { cudaGetLastError(); want = bytes; if (cudaMalloc(&np, want) != cudaSuccess) { cudaGetLastError(); return -1; } }
        if (p && used) cudaMemcpyAsync(np, p, used, cudaMemcpyDeviceToDevice, st);
        cudaStreamSynchronize(st);
        if (p) cudaFree(p);
        p = np; cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostBuf {   // pinned
    void *p = nullptr; size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        if (cudaMallocHost(&p, want) != cudaSuccess) { cudaGetLastError(); return -1; }
        cap = want; return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostTargets {         // 2-bit target reads kept on the host for regime B / step 2 (what mm_idx_t::S is for, ref: minimap2/index.c:152-162)
    std::vector<uint32_t> words, nmask; std::vector<uint64_t> word_off; bool has_n = false;
};

struct ProfEntry { std::string name; double ms = 0; uint64_t launches = 0; };
struct ProfPending { int id; cudaEvent_t a, b; };

struct ReadsDev {            // a batch of reads resident on the device
    DevBuf words, word_off, len, nmask, name_key, name_hash, name_id, chunk_off, tag; bool has_tag = false;
    int n_reads = 0; uint64_t n_words = 0, n_chunks = 0, n_bases = 0; bool has_nmask = false; uint32_t max_len = 0;
    nd_reads_dev view() const {
        nd_reads_dev v;
        v.words = words.as<uint32_t>(); v.word_off = word_off.as<nd_u64>(); v.len = len.as<uint32_t>();
        v.nmask = has_nmask ? nmask.as<uint32_t>() : nullptr;
        v.name_key = name_key.as<nd_u64>(); v.name_hash = name_hash.as<uint32_t>(); v.name_id = name_id.as<uint32_t>(); v.tag = has_tag ? tag.as<uint32_t>() : nullptr; v.n_reads = n_reads;
        return v;
    }
    void release() { words.release(); word_off.release(); len.release(); nmask.release(); name_key.release(); name_hash.release(); name_id.release(); chunk_off.release(); tag.release(); }
};

} // namespace

struct ndg_ctx {
    int device = 0;
    cudaStream_t st = nullptr;
    ndg_opt_t opt;
    std::string err;
    uint64_t sub_batch_bases = 64ull << 20, max_anchors = 0;
    int sort_variant = 1, chain_variant = 1, ksw_variant = 3;   // ksw: 0 scalar, 1 packed memory-resident, 2 block-per-call, 3 register-resident (nd_kswrx.cuh)
    int chunk_bases = 512;
    int n_sm = 148;
    void *hk_scope = nullptr;
    uint64_t launches = 0;
    // profiling
    std::vector<ProfEntry> prof; std::vector<ProfPending> pending; std::vector<cudaEvent_t> ev_pool;
    // index
    DevBuf tmz; uint64_t n_tmz = 0;               // raw target minimizers (x,y)
    DevBuf t_key, t_len, t_id, t_tag; std::vector<uint64_t> h_tkey; std::vector<uint32_t> h_tlen, h_tid, h_ttag; bool t_has_id = true, t_has_tag = false; int tag_shift = 0;
    DevBuf em_keep, em_kidx, em_len, em_off, em_out; HostBuf h_ovl; uint32_t n_targets = 0; uint64_t t_sum_len = 0;
    DevBuf idx_pos, idx_tab, idx_pos_base; uint64_t n_pos = 0; bool index_ready = false;
    uint32_t n_slices = 1; uint64_t slice_cap = 0, slice_slots = 0, key_mul = 0; int key_bits = 0;      // geometry of the ordered table (nd_kernels.cuh)
    // transient buffers of an index build (ndg_index_finalize / ndg_index_shard_*)
    DevBuf ib_lkey, ib_lval, ib_rkey, ib_rval, ib_skey, ib_sval, ib_ukey, ib_cnt, ib_start, ib_d;
    uint64_t ib_n_local = 0, ib_n_recv = 0, ib_n_sorted = 0, ib_n_distinct = 0; std::vector<uint64_t> ib_bounds;
    // occurrence cutoff (ref: minimap2/options.c:66-76): a user value (-f >= 1) wins; otherwise the value derived from the FIRST index
    // part finalized after ndg_set_opt() is kept for the later parts of the run.  idx_hist/idx_distinct describe the current index,
    // so the derivation can be redone (a borrower calling ndg_set_opt() after ndg_index_share()).
    int mid_occ_user = 0, mid_occ_run = 0; bool opt_fresh = true;
    std::vector<uint32_t> idx_hist; uint64_t idx_distinct = 0;
    int mid_occ() const { return mid_occ_user > 0 ? mid_occ_user : mid_occ_run; }
    // mapping workspace
    ReadsDev q;
    DevBuf chunk_owner, staging, cnt, slow, pos, mz, mz_off, m_off, m_n, m_keep, a_pos, a, a2, a_off, avgq, jobs, flags, seg_start,
           ksw_calls, ksw_seqs, ksw_res, ksw_p, ksw_cig, ksw_scratch,
           f, p, v, t, e, hits, hits_sorted, zkey, zkey2, hidx, hidx2, hq, hq2, cubtmp, scalars;
    HostBuf h_scalars, h_hits, h_stage;
    uint64_t n_mz = 0, n_a = 0, n_seg = 0, n_hits = 0;
    int overflow_flag = 0;
    // upload staging
    HostBuf up;
    // a query batch uploaded ahead of its mapping call (ndg_map_prefetch): second read buffer, copy stream, staging of its own
    ReadsDev q_next; HostBuf up_next; cudaStream_t st_copy = nullptr; cudaEvent_t ev_next = nullptr;
    const void *pf_words = nullptr; const void *pf_len = nullptr; int pf_n = 0; bool pf_valid = false;
    ndg_reads_t pf_req; bool pf_req_valid = false;      // a prefetch asked for while the previous one had not been taken yet: started by the mapping call that takes it
    // regime B (base-level alignment)
    bool aln_on = false; ndal::Opt aln; ndg_aln_stat_t aln_stat = {};
    std::shared_ptr<HostTargets> ht = std::make_shared<HostTargets>();   // host copy of the target bases (shared with borrowers of the index)
    std::vector<ndg_hit_t> aln_hits;
    DevBuf ksw_ncig, ksw_coff, ksw_cout, ch_cnt, ch_off, ch_a, ch_fend;
    HostBuf h_kseq, h_kres, h_kcig, h_koff, h_cha, h_chf;
    uint64_t arena_want = 0;
    int n_threads = 1;                            // host threads of the alignment stage
    bool idx_borrowed = false;                    // index buffers belong to another context (ndg_index_share)
    // --step 2
    bool s2_on = false; nds2::Opt s2; nds2::Writer s2w; size_t s2_emitted = 0; std::string s2_bl; ndg_ctx *remap = nullptr;
    uint64_t s2_remaps = 0, s2_groups = 0; double s2_remap_ms = 0;
    // PAF extras (ndg_set_paf)
    bool paf_on = false;
    DevBuf pf_flag, pf_rank, pf_mp, pf_q, pf_he; HostBuf h_pf;
    std::vector<ndg_hit_extra_t> paf_he; std::vector<ndg_query_extra_t> paf_qe; std::vector<uint32_t> paf_cig;
    // ndg_ovl_sort results
    HostBuf os_out; std::string os_bl;
};

// ------------------------------------------------------------------------------------------------
namespace {

// CPUs this process may really use: the affinity mask, capped by a cgroup CPU quota when there is one (a container that shows
// 128 logical CPUs may be allowed 16 CPUs' worth of time; oversubscribing it makes every OpenMP region slower)
int usable_cpus()
{
    int n = omp_get_num_procs();
    long quota = -1, period = 0;
    if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {                       // cgroup v2
        char a[64];
        if (fscanf(f, "%63s %ld", a, &period) == 2 && strcmp(a, "max") != 0) quota = atol(a);
        fclose(f);
    } else {                                                                    // cgroup v1
        if (FILE *g = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) { if (fscanf(g, "%ld", &quota) != 1) quota = -1; fclose(g); }
        if (FILE *g = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) { if (fscanf(g, "%ld", &period) != 1) period = 0; fclose(g); }
    }
    if (quota > 0 && period > 0) { int q = (int)((quota + period - 1) / period); if (q >= 1 && q < n) n = q; }
    return n < 1 ? 1 : n;
}

int fail(ndg_ctx *c, int code, const char *fmt, ...)
{
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    c->err = buf;
    return code;
}
#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail(c, NDG_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); } while (0)
#define RSV(buf, bytes) do { if ((buf).reserve(bytes)) return fail(c, NDG_ERR_NOMEM, "%s:%d out of device memory reserving %zu bytes for " #buf, __FILE__, __LINE__, (size_t)(bytes)); } while (0)

int prof_id(ndg_ctx *c, const char *name)
{
    for (size_t i = 0; i < c->prof.size(); ++i) if (c->prof[i].name == name) return (int)i;
    ProfEntry e; e.name = name; c->prof.push_back(e); return (int)c->prof.size() - 1;
}
cudaEvent_t ev_get(ndg_ctx *c)
{
    if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct ProfScope {
    ndg_ctx *c; ProfPending pp;
    ProfScope(ndg_ctx *c_, const char *name) : c(c_) { pp.id = prof_id(c, name); pp.a = ev_get(c); pp.b = ev_get(c); cudaEventRecord(pp.a, c->st); }
    ~ProfScope() { cudaEventRecord(pp.b, c->st); c->pending.push_back(pp); c->prof[pp.id].launches++; c->launches++; }
};
void prof_flush(ndg_ctx *c)
{
    for (auto &pp : c->pending) {
        float ms = 0; cudaEventSynchronize(pp.b);
        if (cudaEventElapsedTime(&ms, pp.a, pp.b) == cudaSuccess) c->prof[pp.id].ms += ms;
        c->ev_pool.push_back(pp.a); c->ev_pool.push_back(pp.b);
    }
    c->pending.clear();
}
#define KLAUNCH(name, ...) do { ProfScope ps__(c, name); __VA_ARGS__; } while (0)

void hk_begin(void *ctx, const char *name) { ndg_ctx *c = (ndg_ctx *)ctx; c->hk_scope = new ProfScope(c, name); }
void hk_end(void *ctx) { ndg_ctx *c = (ndg_ctx *)ctx; delete (ProfScope *)c->hk_scope; c->hk_scope = nullptr; }

inline unsigned nblk(uint64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// exclusive prefix sum u32 -> u64 (n+1 outputs: out[n] = total)
int scan_u32_to_u64(ndg_ctx *c, const uint32_t *in, nd_u64 *out, uint64_t n)
{
    size_t tmp = 0;
    auto it = cub::TransformInputIterator<nd_u64, cub::CastOp<nd_u64>, const uint32_t *>(in, cub::CastOp<nd_u64>());
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, it, out, (long long)(n + 1), c->st));
    RSV(c->cubtmp, tmp);
    KLAUNCH("scan", CK(cub::DeviceScan::ExclusiveSum(c->cubtmp.p, tmp, it, out, (long long)(n + 1), c->st)));
    return 0;
}

// copy a host read batch (or a slice of it) to the device; computes chunk offsets
int upload_reads(ndg_ctx *c, ReadsDev &D, const ndg_reads_t *R, int r0, int r1, cudaStream_t st_ = nullptr, HostBuf *stage_ = nullptr, bool wait = true)
{
    cudaStream_t st = st_ ? st_ : c->st;
    HostBuf &up = stage_ ? *stage_ : c->up;
    int n = r1 - r0;
    if (n <= 0) return fail(c, NDG_ERR_ARG, "empty read batch");
    uint64_t w0 = R->word_off[r0], w1 = R->word_off[r1], nw = w1 - w0;
    size_t hdr = (size_t)(n + 1) * 8 * 2 + (size_t)n * (4 + 8 + 4 + 4);
    if (up.reserve(hdr)) return fail(c, NDG_ERR_NOMEM, "pinned staging");
    uint64_t *h_woff = up.as<uint64_t>(), *h_coff = h_woff + (n + 1), *h_key = h_coff + (n + 1);
    uint32_t *h_len = reinterpret_cast<uint32_t *>(h_key + n), *h_hash = h_len + n, *h_id = h_hash + n;
    uint64_t chunks = 0, bases = 0;
    D.max_len = 0;
    for (int i = 0; i < n; ++i) {
        h_woff[i] = R->word_off[r0 + i] - w0;
        h_coff[i] = chunks;
        uint32_t L = R->len[r0 + i];
        if (L > D.max_len) D.max_len = L;
        chunks += (L + c->chunk_bases - 1) / c->chunk_bases;
        bases += L;
        h_len[i] = L;
        h_key[i] = R->name_key ? R->name_key[r0 + i] : (uint64_t)(r0 + i);
        h_hash[i] = R->name_hash ? R->name_hash[r0 + i] : 0;
        h_id[i] = R->name_id ? R->name_id[r0 + i] : (uint32_t)(r0 + i);
    }
    h_woff[n] = nw; h_coff[n] = chunks;
    D.n_reads = n; D.n_words = nw; D.n_chunks = chunks; D.n_bases = bases; D.has_nmask = R->nmask != nullptr;
    RSV(D.words, (nw + 4) * 4); RSV(D.word_off, (size_t)(n + 1) * 8); RSV(D.chunk_off, (size_t)(n + 1) * 8);
    RSV(D.len, (size_t)n * 4); RSV(D.name_key, (size_t)n * 8); RSV(D.name_hash, (size_t)n * 4); RSV(D.name_id, (size_t)n * 4);
    CK(cudaMemcpyAsync(D.words.p, R->words + w0, nw * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(D.word_off.p, h_woff, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(D.chunk_off.p, h_coff, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(D.len.p, h_len, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(D.name_key.p, h_key, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(D.name_hash.p, h_hash, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(D.name_id.p, h_id, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    if (R->nmask) {
        for (int i = 0; i <= n; ++i) if (R->word_off[r0 + i] & 1) return fail(c, NDG_ERR_ARG, "nmask requires an even word offset for every read (read %d)", r0 + i);
        RSV(D.nmask, (nw / 2 + 4) * 4);
        CK(cudaMemcpyAsync(D.nmask.p, R->nmask + w0 / 2, ((nw + 1) / 2) * 4, cudaMemcpyHostToDevice, st));   // exactly the mask words the slice covers
    }
    D.has_tag = R->tag != nullptr;
    if (R->tag) { RSV(D.tag, (size_t)n * 4); CK(cudaMemcpyAsync(D.tag.p, R->tag + r0, (size_t)n * 4, cudaMemcpyHostToDevice, st)); }
    if (wait) CK(cudaStreamSynchronize(st));   // staging buffer is reused by the next call
    return 0;
}

// sketch the device-resident reads D; result packed into `dst` (+dst_used entries kept), offsets into off_out
// (n_reads+1 entries, relative to dst_used).  Returns number of minimizers via *n_out.
int sketch_reads(ndg_ctx *c, const ReadsDev &D, uint32_t rid0, int use_rid, DevBuf &dst, uint64_t dst_used, bool keep,
                 nd_u64 *off_out, uint64_t *n_out)
{
    const ndg_opt_t &o = c->opt;
    if (o.k < 1 || o.k > 128 || o.w < 1 || o.w > 255) return fail(c, NDG_ERR_ARG, "k must be in [1,128] and w in [1,255]");
    int C = c->chunk_bases, S = C + o.w + 8;
    uint64_t nc = D.n_chunks;
    *n_out = 0;
    RSV(c->scalars, 1024);
    if (nc == 0) { if (off_out) CK(cudaMemsetAsync(off_out, 0, (size_t)(D.n_reads + 1) * 8, c->st)); return 0; }
    RSV(c->staging, nc * (size_t)S * 16); RSV(c->cnt, (nc + 1) * 4); RSV(c->pos, (nc + 1) * 8);
    int *d_over = c->scalars.as<int>();
    CK(cudaMemsetAsync(d_over, 0, 4, c->st));
    CK(cudaMemsetAsync(c->cnt.as<uint32_t>() + nc, 0, 4, c->st));
    nd_reads_dev V = D.view();
    // fast path (closed form) where its preconditions can hold; flagged reads fall through to the state machine
    const bool fast = !o.hpc && !D.has_nmask && o.w <= NDS_MAXW && C == NDS_C && o.k <= 28;
    RSV(c->slow, (size_t)D.n_reads * 4 + 4);
    CK(cudaMemsetAsync(c->slow.p, fast ? 0 : 1, (size_t)D.n_reads * 4, c->st));
    static const bool old_fast = getenv("NDG_SKETCH_OLD") != nullptr;
    if (fast && !old_fast && (o.w == 5 || o.w == 10)) {
        RSV(c->chunk_owner, (nc + 1) * 4);
        KLAUNCH("misc", (k_chunk_owner<<<nblk(nc, 256), 256, 0, c->st>>>(D.chunk_off.as<nd_u64>(), D.n_reads, nc, c->chunk_owner.as<uint32_t>())));
    }
    if (fast && !old_fast && o.w == 5)
        KLAUNCH("sketch_fast", (k_sketch_fast4<5><<<(unsigned)nc, NDS_T, 0, c->st>>>(V, D.chunk_off.as<nd_u64>(), c->chunk_owner.as<uint32_t>(), nc, o.k, rid0, use_rid,
                                c->staging.as<nd_u128>(), S, c->cnt.as<uint32_t>(), c->slow.as<uint32_t>())));
    else if (fast && !old_fast && o.w == 10)
        KLAUNCH("sketch_fast", (k_sketch_fast4<10><<<(unsigned)nc, NDS_T, 0, c->st>>>(V, D.chunk_off.as<nd_u64>(), c->chunk_owner.as<uint32_t>(), nc, o.k, rid0, use_rid,
                                c->staging.as<nd_u128>(), S, c->cnt.as<uint32_t>(), c->slow.as<uint32_t>())));
    else if (fast)
        KLAUNCH("sketch_fast", (k_sketch_fast<<<(unsigned)nc, NDS_T, 0, c->st>>>(V, D.chunk_off.as<nd_u64>(), nc, o.w, o.k, rid0, use_rid,
                                c->staging.as<nd_u128>(), S, c->cnt.as<uint32_t>(), c->slow.as<uint32_t>())));
#define ND_SKETCH_LAUNCH(RW, LK) KLAUNCH("sketch", (k_sketch_chunks<RW, LK><<<nblk(nc, 128), 128, 0, c->st>>>(V, D.chunk_off.as<nd_u64>(), nc, C, o.w, o.k, o.hpc, rid0, use_rid, \
                                        c->staging.as<nd_u128>(), S, c->cnt.as<uint32_t>(), d_over, c->slow.as<uint32_t>())))
    if (o.k <= 28) { if (o.w < 16) ND_SKETCH_LAUNCH(16, false); else ND_SKETCH_LAUNCH(256, false); }
    else { if (o.w < 64) ND_SKETCH_LAUNCH(64, true); else ND_SKETCH_LAUNCH(256, true); }           // long k-mers (ava-hifi: k = w = 51)
#undef ND_SKETCH_LAUNCH
    CK(cudaGetLastError());
    if (int rc = scan_u32_to_u64(c, c->cnt.as<uint32_t>(), c->pos.as<nd_u64>(), nc)) return rc;
    uint64_t *h = c->h_scalars.as<uint64_t>();
    CK(cudaMemcpyAsync(h, c->pos.as<nd_u64>() + nc, 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(h + 1, d_over, 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    uint64_t total = h[0];
    if ((int)h[1]) return fail(c, NDG_ERR_OVERFLOW, "sketch staging overflow");
    if (keep) { if (dst.grow_keep((dst_used + total + 1) * 16, dst_used * 16, c->st)) return fail(c, NDG_ERR_NOMEM, "minimizer buffer"); }
    else RSV(dst, (dst_used + total + 1) * 16);
    KLAUNCH("sketch_compact", (k_sketch_compact<<<nblk(nc * 32, 256), 256, 0, c->st>>>(c->staging.as<nd_u128>(), S, c->cnt.as<uint32_t>(), c->pos.as<nd_u64>(), nc,
                               dst.as<nd_u128>() + dst_used)));
    CK(cudaGetLastError());
    if (off_out) {
        KLAUNCH("misc", (k_gather_off<<<nblk(D.n_reads + 1, 256), 256, 0, c->st>>>(c->pos.as<nd_u64>(), D.chunk_off.as<nd_u64>(), D.n_reads + 1, off_out)));
        CK(cudaGetLastError());
    }
    *n_out = total;
    return 0;
}

// ref: minimap2/index.c:164-185 (mm_idx_cal_max_occ) + minimap2/options.c:66-76: value at rank (1-f)*n_distinct of the
// per-minimizer occurrence counts, +1, floored by --min-occ-floor.  idx_hist[c] = number of distinct minimizers occurring c times.
int derive_mid_occ(ndg_ctx *c, int *out)
{
    int v;
    if (c->opt.mid_occ_frac <= 0.f) v = INT32_MAX;
    else {
        const int HB = (int)c->idx_hist.size();
        if (HB == 0) return fail(c, NDG_ERR_STATE, "no index statistics to derive mid_occ from");
        uint32_t kth = (uint32_t)((1. - c->opt.mid_occ_frac) * c->idx_distinct);
        uint64_t acc = 0; int vbin = -1;
        for (int b = 0; b < HB; ++b) { acc += c->idx_hist[b]; if (acc > kth) { vbin = b; break; } }
        if (vbin < 0 || vbin == HB - 1) return fail(c, NDG_ERR_OVERFLOW, "mid_occ order statistic beyond histogram range");
        v = vbin + 1;
    }
    if (v < c->opt.min_mid_occ) v = c->opt.min_mid_occ;
    *out = v;
    return 0;
}

} // namespace

// ------------------------------------------------------------------------------------------------
// lifecycle
// ------------------------------------------------------------------------------------------------
extern "C" const char *ndg_version(void) { return "ndgpu 0.1 (sm_100a)"; }

extern "C" int ndg_init(int device, ndg_ctx_t **out)
{
    if (!out) return NDG_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) { cudaGetLastError(); return NDG_ERR_CUDA; }
    if (cudaSetDevice(device) != cudaSuccess) return NDG_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NDG_ERR_CUDA;
    if (prop.major != 10) { fprintf(stderr, "[ndgpu] device %d is sm_%d%d; this library carries sm_100a code only\n", device, prop.major, prop.minor); return NDG_ERR_CUDA; }
    ndg_ctx *c = new ndg_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) { delete c; return NDG_ERR_CUDA; }
    memset(&c->opt, 0, sizeof c->opt);
    c->opt.k = 15; c->opt.w = 10; c->opt.min_cnt = 3; c->opt.min_chain_score = 40; c->opt.bw = 500; c->opt.max_gap = 5000;
    c->opt.max_chain_skip = 25; c->opt.max_chain_iter = 5000; c->opt.seed = 11; c->opt.mid_occ_frac = 2e-4f;
    if (c->h_scalars.reserve(4096)) { delete c; return NDG_ERR_NOMEM; }
    c->n_threads = usable_cpus();
    c->n_sm = prop.multiProcessorCount;
    cudaDeviceSetLimit(cudaLimitStackSize, 8192);
    *out = c;
    return NDG_OK;
}

static void drop_index_loan(ndg_ctx *c)
{
    if (!c->idx_borrowed) return;
    DevBuf *b[] = { &c->idx_pos, &c->idx_tab, &c->idx_pos_base, &c->t_key, &c->t_len, &c->t_id, &c->t_tag };
    for (auto x : b) { x->p = nullptr; x->cap = 0; }
    c->idx_borrowed = false; c->index_ready = false;
    c->ht = std::make_shared<HostTargets>();
}

extern "C" void ndg_destroy(ndg_ctx_t *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->st);
    drop_index_loan(c);
    prof_flush(c);
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    DevBuf *bufs[] = { &c->tmz, &c->t_key, &c->t_len, &c->t_id, &c->em_keep, &c->em_kidx, &c->em_len, &c->em_off, &c->em_out, &c->idx_pos, &c->idx_tab, &c->idx_pos_base, &c->ib_lkey, &c->ib_lval, &c->ib_rkey, &c->ib_rval, &c->ib_skey, &c->ib_sval, &c->ib_ukey, &c->ib_cnt, &c->ib_start, &c->ib_d, &c->staging, &c->cnt, &c->slow, &c->pos, &c->mz, &c->mz_off, &c->m_off, &c->m_n,
                       &c->m_keep, &c->a_pos, &c->a, &c->a2, &c->a_off, &c->ksw_calls, &c->ksw_seqs, &c->ksw_res, &c->ksw_p, &c->ksw_cig, &c->ksw_scratch, &c->avgq, &c->jobs, &c->flags, &c->seg_start, &c->f, &c->p, &c->v, &c->t, &c->e,
                       &c->hits, &c->hits_sorted, &c->zkey, &c->zkey2, &c->hidx, &c->hidx2, &c->hq, &c->hq2, &c->cubtmp, &c->scalars,
                       &c->ksw_ncig, &c->ksw_coff, &c->ksw_cout, &c->ch_cnt, &c->ch_off, &c->ch_a, &c->ch_fend, &c->t_tag,
                       &c->pf_flag, &c->pf_rank, &c->pf_mp, &c->pf_q, &c->pf_he, &c->chunk_owner };
    for (auto b : bufs) b->release();
    c->q.release(); c->q_next.release(); c->up_next.release();
    if (c->st_copy) { cudaStreamDestroy(c->st_copy); cudaEventDestroy(c->ev_next); }
    c->h_ovl.release(); c->h_scalars.release(); c->h_hits.release(); c->h_stage.release(); c->up.release();
    c->h_kseq.release(); c->h_kres.release(); c->h_kcig.release(); c->h_koff.release(); c->h_cha.release(); c->h_chf.release(); c->h_pf.release(); c->os_out.release();
    if (c->remap) { ndg_destroy(c->remap); c->remap = nullptr; }
    cudaStreamDestroy(c->st);
    delete c;
}

extern "C" void *ndg_host_alloc(uint64_t bytes) { void *p = nullptr; if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
extern "C" void ndg_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" const char *ndg_last_error(const ndg_ctx_t *c) { return c ? c->err.c_str() : "null context"; }
extern "C" uint64_t ndg_kernel_launches(const ndg_ctx_t *c) { return c ? c->launches : 0; }

extern "C" int ndg_set_opt(ndg_ctx_t *c, const ndg_opt_t *o)
{
    if (!c || !o) return NDG_ERR_ARG;
    c->opt = *o;
    c->mid_occ_user = o->mid_occ > 0 ? o->mid_occ : 0;
    // the derived cutoff belongs to a run: the first index part finalized after this call sets it anew; if an index is
    // already in place (finalized or borrowed) it is re-derived from that index right away, so mapping never sees 0
    c->mid_occ_run = 0; c->opt_fresh = true;
    if (c->index_ready && derive_mid_occ(c, &c->mid_occ_run)) return NDG_ERR_OVERFLOW;
    return NDG_OK;
}
extern "C" int ndg_set_limits(ndg_ctx_t *c, uint64_t sub_batch_bases, uint64_t max_anchors)
{
    if (!c) return NDG_ERR_ARG;
    if (sub_batch_bases) c->sub_batch_bases = sub_batch_bases;
    c->max_anchors = max_anchors;
    return NDG_OK;
}
extern "C" int ndg_set_ksw_variant(ndg_ctx_t *c, int v) { if (!c) return NDG_ERR_ARG; c->ksw_variant = v; return NDG_OK; }
extern "C" int ndg_set_variant(ndg_ctx_t *c, int sort_variant, int chain_variant)
{
    if (!c) return NDG_ERR_ARG;
    c->sort_variant = sort_variant; c->chain_variant = chain_variant;
    return NDG_OK;
}

extern "C" int ndg_prof_get(ndg_ctx_t *c, int max, const char **names, double *ms, uint64_t *launches)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    prof_flush(c);
    int n = std::min<int>(max, (int)c->prof.size());
    for (int i = 0; i < n; ++i) { names[i] = c->prof[i].name.c_str(); ms[i] = c->prof[i].ms; launches[i] = c->prof[i].launches; }
    return n;
}
extern "C" int ndg_prof_reset(ndg_ctx_t *c)
{
    if (!c) return NDG_ERR_ARG;
    prof_flush(c);
    for (auto &e : c->prof) { e.ms = 0; e.launches = 0; }
    c->aln_stat = ndg_aln_stat_t();
    return NDG_OK;
}

// ------------------------------------------------------------------------------------------------
// index
// ------------------------------------------------------------------------------------------------
extern "C" int ndg_index_begin(ndg_ctx_t *c)
{
    if (!c) return NDG_ERR_ARG;
    cudaSetDevice(c->device);
    drop_index_loan(c);
    c->n_tmz = 0; c->n_targets = 0; c->t_sum_len = 0; c->h_tkey.clear(); c->h_tlen.clear(); c->h_tid.clear(); c->h_ttag.clear(); c->t_has_id = true; c->t_has_tag = false; c->tag_shift = 0;
    c->index_ready = false; c->n_pos = 0;
    c->ht = std::make_shared<HostTargets>();          // never clear in place: a borrower may still hold the previous store
    return NDG_OK;
}

static int index_meta(ndg_ctx *c, const ndg_reads_t *T)
{
    for (int i = 0; i < T->n_reads; ++i) {
        c->h_tkey.push_back(T->name_key ? T->name_key[i] : (uint64_t)(c->n_targets + i));
        c->h_tlen.push_back(T->len[i]);
        c->h_tid.push_back(T->name_id ? T->name_id[i] : 0);
        c->h_ttag.push_back(T->tag ? T->tag[i] : 0);
        c->t_sum_len += T->len[i];
    }
    if (T->tag) c->t_has_tag = true;
    if (!T->name_id) c->t_has_id = false;
    c->n_targets += T->n_reads;
    return 0;
}

extern "C" int ndg_index_add_meta(ndg_ctx_t *c, const ndg_reads_t *T)
{
    if (!c || !T) return NDG_ERR_ARG;
    return index_meta(c, T);
}

extern "C" int ndg_index_add(ndg_ctx_t *c, const ndg_reads_t *T)
{
    if (!c || !T || T->n_reads <= 0) return NDG_ERR_ARG;
    if (c->idx_borrowed) return fail(c, NDG_ERR_STATE, "this context borrows another context's index: call ndg_index_begin() before building its own");
    if (T->nmask) for (int i = 0; i <= T->n_reads; ++i) if (T->word_off[i] & 1) return fail(c, NDG_ERR_ARG, "nmask requires an even word offset for every read (read %d)", i);
    cudaSetDevice(c->device);
    uint32_t rid0 = c->n_targets;
    int r0 = 0;
    while (r0 < T->n_reads) {           // device sub-batches bounded by bases
        int r1 = r0; uint64_t bases = 0;
        while (r1 < T->n_reads && (r1 == r0 || bases + T->len[r1] <= c->sub_batch_bases)) bases += T->len[r1++];
        if (int rc = upload_reads(c, c->q, T, r0, r1)) return rc;
        uint64_t n = 0;
        if (int rc = sketch_reads(c, c->q, rid0 + r0, 1, c->tmz, c->n_tmz, true, nullptr, &n)) return rc;
        c->n_tmz += n;
        r0 = r1;
    }
    if (c->aln_on || c->s2_on) {        // regime B and step 2 read target bases on the host (ref: mm_idx_getseq, minimap2/index.c:152-162)
        const uint64_t wb = T->word_off[0], nw = T->word_off[T->n_reads] - wb, w0 = c->ht->words.size();   // offsets may be absolute into a larger array
        if (T->nmask && !c->ht->has_n) { c->ht->nmask.assign(w0 / 2, 0); c->ht->has_n = true; }
        c->ht->words.insert(c->ht->words.end(), T->words + wb, T->words + wb + nw);
        if (c->ht->words.size() & 1) c->ht->words.push_back(0);
        if (c->ht->has_n) {
            c->ht->nmask.resize(c->ht->words.size() / 2 + 1, 0);
            if (T->nmask) memcpy(c->ht->nmask.data() + w0 / 2, T->nmask + wb / 2, (size_t)((nw + 1) / 2) * 4);
        }
        for (int i = 0; i < T->n_reads; ++i) c->ht->word_off.push_back(w0 + T->word_off[i] - wb);
    }
    return index_meta(c, T);
}

// ------------------------------------------------------------------------------------------------
// index build: sort by minimizer -> run lengths -> ordered table (nd_kernels.cuh).  The same stages serve one GPU (one slice)
// and N GPUs (slice r built on rank r from the keys of its range, finished slices all-gathered): ndg_index_shard_*.
// ------------------------------------------------------------------------------------------------
namespace {

void idx_stage_release(ndg_ctx *c)
{
    DevBuf *b[] = { &c->ib_lkey, &c->ib_lval, &c->ib_rkey, &c->ib_rval, &c->ib_skey, &c->ib_sval, &c->ib_ukey, &c->ib_cnt, &c->ib_start, &c->ib_d };
    for (auto x : b) x->release();
    c->ib_n_local = c->ib_n_recv = c->ib_n_sorted = c->ib_n_distinct = 0;
}

// key geometry of the part being built: keys are < key_range (2k hash bits, group tags above them for a tagged index)
int idx_key_geometry(ndg_ctx *c)
{
    int bits = std::min(2 * c->opt.k, 56);             // long k-mers (k > 28) are hashed down to 64 bits and keep the low 56 (x = hash << 8 | span)
    c->tag_shift = 0;
    if (c->t_has_tag) {
        uint32_t mx = 0; for (uint32_t t : c->h_ttag) mx = std::max(mx, t);
        int tb = 1; while (tb < 32 && (mx >> tb)) ++tb;
        if (bits + tb > 63) return fail(c, NDG_ERR_ARG, "group tags up to %u do not fit above %d hash bits", mx, bits);
        c->tag_shift = bits; bits += tb;
        RSV(c->t_tag, (size_t)c->n_targets * 4 + 4);
        CK(cudaMemcpyAsync(c->t_tag.p, c->h_ttag.data(), (size_t)c->n_targets * 4, cudaMemcpyHostToDevice, c->st));
    }
    c->key_bits = bits; c->key_mul = ~0ull / (1ull << bits);
    return 0;
}

// (key, val) pairs sorted by key on key_bits bits; stable LSD radix => ascending position inside a key (ref: index.c:230)
int idx_sort(ndg_ctx *c, const nd_u64 *kin, const nd_u64 *vin, uint64_t n, nd_u64 *kout, nd_u64 *vout)
{
    size_t tmp = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin, kout, vin, vout, (long long)n, 0, c->key_bits, c->st));
    RSV(c->cubtmp, tmp);
    KLAUNCH("index_sort", CK(cub::DeviceRadixSort::SortPairs(c->cubtmp.p, tmp, kin, kout, vin, vout, (long long)n, 0, c->key_bits, c->st)));
    return 0;
}

// distinct keys, their counts and run starts of a sorted key array
int idx_runs(ndg_ctx *c, const nd_u64 *skey, uint64_t n)
{
    RSV(c->ib_ukey, (n + 1) * 8); RSV(c->ib_cnt, (n + 1) * 4); RSV(c->ib_start, (n + 2) * 8); RSV(c->scalars, 1024);
    nd_u64 *d_nd = c->scalars.as<nd_u64>() + 8;
    size_t tmp = 0;
    CK(cub::DeviceRunLengthEncode::Encode(nullptr, tmp, skey, c->ib_ukey.as<nd_u64>(), c->ib_cnt.as<uint32_t>(), d_nd, (long long)n, c->st));
    RSV(c->cubtmp, tmp);
    KLAUNCH("index_runs", CK(cub::DeviceRunLengthEncode::Encode(c->cubtmp.p, tmp, skey, c->ib_ukey.as<nd_u64>(), c->ib_cnt.as<uint32_t>(), d_nd, (long long)n, c->st)));
    uint64_t *h = c->h_scalars.as<uint64_t>();
    CK(cudaMemcpyAsync(h, d_nd, 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    c->ib_n_distinct = h[0];
    CK(cudaMemsetAsync(c->ib_cnt.as<uint32_t>() + c->ib_n_distinct, 0, 4, c->st));
    return scan_u32_to_u64(c, c->ib_cnt.as<uint32_t>(), c->ib_start.as<nd_u64>(), c->ib_n_distinct);
}

struct MaxI64 { __host__ __device__ nd_i64 operator()(nd_i64 a, nd_i64 b) const { return a > b ? a : b; } };

// place the distinct keys (ib_ukey/cnt/start) into slice `sl` of the table; accumulates the run-length histogram in d_hist
int idx_place(ndg_ctx *c, uint32_t sl, const nd_u64 *svals, uint32_t *d_hist, int HB)
{
    const uint64_t nd = c->ib_n_distinct;
    nd_ht_slot *slice = c->idx_tab.as<nd_ht_slot>() + (uint64_t)sl * c->slice_slots;
    CK(cudaMemsetAsync(slice, 0xff, c->slice_slots * sizeof(nd_ht_slot), c->st));
    if (nd == 0) return 0;
    RSV(c->ib_d, nd * 8);
    KLAUNCH("index_table", (k_index_ideal<<<nblk(nd, 256), 256, 0, c->st>>>(c->ib_ukey.as<nd_u64>(), nd, c->key_mul, c->n_slices, c->slice_cap, c->ib_d.as<nd_i64>())));
    size_t tmp = 0;
    CK(cub::DeviceScan::InclusiveScan(nullptr, tmp, c->ib_d.as<nd_i64>(), c->ib_d.as<nd_i64>(), MaxI64(), (long long)nd, c->st));
    RSV(c->cubtmp, tmp);
    KLAUNCH("index_table", CK(cub::DeviceScan::InclusiveScan(c->cubtmp.p, tmp, c->ib_d.as<nd_i64>(), c->ib_d.as<nd_i64>(), MaxI64(), (long long)nd, c->st)));
    int *d_over = c->scalars.as<int>() + 2;
    CK(cudaMemsetAsync(d_over, 0, 4, c->st));
    KLAUNCH("index_table", (k_index_place<<<nblk(nd, 256), 256, 0, c->st>>>(c->ib_ukey.as<nd_u64>(), c->ib_cnt.as<uint32_t>(), c->ib_start.as<nd_u64>(), c->ib_d.as<nd_i64>(), nd, svals,
                            slice, c->slice_slots, d_hist, HB, d_over)));
    CK(cudaGetLastError());
    uint64_t *h = c->h_scalars.as<uint64_t>();
    CK(cudaMemcpyAsync(h, d_over, 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if ((int)h[0]) return fail(c, NDG_ERR_OVERFLOW, "index table: a key was displaced beyond the %llu-slot pad of its slice (non-uniform minimizer keys?)", (unsigned long long)ND_HT_PAD);
    return 0;
}

uint64_t idx_slice_cap(uint64_t max_distinct) { uint64_t cap = (uint64_t)(max_distinct / 0.70) + 64; return (cap + 7) & ~7ull; }

int idx_upload_meta(ndg_ctx *c)
{
    RSV(c->t_key, (size_t)c->n_targets * 8 + 8); RSV(c->t_len, (size_t)c->n_targets * 4 + 4); RSV(c->t_id, (size_t)c->n_targets * 4 + 4);
    CK(cudaMemcpyAsync(c->t_id.p, c->h_tid.data(), (size_t)c->n_targets * 4, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(c->t_key.p, c->h_tkey.data(), (size_t)c->n_targets * 8, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(c->t_len.p, c->h_tlen.data(), (size_t)c->n_targets * 4, cudaMemcpyHostToDevice, c->st));
    CK(cudaStreamSynchronize(c->st));
    return 0;
}

// last stage: statistics of the whole part are known (hist, n_distinct): cutoff, metadata, ready
int idx_finish(ndg_ctx *c, const std::vector<uint32_t> &hist, uint64_t n_distinct, uint64_t n_minimizers, ndg_index_stat_t *stat)
{
    if (int rc = idx_upload_meta(c)) return rc;
    // mid_occ: ref minimap2/index.c:164-185 + options.c:66-76 (computed on the first part of a run only)
    c->idx_hist = hist; c->idx_distinct = n_distinct;
    if (c->opt_fresh || c->mid_occ_run <= 0) { if (int rc = derive_mid_occ(c, &c->mid_occ_run)) return rc; }
    c->opt_fresh = false;
    c->n_pos = n_minimizers; c->index_ready = true;
    c->tmz.release(); c->n_tmz = 0;
    idx_stage_release(c);
    if (stat) {
        stat->n_seq = c->n_targets; stat->n_minimizers = n_minimizers; stat->n_distinct = n_distinct; stat->n_singletons = hist.size() > 1 ? hist[1] : 0;
        stat->sum_len = c->t_sum_len; stat->mid_occ = c->mid_occ();
    }
    return NDG_OK;
}

nd_idx_view idx_view(const ndg_ctx *c)
{
    nd_idx_view I; I.tab = c->idx_tab.as<nd_ht_slot>(); I.pos = c->idx_pos.as<nd_u64>(); I.pos_base = c->idx_pos_base.as<nd_u64>();
    I.slice_cap = c->slice_cap; I.slice_slots = c->slice_slots; I.key_mul = c->key_mul; I.n_slices = c->n_slices; I.key_bits = c->key_bits;
    return I;
}

const int IDX_HB = 65536;

} // namespace

extern "C" int ndg_index_finalize(ndg_ctx_t *c, ndg_index_stat_t *stat)
{
    if (!c) return NDG_ERR_ARG;
    if (c->idx_borrowed) return fail(c, NDG_ERR_STATE, "this context borrows another context's index: call ndg_index_begin() before building its own");
    cudaSetDevice(c->device);
    const uint64_t n = c->n_tmz;
    if (n == 0) return fail(c, NDG_ERR_STATE, "index has no minimizers");
    if (int rc = idx_key_geometry(c)) return rc;
    RSV(c->scalars, 1024);
    // one slice: split (x, y) -> (key, position), sort straight into the final position list
    RSV(c->ib_lkey, n * 8); RSV(c->ib_lval, n * 8); RSV(c->ib_skey, n * 8); RSV(c->idx_pos, n * 8 + 8);
    KLAUNCH("index_split", (k_split_xy<<<nblk(n, 256), 256, 0, c->st>>>(c->tmz.as<nd_u128>(), n, c->ib_lkey.as<nd_u64>(), c->ib_lval.as<nd_u64>(),
                            c->t_has_tag ? c->t_tag.as<uint32_t>() : nullptr, c->tag_shift, c->key_bits)));
    c->tmz.release(); c->n_tmz = 0;
    if (int rc = idx_sort(c, c->ib_lkey.as<nd_u64>(), c->ib_lval.as<nd_u64>(), n, c->ib_skey.as<nd_u64>(), c->idx_pos.as<nd_u64>())) return rc;
    CK(cudaStreamSynchronize(c->st));
    c->ib_lkey.release(); c->ib_lval.release();
    if (int rc = idx_runs(c, c->ib_skey.as<nd_u64>(), n)) return rc;
    c->ib_skey.release();
    c->n_slices = 1; c->slice_cap = idx_slice_cap(c->ib_n_distinct); c->slice_slots = c->slice_cap + ND_HT_PAD;
    RSV(c->idx_tab, c->slice_slots * sizeof(nd_ht_slot)); RSV(c->idx_pos_base, 64);
    CK(cudaMemsetAsync(c->idx_pos_base.p, 0, 64, c->st));
    DevBuf hist;
    if (hist.reserve(IDX_HB * 4)) return fail(c, NDG_ERR_NOMEM, "index histogram");
    cudaMemsetAsync(hist.p, 0, IDX_HB * 4, c->st);
    int rc = idx_place(c, 0, c->idx_pos.as<nd_u64>(), hist.as<uint32_t>(), IDX_HB);
    std::vector<uint32_t> hh(IDX_HB);
    if (!rc) { cudaMemcpyAsync(hh.data(), hist.p, IDX_HB * 4, cudaMemcpyDeviceToHost, c->st); if (cudaStreamSynchronize(c->st) != cudaSuccess) rc = fail(c, NDG_ERR_CUDA, "index table"); }
    hist.release();
    if (rc) return rc;
    return idx_finish(c, hh, c->ib_n_distinct, n, stat);
}

// ---- N-GPU build (no counterpart in the reference, whose index is built by one process) ----------------------------------
// rank r: ndg_index_add(own targets) + ndg_index_add_meta(the others), then
//   split(N)            local minimizers sorted by key and cut into the N key ranges              -> counts[N]
//   bucket(j)           device pointers of the pairs destined for rank j                           (all-to-all by the caller: NCCL or peer copies)
//   recv_reserve(total) device buffers the caller fills with everything destined for this rank
//   sort()              sort + run lengths of the received pairs                                   -> n_distinct of this slice
//   place(r, cap)       all N agree on slice_cap (max n_distinct / load); this rank's slice is built inside the full-size table
//   table_slice(j)      device pointer / bytes of slice j                                          (all-gather by the caller)
//   pos_reserve(n[N])   full position list allocated, own segment in place                         -> device pointer of segment j (all-gather by the caller)
//   finish(hist, nd)    global run-length histogram and distinct count (summed by the caller)      -> index ready
// device-to-device copy between buffers of two contexts (same or different GPUs; peer access or staged by the driver), synchronous.
// The collectives of a one-process N-GPU build (bin/minimap2-nd --gpus N) are made of these.
extern "C" int ndg_copy_between(ndg_ctx_t *dst_ctx, void *dst, ndg_ctx_t *src_ctx, const void *src, uint64_t bytes)
{
    if (!dst_ctx || !src_ctx || (bytes && (!dst || !src))) return NDG_ERR_ARG;
    ndg_ctx *c = dst_ctx;
    cudaSetDevice(c->device);
    if (bytes) CK(cudaMemcpyPeerAsync(dst, dst_ctx->device, src, src_ctx->device, bytes, c->st));
    CK(cudaStreamSynchronize(c->st));
    return NDG_OK;
}

// host helper (tests, planning): the slice / rank a minimizer key belongs to in an N-way sharded build of an untagged index
extern "C" int32_t ndg_index_key_slice(uint64_t key, int32_t k, int32_t n_slices)
{
    if (k < 1 || k > 128 || n_slices < 1) return -1;
    const int bits = std::min(2 * k, 56);
    return (int32_t)nd_idx_slice(nd_idx_mix(key, bits), ~0ull / (1ull << bits), (uint32_t)n_slices);
}

extern "C" int ndg_index_shard_split(ndg_ctx_t *c, int32_t n_slices, uint64_t *counts)
{
    if (!c || n_slices < 1 || !counts) return NDG_ERR_ARG;
    if (c->idx_borrowed) return fail(c, NDG_ERR_STATE, "borrowed index");
    if (c->t_has_tag) return fail(c, NDG_ERR_ARG, "a tagged index is built on one GPU");
    cudaSetDevice(c->device);
    if (int rc = idx_key_geometry(c)) return rc;
    RSV(c->scalars, 1024);
    const uint64_t n = c->n_tmz;
    c->n_slices = (uint32_t)n_slices; c->ib_n_local = n;
    RSV(c->ib_lkey, (n + 1) * 8); RSV(c->ib_lval, (n + 1) * 8); RSV(c->ib_skey, (n + 1) * 8); RSV(c->ib_sval, (n + 1) * 8);
    RSV(c->idx_pos_base, (size_t)(n_slices + 2) * 8);
    std::vector<uint64_t> bounds((size_t)n_slices + 1, 0);
    if (n) {
        KLAUNCH("index_split", (k_split_xy<<<nblk(n, 256), 256, 0, c->st>>>(c->tmz.as<nd_u128>(), n, c->ib_skey.as<nd_u64>(), c->ib_sval.as<nd_u64>(), nullptr, 0, c->key_bits)));
        if (int rc = idx_sort(c, c->ib_skey.as<nd_u64>(), c->ib_sval.as<nd_u64>(), n, c->ib_lkey.as<nd_u64>(), c->ib_lval.as<nd_u64>())) return rc;
        KLAUNCH("index_split", (k_index_slice_bounds<<<nblk(n + 1, 256), 256, 0, c->st>>>(c->ib_lkey.as<nd_u64>(), n, c->key_mul, c->n_slices, c->idx_pos_base.as<nd_u64>())));
        CK(cudaMemcpyAsync(bounds.data(), c->idx_pos_base.p, (size_t)(n_slices + 1) * 8, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
    }
    c->tmz.release(); c->n_tmz = 0;
    c->ib_bounds = bounds;
    for (int j = 0; j < n_slices; ++j) counts[j] = bounds[(size_t)j + 1] - bounds[(size_t)j];
    return NDG_OK;
}

extern "C" int ndg_index_shard_bucket(ndg_ctx_t *c, int32_t j, const void **keys, const void **vals, uint64_t *n)
{
    if (!c || !keys || !vals || !n || j < 0 || (size_t)j + 1 >= c->ib_bounds.size() + 0) return NDG_ERR_ARG;
    *keys = c->ib_lkey.as<nd_u64>() + c->ib_bounds[(size_t)j]; *vals = c->ib_lval.as<nd_u64>() + c->ib_bounds[(size_t)j];
    *n = c->ib_bounds[(size_t)j + 1] - c->ib_bounds[(size_t)j];
    return NDG_OK;
}

extern "C" int ndg_index_shard_recv_reserve(ndg_ctx_t *c, uint64_t total, void **keys, void **vals)
{
    if (!c || !keys || !vals) return NDG_ERR_ARG;
    cudaSetDevice(c->device);
    RSV(c->ib_rkey, (total + 1) * 8); RSV(c->ib_rval, (total + 1) * 8);
    c->ib_n_recv = total; *keys = c->ib_rkey.p; *vals = c->ib_rval.p;
    return NDG_OK;
}

extern "C" int ndg_index_shard_sort(ndg_ctx_t *c, uint64_t *n_distinct)
{
    if (!c || !n_distinct) return NDG_ERR_ARG;
    cudaSetDevice(c->device);
    const uint64_t n = c->ib_n_recv;
    c->ib_lkey.release(); c->ib_lval.release();      // the local buckets have been sent
    RSV(c->ib_skey, (n + 1) * 8); RSV(c->ib_sval, (n + 1) * 8);
    c->ib_n_sorted = n; c->ib_n_distinct = 0;
    if (n) {
        // pairs arrive as N sorted runs in rank order = target order, and the sort is stable: positions inside a key stay ascending
        if (int rc = idx_sort(c, c->ib_rkey.as<nd_u64>(), c->ib_rval.as<nd_u64>(), n, c->ib_skey.as<nd_u64>(), c->ib_sval.as<nd_u64>())) return rc;
        CK(cudaStreamSynchronize(c->st));
        c->ib_rkey.release(); c->ib_rval.release();
        if (int rc = idx_runs(c, c->ib_skey.as<nd_u64>(), n)) return rc;
    }
    c->ib_skey.release();
    *n_distinct = c->ib_n_distinct;
    return NDG_OK;
}

extern "C" int ndg_index_shard_place(ndg_ctx_t *c, int32_t my_slice, uint64_t max_distinct_per_slice, uint32_t *hist65536)
{
    if (!c || !hist65536 || my_slice < 0 || (uint32_t)my_slice >= c->n_slices) return NDG_ERR_ARG;
    cudaSetDevice(c->device);
    c->slice_cap = idx_slice_cap(max_distinct_per_slice); c->slice_slots = c->slice_cap + ND_HT_PAD;
    RSV(c->idx_tab, (size_t)c->n_slices * c->slice_slots * sizeof(nd_ht_slot));
    DevBuf hist;
    if (hist.reserve(IDX_HB * 4)) return fail(c, NDG_ERR_NOMEM, "index histogram");
    cudaMemsetAsync(hist.p, 0, IDX_HB * 4, c->st);
    int rc = idx_place(c, (uint32_t)my_slice, c->ib_sval.as<nd_u64>(), hist.as<uint32_t>(), IDX_HB);
    if (!rc) { cudaMemcpyAsync(hist65536, hist.p, IDX_HB * 4, cudaMemcpyDeviceToHost, c->st); if (cudaStreamSynchronize(c->st) != cudaSuccess) rc = fail(c, NDG_ERR_CUDA, "index table"); }
    hist.release();
    return rc;
}

extern "C" int ndg_index_shard_table_slice(ndg_ctx_t *c, int32_t j, void **dev, uint64_t *bytes)
{
    if (!c || !dev || !bytes || j < 0 || (uint32_t)j >= c->n_slices) return NDG_ERR_ARG;
    *dev = c->idx_tab.as<nd_ht_slot>() + (uint64_t)j * c->slice_slots; *bytes = c->slice_slots * sizeof(nd_ht_slot);
    return NDG_OK;
}

extern "C" int ndg_index_shard_pos_reserve(ndg_ctx_t *c, int32_t my_slice, const uint64_t *n_pos_per_slice, void **segments)
{
    if (!c || !n_pos_per_slice || !segments || my_slice < 0 || (uint32_t)my_slice >= c->n_slices) return NDG_ERR_ARG;
    cudaSetDevice(c->device);
    std::vector<uint64_t> base((size_t)c->n_slices + 1, 0);
    for (uint32_t j = 0; j < c->n_slices; ++j) base[j + 1] = base[j] + n_pos_per_slice[j];
    if (n_pos_per_slice[my_slice] != c->ib_n_sorted) return fail(c, NDG_ERR_ARG, "position count of the own slice does not match what was sorted");
    RSV(c->idx_pos, (base[c->n_slices] + 1) * 8); RSV(c->idx_pos_base, (size_t)(c->n_slices + 2) * 8);
    CK(cudaMemcpyAsync(c->idx_pos_base.p, base.data(), (size_t)(c->n_slices + 1) * 8, cudaMemcpyHostToDevice, c->st));
    if (c->ib_n_sorted) CK(cudaMemcpyAsync(c->idx_pos.as<nd_u64>() + base[(size_t)my_slice], c->ib_sval.p, c->ib_n_sorted * 8, cudaMemcpyDeviceToDevice, c->st));
    CK(cudaStreamSynchronize(c->st));
    c->ib_sval.release();
    for (uint32_t j = 0; j < c->n_slices; ++j) segments[j] = c->idx_pos.as<nd_u64>() + base[j];
    c->n_pos = base[c->n_slices];
    return NDG_OK;
}

extern "C" int ndg_index_shard_finish(ndg_ctx_t *c, const uint32_t *hist65536, uint64_t n_distinct, ndg_index_stat_t *stat)
{
    if (!c || !hist65536) return NDG_ERR_ARG;
    cudaSetDevice(c->device);
    std::vector<uint32_t> hh(hist65536, hist65536 + IDX_HB);
    return idx_finish(c, hh, n_distinct, c->n_pos, stat);
}

extern "C" int ndg_index_share(ndg_ctx_t *c, ndg_ctx_t *src)
{
    if (!c || !src || c == src) return NDG_ERR_ARG;
    if (!src->index_ready) return fail(c, NDG_ERR_STATE, "the source context has no finalized index");
    if (c->device != src->device) return fail(c, NDG_ERR_ARG, "contexts are on different devices");
    cudaSetDevice(c->device);
    cudaStreamSynchronize(src->st);
    drop_index_loan(c);
    DevBuf *own[] = { &c->idx_pos, &c->idx_tab, &c->idx_pos_base, &c->t_key, &c->t_len, &c->t_id, &c->t_tag };
    DevBuf *from[] = { &src->idx_pos, &src->idx_tab, &src->idx_pos_base, &src->t_key, &src->t_len, &src->t_id, &src->t_tag };
    for (int i = 0; i < 7; ++i) { own[i]->release(); own[i]->p = from[i]->p; own[i]->cap = from[i]->cap; }
    c->n_slices = src->n_slices; c->slice_cap = src->slice_cap; c->slice_slots = src->slice_slots; c->key_mul = src->key_mul; c->key_bits = src->key_bits;
    c->n_pos = src->n_pos; c->tag_shift = src->tag_shift;
    c->idx_hist = src->idx_hist; c->idx_distinct = src->idx_distinct;
    c->mid_occ_run = src->mid_occ();                  // the run's cutoff travels with the index; ndg_set_opt() on the borrower re-derives it
    c->ht = src->ht;                                  // regime B / step 2 on the borrower read the owner's host copy of the target bases
    c->n_targets = src->n_targets; c->t_sum_len = src->t_sum_len; c->t_has_id = src->t_has_id; c->t_has_tag = src->t_has_tag;
    c->h_tkey = src->h_tkey; c->h_tlen = src->h_tlen; c->h_tid = src->h_tid; c->h_ttag = src->h_ttag;
    c->idx_borrowed = true; c->index_ready = true;
    return NDG_OK;
}

extern "C" int ndg_index_get(ndg_ctx_t *c, uint64_t minier, uint64_t *pos, int cap, int *n)
{
    if (!c || !n) return NDG_ERR_ARG;
    if (!c->index_ready) return fail(c, NDG_ERR_STATE, "index not finalized");
    cudaSetDevice(c->device);
    RSV(c->scalars, 1024);
    nd_u64 *d = c->scalars.as<nd_u64>() + 16;
    k_index_get<<<1, 1, 0, c->st>>>(idx_view(c), minier, d);
    uint64_t *h = c->h_scalars.as<uint64_t>();
    CK(cudaMemcpyAsync(h, d, 24, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    *n = (int)h[1];
    int m = std::min<int>(cap, *n);
    if (m > 0 && pos) {
        if (h[2]) pos[0] = h[0];                       // a singleton carries its position in the slot
        else { CK(cudaMemcpyAsync(pos, c->idx_pos.as<nd_u64>() + h[0], (size_t)m * 8, cudaMemcpyDeviceToHost, c->st)); CK(cudaStreamSynchronize(c->st)); }
    }
    return NDG_OK;
}

// ------------------------------------------------------------------------------------------------
// sketch only (tests)
// ------------------------------------------------------------------------------------------------
extern "C" int ndg_sketch(ndg_ctx_t *c, const ndg_reads_t *R, uint32_t rid0, uint64_t *xy, uint64_t cap, uint64_t *off)
{
    if (!c || !R || !off) return NDG_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc = upload_reads(c, c->q, R, 0, R->n_reads)) return rc;
    RSV(c->mz_off, (size_t)(R->n_reads + 1) * 8);
    uint64_t n = 0;
    if (int rc = sketch_reads(c, c->q, rid0, 1, c->mz, 0, false, c->mz_off.as<nd_u64>(), &n)) return rc;
    CK(cudaMemcpyAsync(off, c->mz_off.p, (size_t)(R->n_reads + 1) * 8, cudaMemcpyDeviceToHost, c->st));
    if (n > cap) { cudaStreamSynchronize(c->st); return fail(c, NDG_ERR_OVERFLOW, "sketch output needs %llu entries", (unsigned long long)n); }
    if (n && xy) CK(cudaMemcpyAsync(xy, c->mz.p, n * 16, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return NDG_OK;
}

// ------------------------------------------------------------------------------------------------
// mapping
// ------------------------------------------------------------------------------------------------
namespace {

// all kernels of one device-resident sub-batch; hits left (sorted) in c->hits_sorted, count in c->n_hits
int map_resident(ndg_ctx *c, int qidx_base)
{
    const ndg_opt_t &o = c->opt;
    if (!c->index_ready) return fail(c, NDG_ERR_STATE, "index not finalized");
    if (c->mid_occ() <= 0) return fail(c, NDG_ERR_STATE, "no occurrence cutoff (mid_occ): neither given (-f >= 1) nor derived from an index part");
    ReadsDev &Q = c->q;
    const int nr = Q.n_reads;
    uint64_t *h = c->h_scalars.as<uint64_t>();
    c->n_mz = c->n_a = c->n_seg = c->n_hits = 0;
    RSV(c->scalars, 1024);

    // 1. sketch
    RSV(c->mz_off, (size_t)(nr + 1) * 8);
    if (int rc = sketch_reads(c, Q, 0, 0, c->mz, 0, false, c->mz_off.as<nd_u64>(), &c->n_mz)) return rc;
    const uint64_t M = c->n_mz;
    if (M == 0) return 0;

    // 2. probe + count
    nd_seed_par P; P.idx = idx_view(c);
    P.t_key = c->t_key.as<nd_u64>(); P.t_len = c->t_len.as<uint32_t>(); P.mid_occ = c->mid_occ(); P.no_diag = o.flag_no_diag; P.no_dual = o.flag_no_dual;
    P.tag_shift = c->tag_shift;
    if (c->tag_shift > 0 && !Q.has_tag) return fail(c, NDG_ERR_ARG, "the index is tagged: query reads need tags too");
    nd_reads_dev V = Q.view();
    if (M >= (1ULL << 32) - 64) return fail(c, NDG_ERR_OVERFLOW, "sub-batch yields %llu minimizers; lower sub_batch_bases", (unsigned long long)M);
    RSV(c->m_off, M * 8); RSV(c->m_n, M * 4); RSV(c->m_keep, (M + 1) * 4); RSV(c->a_pos, (M + 1) * 8);
    CK(cudaMemsetAsync(c->m_keep.as<uint32_t>() + M, 0, 4, c->st));
    KLAUNCH("probe_count", (k_probe_count<<<nblk(M, 256), 256, 0, c->st>>>(P, V, c->mz.as<nd_u128>(), c->mz_off.as<nd_u64>(), M, c->m_off.as<nd_u64>(),
                            c->m_n.as<uint32_t>(), c->m_keep.as<uint32_t>())));
    CK(cudaGetLastError());
    if (int rc = scan_u32_to_u64(c, c->m_keep.as<uint32_t>(), c->a_pos.as<nd_u64>(), M)) return rc;
    CK(cudaMemcpyAsync(h, c->a_pos.as<nd_u64>() + M, 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    const uint64_t NA = h[0];
    c->n_a = NA;
    if (NA == 0) return 0;
    if (NA >= (1ULL << 32) - 64) return fail(c, NDG_ERR_OVERFLOW, "sub-batch yields %llu anchors; lower sub_batch_bases", (unsigned long long)NA);

    // 3. fill + per-read stats
    RSV(c->a, (NA + 1) * 16); RSV(c->a_off, (size_t)(nr + 1) * 8); RSV(c->avgq, (size_t)nr * 4 + 4);
    KLAUNCH("fill_anchors", (k_fill_anchors<<<nblk(M, 256), 256, 0, c->st>>>(P, V, c->mz.as<nd_u128>(), c->mz_off.as<nd_u64>(), M, c->m_off.as<nd_u64>(),
                             c->m_n.as<uint32_t>(), c->m_keep.as<uint32_t>(), c->a_pos.as<nd_u64>(), c->a.as<nd_u128>())));
    KLAUNCH("read_stats", (k_read_stats<<<nblk((uint64_t)(nr + 1) * 32, 128), 128, 0, c->st>>>(c->mz.as<nd_u128>(), c->mz_off.as<nd_u64>(), c->m_keep.as<uint32_t>(),
                           c->a_pos.as<nd_u64>(), nr, c->a_off.as<nd_u64>(), c->avgq.as<float>())));
    CK(cudaGetLastError());

    // 4. exact sort (ref: minimap2/map.c:246 radix_sort_128x; tie order must match)
    if (c->sort_variant == 0) {
        RSV(c->jobs, (2 * (NA / 65) + 4 * (size_t)nr + 8) * sizeof(nd_rs_job));
        KLAUNCH("sort_seq", (k_sort_seq<<<nblk(nr, 64), 64, 0, c->st>>>(c->a.as<nd_u128>(), c->a_off.as<nd_u64>(), nr, c->jobs.as<nd_rs_job>())));
    } else {
        RSV(c->jobs, 2 * (NA / 65 + (size_t)nr + 2) * sizeof(nd_rs_job));
        // scratch of the warp sort aliases buffers the chain stage allocates anyway
        RSV(c->f, NA * 4); RSV(c->p, NA * 4); RSV(c->v, NA * 4); RSV(c->hits, (NA / (o.min_cnt > 0 ? o.min_cnt : 1) + 16) * sizeof(nd_hit));
        RSV(c->a2, (NA + 1) * 16);
        ndw_sort_bufs W = { c->a.as<nd_u128>(), c->a2.as<nd_u128>(), c->v.as<uint8_t>(), c->f.as<uint32_t>(), c->p.as<uint32_t>() };
        ndw_launch_hooks hk = { hk_begin, hk_end, c };
        // group filter scratch: c->flags (segment flags later, NA bytes) and c->t (chain scratch later, >= nr u32)
        RSV(c->flags, NA + 16); RSV(c->t, std::max<size_t>(NA * 4, (size_t)nr * 4 + 16));
        static const bool no_filter = getenv("NDG_NO_GROUP_FILTER") != nullptr;
        ndw_filter_par FP = { no_filter ? nullptr : c->flags.as<uint8_t>(), no_filter ? nullptr : c->t.as<uint32_t>(), o.min_cnt, o.min_chain_score, o.hpc ? 255 : o.k, (uint32_t)o.max_gap };
        ndw_sort_launch(c->st, W, c->a_off.as<nd_u64>(), nr, NA, c->jobs.as<nd_rs_job>(), c->scalars.as<uint32_t>() + 64, c->n_sm, hk, c->sort_variant != 2, FP);
    }
    CK(cudaGetLastError());

    // 5. segments
    RSV(c->flags, NA + 16); RSV(c->seg_start, (NA / 2 + 16) * 4);
    KLAUNCH("seg_flags", (k_seg_flags<<<nblk(NA, 256), 256, 0, c->st>>>(c->a.as<nd_u128>(), NA, (nd_u64)o.max_gap, c->flags.as<uint8_t>())));
    KLAUNCH("seg_flags", (k_seg_read_starts<<<nblk(nr, 256), 256, 0, c->st>>>(c->a_off.as<nd_u64>(), nr, c->flags.as<uint8_t>())));
    {
        // segments shorter than min_cnt can never produce a chain: they are dropped inside the chain kernel;
        // the list itself can have up to NA entries, so select into a buffer of NA u32.
        RSV(c->seg_start, (NA + 16) * 4);
        size_t tmp = 0;
        cub::CountingInputIterator<uint32_t> it(0);
        uint32_t *d_nsel = c->scalars.as<uint32_t>() + 4;
        CK(cub::DeviceSelect::Flagged(nullptr, tmp, it, c->flags.as<uint8_t>(), c->seg_start.as<uint32_t>(), d_nsel, (long long)NA, c->st));
        RSV(c->cubtmp, tmp);
        KLAUNCH("seg_select", CK(cub::DeviceSelect::Flagged(c->cubtmp.p, tmp, it, c->flags.as<uint8_t>(), c->seg_start.as<uint32_t>(), d_nsel, (long long)NA, c->st)));
        CK(cudaMemcpyAsync(h, d_nsel, 4, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        c->n_seg = (uint32_t)h[0];
    }

    // 6. chain
    const uint64_t hit_cap = NA / (o.min_cnt > 0 ? o.min_cnt : 1) + 16;
    RSV(c->f, NA * 4); RSV(c->p, NA * 4); RSV(c->v, NA * 4); RSV(c->t, NA * 4); RSV(c->e, NA * 4);
    RSV(c->hits, hit_cap * sizeof(nd_hit));
    nd_u64 *d_nh = c->scalars.as<nd_u64>() + 4;
    CK(cudaMemsetAsync(d_nh, 0, 8, c->st));
    nd_chain_par par = { o.max_gap, o.max_gap, o.bw, o.max_chain_skip, o.max_chain_iter, o.min_cnt, o.min_chain_score };
    nd_chain_bufs B = { c->f.as<int32_t>(), c->p.as<int32_t>(), c->v.as<int32_t>(), c->t.as<int32_t>(), c->e.as<int32_t>() };
    const int max_span = o.hpc ? 255 : o.k;
    nd_seg_view SV = { c->a.as<nd_u128>(), NA, c->seg_start.as<uint32_t>(), (uint32_t)c->n_seg, c->a_off.as<nd_u64>() };
    if (c->chain_variant == 0) {
        KLAUNCH("chain_fill_seq", (k_chain_fill_seq<<<nblk(c->n_seg, 64), 64, 0, c->st>>>(SV, V, c->avgq.as<float>(), par, max_span, B)));
    } else {
        KLAUNCH("chain_fill_warp", ndw_chain_fill_launch(c->st, SV, V, c->avgq.as<float>(), par, max_span, o.hpc ? 0 : o.k, B, c->scalars.as<uint32_t>() + 48, c->n_sm));
    }
    CK(cudaGetLastError());
    KLAUNCH("chain_backtrack", (k_chain_backtrack<<<nblk(c->n_seg, 64), 64, 0, c->st>>>(SV, V, par, max_span, o.seed, B, c->hits.as<nd_hit>(), hit_cap, d_nh)));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h, d_nh, 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    const uint64_t NH = h[0];
    if (NH > hit_cap) return fail(c, NDG_ERR_OVERFLOW, "hit buffer overflow (%llu > %llu)", (unsigned long long)NH, (unsigned long long)hit_cap);
    c->n_hits = NH;
    if (NH == 0) return 0;

    // 7. order hits: by query, then by zkey descending (ref: minimap2/hit.c:69-71)
    RSV(c->zkey, NH * 8); RSV(c->zkey2, NH * 8); RSV(c->hidx, NH * 4); RSV(c->hidx2, NH * 4); RSV(c->hq, NH * 4); RSV(c->hq2, NH * 4);
    RSV(c->hits_sorted, NH * sizeof(nd_hit));
    int qbits = 1; while ((1 << qbits) < nr + 1 && qbits < 31) ++qbits;
    int sbits = 1; while ((1ull << sbits) <= (uint64_t)Q.max_len && sbits < 32) ++sbits;       // chain score <= query length < 2^sbits
    if (32 + sbits + qbits <= 64) {                                                           // one sort on a composite key
        size_t t1 = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, t1, c->zkey.as<nd_u64>(), c->zkey2.as<nd_u64>(), c->hidx2.as<uint32_t>(), c->hidx.as<uint32_t>(), (long long)NH, 0, 32 + sbits + qbits, c->st));
        RSV(c->cubtmp, t1);
        KLAUNCH("hit_order", (k_hit_keys1<<<nblk(NH, 256), 256, 0, c->st>>>(c->hits.as<nd_hit>(), NH, 32 + sbits, c->zkey.as<nd_u64>(), c->hidx2.as<uint32_t>())));
        KLAUNCH("hit_order", CK(cub::DeviceRadixSort::SortPairs(c->cubtmp.p, t1, c->zkey.as<nd_u64>(), c->zkey2.as<nd_u64>(), c->hidx2.as<uint32_t>(), c->hidx.as<uint32_t>(), (long long)NH, 0, 32 + sbits + qbits, c->st)));
    } else {
    KLAUNCH("hit_order", (k_hit_keys<<<nblk(NH, 256), 256, 0, c->st>>>(c->hits.as<nd_hit>(), NH, c->zkey.as<nd_u64>(), c->hidx.as<uint32_t>())));
    {
        size_t t1 = 0, t2 = 0;
        CK(cub::DeviceRadixSort::SortPairsDescending(nullptr, t1, c->zkey.as<nd_u64>(), c->zkey2.as<nd_u64>(), c->hidx.as<uint32_t>(), c->hidx2.as<uint32_t>(), (long long)NH, 0, 64, c->st));
        CK(cub::DeviceRadixSort::SortPairs(nullptr, t2, c->hq.as<uint32_t>(), c->hq2.as<uint32_t>(), c->hidx2.as<uint32_t>(), c->hidx.as<uint32_t>(), (long long)NH, 0, qbits, c->st));
        RSV(c->cubtmp, std::max(t1, t2));
        KLAUNCH("hit_order", CK(cub::DeviceRadixSort::SortPairsDescending(c->cubtmp.p, t1, c->zkey.as<nd_u64>(), c->zkey2.as<nd_u64>(), c->hidx.as<uint32_t>(), c->hidx2.as<uint32_t>(), (long long)NH, 0, 64, c->st)));
        KLAUNCH("hit_order", (k_hit_qidx<<<nblk(NH, 256), 256, 0, c->st>>>(c->hits.as<nd_hit>(), c->hidx2.as<uint32_t>(), NH, c->hq.as<uint32_t>())));
        KLAUNCH("hit_order", CK(cub::DeviceRadixSort::SortPairs(c->cubtmp.p, t2, c->hq.as<uint32_t>(), c->hq2.as<uint32_t>(), c->hidx2.as<uint32_t>(), c->hidx.as<uint32_t>(), (long long)NH, 0, qbits, c->st)));
    }
    }
    KLAUNCH("hit_order", (k_hit_gather<<<nblk(NH, 256), 256, 0, c->st>>>(c->hits.as<nd_hit>(), c->hidx.as<uint32_t>(), NH, qidx_base, c->hits_sorted.as<nd_hit>())));
    CK(cudaGetLastError());
    return 0;
}

} // namespace

namespace {
// filter + encode the sorted hits of the resident sub-batch on the device; appends to c->h_ovl at h_used
int emit_resident(ndg_ctx *c, int qidx_base, const ndg_emit_opt_t *eo, uint32_t *d_state, uint64_t *h_used, uint64_t *n_rec, uint64_t *sum_qspan)
{
    const uint64_t NH = c->n_hits;
    if (NH == 0) return 0;
    uint64_t *h = c->h_scalars.as<uint64_t>();
    nd_emit_par P;
    P.q_id = c->q.name_id.as<uint32_t>(); P.q_len = c->q.len.as<uint32_t>(); P.q_key = c->q.name_key.as<nd_u64>();
    P.t_id = c->t_id.as<uint32_t>(); P.t_len = c->t_len.as<uint32_t>(); P.t_key = c->t_key.as<nd_u64>();
    P.qidx_base = qidx_base; P.minlen = eo->minlen; P.dvt = eo->dvt; P.maxhan1 = eo->maxhan1; P.maxhan2 = eo->maxhan2;
    RSV(c->em_keep, NH + 16); RSV(c->em_kidx, (NH + 16) * 4); RSV(c->em_len, (NH + 16) * 4); RSV(c->em_off, (NH + 16) * 8);
    const nd_hit *hits = c->hits_sorted.as<nd_hit>();
    KLAUNCH("emit", (k_emit_flags<<<nblk(NH, 256), 256, 0, c->st>>>(hits, NH, P, c->em_keep.as<uint8_t>())));
    uint32_t *d_nk = c->scalars.as<uint32_t>() + 160;
    nd_u64 *d_span = c->scalars.as<nd_u64>() + 96;      // byte 768; d_nk @640, new_state @656, d_state @672
    {
        size_t tmp = 0;
        cub::CountingInputIterator<uint32_t> it(0);
        CK(cub::DeviceSelect::Flagged(nullptr, tmp, it, c->em_keep.as<uint8_t>(), c->em_kidx.as<uint32_t>(), d_nk, (long long)NH, c->st));
        RSV(c->cubtmp, tmp);
        KLAUNCH("emit", CK(cub::DeviceSelect::Flagged(c->cubtmp.p, tmp, it, c->em_keep.as<uint8_t>(), c->em_kidx.as<uint32_t>(), d_nk, (long long)NH, c->st)));
    }
    CK(cudaMemcpyAsync(h, d_nk, 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    const uint64_t NK = (uint32_t)h[0];
    if (NK == 0) return 0;
    CK(cudaMemsetAsync(d_span, 0, 8, c->st));
    CK(cudaMemsetAsync(c->em_len.as<uint32_t>() + NK, 0, 4, c->st));
    KLAUNCH("emit", (k_emit_len<<<nblk(NK, 256), 256, 0, c->st>>>(hits, c->em_kidx.as<uint32_t>(), NK, P, d_state, c->em_len.as<uint32_t>(), d_span)));
    if (int rc = scan_u32_to_u64(c, c->em_len.as<uint32_t>(), c->em_off.as<nd_u64>(), NK)) return rc;
    RSV(c->em_out, NK * 40 + 64);
    KLAUNCH("emit", (k_emit_write<<<nblk(NK, 256), 256, 0, c->st>>>(hits, c->em_kidx.as<uint32_t>(), NK, P, d_state, c->em_off.as<nd_u64>(), c->em_out.as<uint8_t>(),
                     c->scalars.as<uint32_t>() + 164)));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h, c->em_off.as<nd_u64>() + NK, 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(h + 1, d_span, 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(d_state, c->scalars.as<uint32_t>() + 164, 8, cudaMemcpyDeviceToDevice, c->st));
    CK(cudaStreamSynchronize(c->st));
    const uint64_t nb = h[0];
    if (c->h_ovl.cap < *h_used + nb + 64) {
        HostBuf nbuf; if (nbuf.reserve((*h_used + nb + 64) * 2)) return fail(c, NDG_ERR_NOMEM, "pinned ovl buffer");
        if (*h_used) memcpy(nbuf.p, c->h_ovl.p, *h_used);
        c->h_ovl.release(); c->h_ovl = nbuf;
    }
    CK(cudaMemcpyAsync(c->h_ovl.as<uint8_t>() + *h_used, c->em_out.p, nb, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    *h_used += nb; *n_rec += NK; *sum_qspan += h[1];
    return 0;
}
} // namespace

namespace {
int align_resident(ndg_ctx *c, const ndg_reads_t *Q, int r0, int r1);

// anchors of every kept chain in chain order, for the sorted hits of the resident sub-batch (c->ch_off, c->ch_a, c->ch_fend)
int collect_chains(ndg_ctx *c, int r0, uint64_t *n_chain_anchors)
{
    const uint64_t NH = c->n_hits;
    uint64_t *h = c->h_scalars.as<uint64_t>();
    RSV(c->ch_cnt, (NH + 1) * 4); RSV(c->ch_off, (NH + 1) * 8); RSV(c->ch_fend, NH * 4);
    KLAUNCH("chain_collect", (k_hit_cnt<<<nblk(NH + 1, 256), 256, 0, c->st>>>(c->hits_sorted.as<nd_hit>(), NH, c->ch_cnt.as<uint32_t>())));
    if (int rc = scan_u32_to_u64(c, c->ch_cnt.as<uint32_t>(), c->ch_off.as<nd_u64>(), NH)) return rc;
    CK(cudaMemcpyAsync(h, c->ch_off.as<nd_u64>() + NH, 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    const uint64_t NCA = h[0];
    RSV(c->ch_a, (NCA + 1) * 16);
    KLAUNCH("chain_collect", (k_chain_collect<<<nblk(NH, 128), 128, 0, c->st>>>(c->hits_sorted.as<nd_hit>(), NH, r0, c->a_off.as<nd_u64>(), c->seg_start.as<uint32_t>(), (uint32_t)c->n_seg,
                              c->a.as<nd_u128>(), c->p.as<int32_t>(), c->f.as<int32_t>(), c->ch_off.as<nd_u64>(), c->ch_a.as<nd_u128>(), c->ch_fend.as<int32_t>())));
    CK(cudaGetLastError());
    *n_chain_anchors = NCA;
    return 0;
}

// PAF extras of the resident sub-batch (reads r0..r1 of the batch): per-query records appended to c->paf_qe; in regime A also the
// mm_est_err counts of every sorted hit, appended to c->paf_he (regime B fills paf_he from its aligned hits)
int paf_resident(ndg_ctx *c, int r0, int r1)
{
    const int nr = r1 - r0;
    const uint64_t M = c->n_mz, NH = c->n_hits;
    const size_t q_at = c->paf_qe.size();
    c->paf_qe.resize(q_at + (size_t)nr);
    memset(c->paf_qe.data() + q_at, 0, (size_t)nr * sizeof(ndg_query_extra_t));
    if (M == 0) return 0;
    RSV(c->pf_flag, (M + 1) * 4); RSV(c->pf_rank, (M + 1) * 8); RSV(c->pf_mp, (M + 1) * 8); RSV(c->pf_q, (size_t)nr * sizeof(nd_paf_query));
    KLAUNCH("paf", (k_paf_flags<<<nblk(M + 1, 256), 256, 0, c->st>>>(c->m_n.as<uint32_t>(), M, c->pf_flag.as<uint32_t>())));
    if (int rc = scan_u32_to_u64(c, c->pf_flag.as<uint32_t>(), c->pf_rank.as<nd_u64>(), M)) return rc;
    KLAUNCH("paf", (k_paf_minipos<<<nblk(M, 256), 256, 0, c->st>>>(c->mz.as<nd_u128>(), c->m_n.as<uint32_t>(), c->pf_rank.as<nd_u64>(), M, c->pf_mp.as<nd_u64>())));
    KLAUNCH("paf", (k_paf_query<<<nblk(nr, 64), 64, 0, c->st>>>(c->mz.as<nd_u128>(), c->mz_off.as<nd_u64>(), c->m_n.as<uint32_t>(), c->pf_rank.as<nd_u64>(), nr, c->pf_q.as<nd_paf_query>())));
    CK(cudaGetLastError());
    const bool est = !c->aln_on && NH > 0;
    if (est) {
        uint64_t NCA = 0;
        if (int rc = collect_chains(c, r0, &NCA)) return rc;
        RSV(c->pf_he, NH * 8);
        KLAUNCH("paf", (k_paf_esterr<<<nblk(NH, 128), 128, 0, c->st>>>(c->hits_sorted.as<nd_hit>(), NH, r0, c->q.len.as<uint32_t>(), c->mz_off.as<nd_u64>(), c->pf_rank.as<nd_u64>(),
                        c->pf_mp.as<nd_u64>(), c->ch_off.as<nd_u64>(), c->ch_a.as<nd_u128>(), c->pf_he.as<int32_t>())));
        CK(cudaGetLastError());
    }
    if (c->h_pf.reserve((size_t)nr * sizeof(nd_paf_query) + NH * 8 + 64)) return fail(c, NDG_ERR_NOMEM, "pinned PAF staging");
    CK(cudaMemcpyAsync(c->h_pf.p, c->pf_q.p, (size_t)nr * sizeof(nd_paf_query), cudaMemcpyDeviceToHost, c->st));
    int32_t *hhe = (int32_t *)((char *)c->h_pf.p + (size_t)nr * sizeof(nd_paf_query));
    if (est) CK(cudaMemcpyAsync(hhe, c->pf_he.p, NH * 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    const nd_paf_query *hq = (const nd_paf_query *)c->h_pf.p;
    for (int i = 0; i < nr; ++i) { ndg_query_extra_t &o = c->paf_qe[q_at + (size_t)i]; o.rep_len = hq[i].rep_len; o.n_mini = hq[i].n_mini; o.sum_span = hq[i].sum_span; }
    if (est) {
        const size_t h_at = c->paf_he.size();
        c->paf_he.resize(h_at + (size_t)NH);
        for (uint64_t i = 0; i < NH; ++i) {
            ndg_hit_extra_t &o = c->paf_he[h_at + (size_t)i]; memset(&o, 0, sizeof o);
            o.n_match = hhe[2 * i]; o.n_span = hhe[2 * i + 1];
        }
    }
    return 0;
}
}

namespace {
// the batch Q was uploaded ahead (same buffers, same size): make it the resident batch once its copy has finished
bool take_prefetched(ndg_ctx *c, const ndg_reads_t *Q)
{
    if (!c->pf_valid) return false;
    c->pf_valid = false;
    if (c->pf_words != (const void *)Q->words || c->pf_len != (const void *)Q->len || c->pf_n != Q->n_reads) return false;
    if (cudaStreamWaitEvent(c->st, c->ev_next, 0) != cudaSuccess) return false;
    std::swap(c->q, c->q_next);
    return true;
}
}

namespace {
int start_prefetch(ndg_ctx *c, const ndg_reads_t *Q)
{
    c->pf_valid = false;
    uint64_t bases = 0; for (int i = 0; i < Q->n_reads; ++i) bases += Q->len[i];
    if (bases > c->sub_batch_bases) return NDG_OK;                 // would be cut into sub-batches: the mapping call uploads as usual
    if (!c->st_copy) { CK(cudaStreamCreateWithFlags(&c->st_copy, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&c->ev_next, cudaEventDisableTiming)); }
    else CK(cudaEventSynchronize(c->ev_next));                     // the staging block of the previous prefetch is free again
    if (int rc = upload_reads(c, c->q_next, Q, 0, Q->n_reads, c->st_copy, &c->up_next, false)) return rc;
    CK(cudaEventRecord(c->ev_next, c->st_copy));
    c->pf_words = Q->words; c->pf_len = Q->len; c->pf_n = Q->n_reads; c->pf_valid = true;
    return NDG_OK;
}
}

extern "C" int ndg_map_prefetch(ndg_ctx_t *c, const ndg_reads_t *Q)
{
    if (!c || !Q || Q->n_reads <= 0) return NDG_ERR_ARG;
    cudaSetDevice(c->device);
    if (c->pf_valid) { c->pf_req = *Q; c->pf_req_valid = true; return NDG_OK; }     // second buffer still holds the batch about to be mapped
    c->pf_req_valid = false;
    return start_prefetch(c, Q);
}

extern "C" int ndg_map_batch_ovl(ndg_ctx_t *c, const ndg_reads_t *Q, const ndg_emit_opt_t *eo, uint32_t state[2],
                                 const uint8_t **ovl, uint64_t *n_bytes, uint64_t *n_rec, uint64_t *sum_qspan, ndg_map_result_t *work)
{
    if (!c || !Q || !eo || !state || !ovl || !n_bytes || Q->n_reads <= 0) return NDG_ERR_ARG;
    if (!Q->name_id || !c->t_has_id) return fail(c, NDG_ERR_ARG, "ndg_map_batch_ovl needs name_id for targets and queries");
    cudaSetDevice(c->device);
    if (c->aln_on) {                    // regime B: the aligned hits are on the host already; filter + encode there
        ndg_map_result_t R;
        if (int rc = ndg_map_batch(c, Q, &R)) return rc;
        if (c->h_ovl.reserve((size_t)R.n_hits * 48 + 64)) return fail(c, NDG_ERR_NOMEM, "pinned .ovl buffer");
        uint64_t rec = 0, sq = 0;
        const int64_t nb = ndg_ovl_emit_step1(R.hits, R.n_hits, Q->name_id, Q->name_key, Q->len, c->h_tid.data(), c->h_tkey.data(), c->h_tlen.data(),
                                              eo->minlen, eo->dvt, eo->maxhan1, eo->maxhan2, state, c->h_ovl.as<uint8_t>(), c->h_ovl.cap, &rec, &sq);
        if (nb < 0) return fail(c, NDG_ERR_OVERFLOW, "ovl buffer");
        *ovl = c->h_ovl.as<uint8_t>(); *n_bytes = (uint64_t)nb;
        if (n_rec) *n_rec += rec;
        if (sum_qspan) *sum_qspan += sq;
        if (work) { *work = R; work->hits = nullptr; }
        return NDG_OK;
    }
    RSV(c->scalars, 1024);
    uint32_t *d_state = c->scalars.as<uint32_t>() + 168;
    CK(cudaMemcpyAsync(d_state, state, 8, cudaMemcpyHostToDevice, c->st));
    uint64_t used = 0, rec = 0, sq = 0, tot_hits = 0, tot_mz = 0, tot_a = 0, tot_seg = 0;
    int r0 = 0;
    uint64_t budget = c->sub_batch_bases;
    while (r0 < Q->n_reads) {
        int r1 = r0; uint64_t bases = 0;
        while (r1 < Q->n_reads && (r1 == r0 || bases + Q->len[r1] <= budget)) bases += Q->len[r1++];
        if (r0 == 0 && r1 == Q->n_reads && take_prefetched(c, Q)) { /* uploaded ahead by ndg_map_prefetch() */ }
        else if (int rc = upload_reads(c, c->q, Q, r0, r1)) return rc;
        if (c->pf_req_valid) { c->pf_req_valid = false; if (int rc = start_prefetch(c, &c->pf_req)) return rc; }   // the next batch travels while this one is mapped
        int rc = map_resident(c, r0);
        if (rc == NDG_ERR_OVERFLOW && r1 - r0 > 1 && c->err.find("anchors") != std::string::npos) { budget = std::max<uint64_t>(bases / 2, 1); continue; }
        if (rc) return rc;
        if (int rc2 = emit_resident(c, r0, eo, d_state, &used, &rec, &sq)) return rc2;
        tot_hits += c->n_hits; tot_mz += c->n_mz; tot_a += c->n_a; tot_seg += c->n_seg;
        r0 = r1;
    }
    CK(cudaMemcpyAsync(c->h_scalars.p, d_state, 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    memcpy(state, c->h_scalars.p, 8);
    *ovl = c->h_ovl.as<uint8_t>(); *n_bytes = used;
    if (n_rec) *n_rec += rec;
    if (sum_qspan) *sum_qspan += sq;
    if (work) { work->n_hits = (int64_t)tot_hits; work->hits = nullptr; work->n_minimizers = tot_mz; work->n_anchors = tot_a; work->n_segments = tot_seg; }
    return NDG_OK;
}

extern "C" int ndg_map_upload(ndg_ctx_t *c, const ndg_reads_t *Q)
{
    if (!c || !Q || Q->n_reads <= 0) return NDG_ERR_ARG;
    cudaSetDevice(c->device);
    return upload_reads(c, c->q, Q, 0, Q->n_reads);
}

extern "C" int ndg_map_resident(ndg_ctx_t *c, ndg_map_result_t *out)
{
    if (!c) return NDG_ERR_ARG;
    cudaSetDevice(c->device);
    cudaEvent_t e0 = ev_get(c), e1 = ev_get(c);
    CK(cudaEventRecord(e0, c->st));
    int rc = map_resident(c, 0);
    if (rc) return rc;
    CK(cudaEventRecord(e1, c->st));
    CK(cudaStreamSynchronize(c->st));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    c->ev_pool.push_back(e0); c->ev_pool.push_back(e1);
    if (out) { out->n_hits = (int64_t)c->n_hits; out->hits = nullptr; out->n_minimizers = c->n_mz; out->n_anchors = c->n_a; out->n_segments = c->n_seg; out->device_ms = ms; }
    return NDG_OK;
}

extern "C" int ndg_map_download(ndg_ctx_t *c, ndg_map_result_t *out)
{
    if (!c || !out) return NDG_ERR_ARG;
    cudaSetDevice(c->device);
    if (c->h_hits.reserve((c->n_hits + 1) * sizeof(nd_hit))) return fail(c, NDG_ERR_NOMEM, "pinned hit buffer");
    if (c->n_hits) CK(cudaMemcpyAsync(c->h_hits.p, c->hits_sorted.p, c->n_hits * sizeof(nd_hit), cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    out->n_hits = (int64_t)c->n_hits; out->hits = c->h_hits.as<ndg_hit_t>();
    out->n_minimizers = c->n_mz; out->n_anchors = c->n_a; out->n_segments = c->n_seg;
    return NDG_OK;
}

extern "C" int ndg_map_batch(ndg_ctx_t *c, const ndg_reads_t *Q, ndg_map_result_t *out)
{
    if (!c || !Q || !out || Q->n_reads <= 0) return NDG_ERR_ARG;
    cudaSetDevice(c->device);
    if (c->aln_on) {
        if (c->ht->word_off.size() != c->n_targets) return fail(c, NDG_ERR_STATE, "regime B needs ndg_set_aln_opt() before ndg_index_add() (target bases are kept on the host)");
        c->aln_hits.clear();
    }
    c->paf_he.clear(); c->paf_qe.clear(); c->paf_cig.clear();
    uint64_t tot_hits = 0, tot_mz = 0, tot_a = 0, tot_seg = 0;
    int r0 = 0;
    uint64_t budget = c->sub_batch_bases;
    while (r0 < Q->n_reads) {
        int r1 = r0; uint64_t bases = 0;
        while (r1 < Q->n_reads && (r1 == r0 || bases + Q->len[r1] <= budget)) bases += Q->len[r1++];
        if (int rc = upload_reads(c, c->q, Q, r0, r1)) return rc;
        int rc = map_resident(c, r0);
        if (rc == NDG_ERR_OVERFLOW && r1 - r0 > 1 && c->err.find("anchors") != std::string::npos) { budget = std::max<uint64_t>(bases / 2, 1); continue; }
        if (rc) return rc;
        if (c->paf_on) if (int rc3 = paf_resident(c, r0, r1)) return rc3;
        if (c->aln_on) {
            if (int rc2 = align_resident(c, Q, r0, r1)) return rc2;
            tot_mz += c->n_mz; tot_a += c->n_a; tot_seg += c->n_seg;
            r0 = r1;
            continue;
        }
        if (c->h_hits.cap < (tot_hits + c->n_hits + 1) * sizeof(nd_hit)) {       // grow, preserving content
            HostBuf nb; if (nb.reserve((tot_hits + c->n_hits + 1) * sizeof(nd_hit) * 2)) return fail(c, NDG_ERR_NOMEM, "pinned hit buffer");
            if (tot_hits) memcpy(nb.p, c->h_hits.p, tot_hits * sizeof(nd_hit));
            c->h_hits.release(); c->h_hits = nb;
        }
        if (c->n_hits) CK(cudaMemcpyAsync(c->h_hits.as<nd_hit>() + tot_hits, c->hits_sorted.p, c->n_hits * sizeof(nd_hit), cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        if (c->paf_on && c->paf_he.size() == tot_hits + c->n_hits)             // chain scores next to the mm_est_err counts
            for (uint64_t i = 0; i < c->n_hits; ++i) { const nd_hit &hh = c->h_hits.as<nd_hit>()[tot_hits + i]; ndg_hit_extra_t &e = c->paf_he[tot_hits + i]; e.chain_score = e.score0 = hh.score; }
        tot_hits += c->n_hits; tot_mz += c->n_mz; tot_a += c->n_a; tot_seg += c->n_seg;
        r0 = r1;
    }
    out->n_hits = (int64_t)tot_hits; out->hits = c->h_hits.as<ndg_hit_t>();
    if (c->aln_on) { out->n_hits = (int64_t)c->aln_hits.size(); out->hits = c->aln_hits.data(); }
    out->n_minimizers = tot_mz; out->n_anchors = tot_a; out->n_segments = tot_seg;
    return NDG_OK;
}


// ------------------------------------------------------------------------------------------------
// register-resident ksw_extd2 (nd_kswrx.cuh): which ring a call needs, padded sequence layout, launch
// ------------------------------------------------------------------------------------------------
namespace {

const int RX_NCFG = 5;
const int RX_CAP[RX_NCFG] = { 512, 1024, 2048, 3072, 4096 };       // 128 * K * NW: (4,1) (4,2) (4,4) (3,8) (4,8)
#define NDX_FLAG_AMBI 0x10000

// ring class of a call, or -1 when the register-resident kernel does not take it (generic scoring, band wider than every ring)
inline int rx_class(int qlen, int tlen, int w, int flag)
{
    if (flag & NDK_GENERIC_SC) return -1;
    if (qlen <= 0 || tlen <= 0) return 0;
    const int T = (tlen + 15) / 16 * 16;
    if (w < 0) w = std::max(qlen, tlen);
    for (int i = 0; i < RX_NCFG; ++i) if (T <= RX_CAP[i] || (long long)w + 66 <= RX_CAP[i]) return i;
    return -1;
}

// bytes a call occupies in the padded layout (>= 32 zero bytes before the query, >= 48 after query and target, 4-byte aligned starts)
inline uint64_t rx_seq_bytes(int qlen, int tlen) { return 32 + (((uint64_t)qlen + 3) & ~3ull) + 48 + (((uint64_t)tlen + 3) & ~3ull) + 48; }
inline void rx_seq_offsets(uint64_t base, int qlen, uint64_t *q_off, uint64_t *t_off) { *q_off = base + 32; *t_off = base + 32 + (((uint64_t)qlen + 3) & ~3ull) + 48; }

template <int K, int NW, bool EX>
cudaError_t rx_launch_t(int blocks_per_sm_hint, int n_sm, cudaStream_t st, const ndk_call *calls, int n, const uint8_t *seqs, const ndk_par &P, ndk_res *res, uint8_t *pmat,
                        uint32_t *cig, unsigned long long *rowtabs, size_t tab_entries, uint32_t *work, unsigned long long *cur, unsigned long long arena_bytes, int *blocks_out)
{
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ksw_rx<K, NW, EX>, NW * 32, 0);
    if (occ < 1) occ = 1;
    if (blocks_per_sm_hint > 0 && occ > blocks_per_sm_hint) occ = blocks_per_sm_hint;
    int blocks = std::min(n, n_sm * occ);
    if (blocks_out) { *blocks_out = n_sm * occ; return cudaSuccess; }            // query only: how many row tables a launch may need
    k_ksw_rx<K, NW, EX><<<blocks, NW * 32, 0, st>>>(calls, n, seqs, P, res, pmat, cig, rowtabs, tab_entries, work, cur, arena_bytes);
    return cudaGetLastError();
}

cudaError_t rx_launch(int cfg, bool exact, int hint, int n_sm, cudaStream_t st, const ndk_call *calls, int n, const uint8_t *seqs, const ndk_par &P, ndk_res *res, uint8_t *pmat,
                      uint32_t *cig, unsigned long long *rowtabs, size_t tab_entries, uint32_t *work, unsigned long long *cur, unsigned long long arena_bytes, int *blocks_out = nullptr)
{
#define RXL(K, NW) (exact ? rx_launch_t<K, NW, true>(hint, n_sm, st, calls, n, seqs, P, res, pmat, cig, rowtabs, tab_entries, work, cur, arena_bytes, blocks_out) \
                          : rx_launch_t<K, NW, false>(hint, n_sm, st, calls, n, seqs, P, res, pmat, cig, rowtabs, tab_entries, work, cur, arena_bytes, blocks_out))
    switch (cfg) {
    case 0: return RXL(4, 1);
    case 1: return RXL(4, 2);
    case 2: return RXL(4, 4);
    case 3: return RXL(3, 8);
    default: return RXL(4, 8);
    }
#undef RXL
}

} // namespace

// ------------------------------------------------------------------------------------------------
// regime B: base-level alignment.  Host logic in nd_align.hpp; every DP call runs in k_ksw_extd2_px.
// ------------------------------------------------------------------------------------------------
namespace {

// One launch of the DP kernels with everything it owns; two of them are in flight at a time so that staging the next launch
// (sequence bytes, call table) and collecting the previous one overlap the kernels.
struct DpLaunch {
    HostBuf h_seq, h_res, h_cig, h_off;
    DevBuf calls, seqs, res, cig, scratch, ncig, coff, cout, tmp, cur;
    cudaStream_t st = nullptr, st2 = nullptr; cudaEvent_t ev_a = nullptr, ev_b = nullptr, e0 = nullptr, e1 = nullptr;
    std::vector<ndk_call> hc; std::vector<uint32_t> ids;
    struct Grp { int cfg; bool exact; size_t begin, count; };      // register-resident kernel: calls [begin, begin+count) share a ring size / maximum mode
    std::vector<Grp> rx; DevBuf rxtab; size_t n_legacy = 0;
    size_t n = 0, nb = 0; uint64_t seqb = 0, cig_cap = 0, fixed = 0, ext_need = 0;
    int Tmax = 16, rmax = 1, qmax = 1;
    bool busy = false;
    void release() {
        h_seq.release(); h_res.release(); h_cig.release(); h_off.release();
        DevBuf *d[] = { &calls, &seqs, &res, &cig, &scratch, &ncig, &coff, &cout, &tmp, &cur, &rxtab };
        for (auto x : d) x->release();
        if (st) { cudaStreamDestroy(st); cudaStreamDestroy(st2); cudaEventDestroy(ev_a); cudaEventDestroy(ev_b); cudaEventDestroy(e0); cudaEventDestroy(e1); st = nullptr; }
    }
};

struct GpuDp : ndal::Backend {
    ndg_ctx *c; ndk_par P; int rc = 0;
    DpLaunch L[2];
    bool any_ambi = false;                          // some read holds an ambiguous base: the register-resident kernel takes its exact-equality score path
    double ext_frac = 1.0 / 16;                     // share of an extension's worst-case matrix that gets written (adapted from the arena cursor)
    double t_prepare = 0, t_wait = 0, t_hand = 0, t_plan = 0;   // host phases of the back end (NDG_ALN_DEBUG)
    ~GpuDp() { for (auto &l : L) l.release(); }

    static uint64_t p_need(const ndal::DpReq &k)
    {
        int w = k.w < 0 ? std::max(k.qlen, k.tlen) : k.w;
        int ncol = std::min(k.qlen, k.tlen); ncol = (std::min(ncol, w + 1) + 15) / 16 + 1;
        const uint64_t rows = (uint64_t)(k.qlen + k.tlen - 1);
        return ((rows + 63) / 64 * 64) * (uint64_t)ncol * 16;                 // whole 64-row chunks
    }

    // host side of a launch: order the calls by the kernel that takes them (block-per-call and warp-per-call memory-resident kernels
    // first, then one group per ring size / maximum mode of the register-resident kernel), write the sequence bytes
    int prepare(DpLaunch &B, const std::vector<ndal::DpReq> &reqs, const ndal::SeqStore &T, const std::vector<uint32_t> &ids_in)
    {
        static const int coop_min = getenv("NDG_COOP_MIN") ? atoi(getenv("NDG_COOP_MIN")) : 1024;
        const bool use_rx = c->ksw_variant == 3;
        const size_t n = ids_in.size();
        // class of a call: 0 block-per-call, 1 warp-per-call, 2 + 2*cfg + exact register-resident
        std::vector<uint8_t> cls(n);
        size_t cnt[2 + 2 * RX_NCFG] = { 0 };
        for (size_t i = 0; i < n; ++i) {
            const ndal::DpReq &k = reqs[ids_in[i]];
            const int rc = use_rx ? rx_class(k.qlen, k.tlen, k.w, k.flag) : -1;
            if (rc >= 0) cls[i] = (uint8_t)(2 + 2 * rc + ((k.flag & NDK_APPROX_MAX) ? 0 : 1));
            else {
                const int wd = k.w < 0 ? std::max(k.qlen, k.tlen) : k.w;
                cls[i] = std::min(std::min(k.qlen, k.tlen), wd + 1) >= coop_min ? 0 : 1;
            }
            ++cnt[cls[i]];
        }
        size_t start[2 + 2 * RX_NCFG + 1]; start[0] = 0;
        for (int g = 0; g < 2 + 2 * RX_NCFG; ++g) start[g + 1] = start[g] + cnt[g];
        B.ids.assign(n, 0);
        { size_t fill[2 + 2 * RX_NCFG]; for (int g = 0; g < 2 + 2 * RX_NCFG; ++g) fill[g] = start[g];
          for (size_t i = 0; i < n; ++i) B.ids[fill[cls[i]]++] = ids_in[i]; }
        B.nb = cnt[0]; B.n_legacy = cnt[0] + cnt[1];
        B.rx.clear();
        for (int g = 2; g < 2 + 2 * RX_NCFG; ++g) if (cnt[g]) B.rx.push_back(DpLaunch::Grp{ (g - 2) / 2, ((g - 2) & 1) != 0, start[g], cnt[g] });
        B.n = n; B.hc.resize(n);
        uint64_t seqb = 0, cig = 0; B.Tmax = 16; B.rmax = 1; B.qmax = 1;
        for (size_t i = 0; i < n; ++i) {
            const ndal::DpReq &k = reqs[B.ids[i]]; ndk_call &d = B.hc[i];
            d.qlen = k.qlen; d.tlen = k.tlen; d.w = k.w; d.zdrop = k.zdrop; d.end_bonus = k.end_bonus; d.flag = k.flag;
            if (i < B.n_legacy) { d.q_off = seqb; d.t_off = seqb + (uint64_t)k.qlen; seqb += (uint64_t)k.qlen + k.tlen; }
            else {
                seqb = (seqb + 3) & ~3ull;
                unsigned long long qo, to; { uint64_t a, b2; rx_seq_offsets(seqb, k.qlen, &a, &b2); qo = a; to = b2; }
                d.q_off = qo; d.t_off = to; seqb += rx_seq_bytes(k.qlen, k.tlen);
                if (any_ambi) d.flag |= NDX_FLAG_AMBI;
            }
            d.p_off = 0; d.cigar_off = cig;
            cig += (uint64_t)k.qlen + k.tlen + 2;
            if (i < B.n_legacy) { B.Tmax = std::max(B.Tmax, (k.tlen + 15) / 16 * 16); B.qmax = std::max(B.qmax, k.qlen); }
            B.rmax = std::max(B.rmax, k.qlen + k.tlen);
        }
        B.seqb = seqb; B.cig_cap = cig;
        if (B.h_seq.reserve(seqb + 64)) return fail(c, NDG_ERR_NOMEM, "pinned sequence staging");
        uint8_t *hs = B.h_seq.as<uint8_t>();
        if (n > B.n_legacy) {                           // zero padding of the register-resident layout
            const uint64_t z0 = B.hc[B.n_legacy].q_off - 32;
#pragma omp parallel for schedule(static)
            for (long long blk = 0; blk < (long long)((seqb + 64 - z0 + (1 << 20) - 1) >> 20); ++blk) {
                const uint64_t b0 = z0 + ((uint64_t)blk << 20), b1 = std::min<uint64_t>(b0 + (1 << 20), seqb + 64);
                memset(hs + b0, 0, (size_t)(b1 - b0));
            }
        }
#pragma omp parallel for schedule(dynamic, 256)
        for (long long i = 0; i < (long long)n; ++i) ndal::materialize(reqs[B.ids[(size_t)i]], T, hs + B.hc[(size_t)i].q_off, hs + B.hc[(size_t)i].t_off);
        return 0;
    }

    // enqueue the input copies of a prepared launch (they do not touch the arena, so they overlap the previous launch's kernels)
    int submit_copies(DpLaunch &B)
    {
        const size_t n = B.n;
        if (!B.st) {
            CK(cudaStreamCreateWithFlags(&B.st, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&B.st2, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&B.ev_a, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&B.ev_b, cudaEventDisableTiming));
            CK(cudaEventCreate(&B.e0)); CK(cudaEventCreate(&B.e1));
        }
        RSV(B.calls, n * sizeof(ndk_call)); RSV(B.seqs, B.seqb + 128); RSV(B.res, n * sizeof(ndk_res));
        RSV(B.cig, (B.cig_cap + 4) * 4); RSV(B.cur, 256);
        RSV(B.ncig, (n + 1) * 4); RSV(B.coff, (n + 1) * 8);
        if (B.h_off.reserve((n + 2) * 8) || B.h_res.reserve(n * sizeof(ndk_res))) return fail(c, NDG_ERR_NOMEM, "pinned result staging");
        CK(cudaMemcpyAsync(B.calls.p, B.hc.data(), n * sizeof(ndk_call), cudaMemcpyHostToDevice, B.st));
        CK(cudaMemcpyAsync(B.seqs.p, B.h_seq.p, B.seqb + (B.n > B.n_legacy ? 64 : 0), cudaMemcpyHostToDevice, B.st));
        return 0;
    }

    // enqueue the kernels of a launch; the arena must be free (previous launch's kernels finished)
    int submit_kernels(DpLaunch &B, uint8_t *arena, uint64_t arena_bytes)
    {
        const size_t n = B.n, nb = B.nb, ns = B.n_legacy - nb;
        const size_t state_bytes = ((size_t)17 * B.Tmax + ((size_t)(B.qmax + 15) / 16 + 1) * 16 + 96 + 255) / 256 * 256;
        const size_t h_bytes = ((size_t)4 * B.Tmax + 255) / 256 * 256, off_bytes = ((size_t)8 * B.rmax + 255) / 256 * 256;
        const size_t tab_bytes = ((size_t)8 * ((size_t)(B.rmax >> 6) + 2) + 255) / 256 * 256;
        const size_t slot_bytes = state_bytes + h_bytes + off_bytes + tab_bytes;
        static const int bps = getenv("NDG_KSW_BPS") ? atoi(getenv("NDG_KSW_BPS")) : 16;
        const int blocks = ns ? std::min<int>(c->n_sm * bps, (int)((ns + NDK_WARPS - 1) / NDK_WARPS)) : 0;
        const int cblocks = nb ? (int)std::min<size_t>((size_t)c->n_sm * (16 / NDK_COOP_WARPS), nb) : 0;
        const size_t slots = (size_t)blocks * NDK_WARPS + (size_t)cblocks;
        RSV(c->ksw_scratch, slots * slot_bytes);    // shared: launches never run their kernels at the same time
        DevBuf &scr = c->ksw_scratch;
        unsigned long long *d_cur = B.cur.as<unsigned long long>();           // [0] arena cursor, [1] two work counters
        uint32_t *d_work = reinterpret_cast<uint32_t *>(d_cur + 1);
        CK(cudaMemsetAsync(d_cur, 0, 256, B.st));
        CK(cudaEventRecord(B.e0, B.st));
        if (nb) {                                   // second stream: both kernels share the SMs and the arena
            CK(cudaEventRecord(B.ev_a, B.st));
            CK(cudaStreamWaitEvent(B.st2, B.ev_a, 0));
            k_ksw_extd2_coop<<<cblocks, NDK_COOP_WARPS * 32, 0, B.st2>>>(B.calls.as<ndk_call>(), (int)nb, B.seqs.as<uint8_t>(), P, B.res.as<ndk_res>(),
                              arena, B.cig.as<uint32_t>(), scr.as<uint8_t>(), slot_bytes, state_bytes, h_bytes, d_work + 1, d_cur, arena_bytes, off_bytes);
            CK(cudaGetLastError());
            CK(cudaEventRecord(B.ev_b, B.st2));
            ++c->launches;
        }
        if (ns) {
            k_ksw_extd2_px<<<blocks, NDK_WARPS * 32, 0, B.st>>>(B.calls.as<ndk_call>() + nb, (int)ns, B.seqs.as<uint8_t>(), P, B.res.as<ndk_res>() + nb,
                              arena, B.cig.as<uint32_t>(), scr.as<uint8_t>() + (size_t)cblocks * slot_bytes, slot_bytes, state_bytes, h_bytes, d_work, d_cur, arena_bytes, off_bytes);
            CK(cudaGetLastError());
            ++c->launches;
        }
        if (!B.rx.empty()) {                        // register-resident kernel: one launch per ring size / maximum mode, all sharing the arena
            const size_t tab_entries = (size_t)(B.rmax >> 6) + 2;
            std::vector<int> nblk_of(B.rx.size()); size_t tot_blocks = 0;
            for (size_t g = 0; g < B.rx.size(); ++g) {
                int cap = 0;
                rx_launch(B.rx[g].cfg, B.rx[g].exact, 0, c->n_sm, B.st, nullptr, 0, nullptr, P, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, 0, &cap);
                nblk_of[g] = (int)std::min<size_t>((size_t)cap, B.rx[g].count); tot_blocks += (size_t)nblk_of[g];
            }
            RSV(B.rxtab, tot_blocks * tab_entries * 8 + 64);
            size_t tb = 0;
            for (size_t g = 0; g < B.rx.size(); ++g) {
                const DpLaunch::Grp &G = B.rx[g];
                CK(rx_launch(G.cfg, G.exact, 0, c->n_sm, B.st, B.calls.as<ndk_call>() + G.begin, (int)G.count, B.seqs.as<uint8_t>(), P, B.res.as<ndk_res>() + G.begin, arena,
                             B.cig.as<uint32_t>(), B.rxtab.as<unsigned long long>() + tb * tab_entries, tab_entries, d_work + 2 + g, d_cur, arena_bytes));
                tb += (size_t)nblk_of[g];
                ++c->launches;
            }
        }
        if (nb) CK(cudaStreamWaitEvent(B.st, B.ev_b, 0));
        CK(cudaEventRecord(B.e1, B.st));
        k_ksw_ncigar<<<nblk(n + 1, 256), 256, 0, B.st>>>(B.res.as<ndk_res>(), (int)n, B.ncig.as<uint32_t>());
        {
            size_t tmp = 0;
            auto it = cub::TransformInputIterator<nd_u64, cub::CastOp<nd_u64>, const uint32_t *>(B.ncig.as<uint32_t>(), cub::CastOp<nd_u64>());
            CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, it, B.coff.as<nd_u64>(), (long long)(n + 1), B.st));
            RSV(B.tmp, tmp);
            CK(cub::DeviceScan::ExclusiveSum(B.tmp.p, tmp, it, B.coff.as<nd_u64>(), (long long)(n + 1), B.st));
        }
        CK(cudaMemcpyAsync(B.h_off.p, B.coff.p, (n + 1) * 8, cudaMemcpyDeviceToHost, B.st));
        CK(cudaMemcpyAsync(B.h_off.as<uint64_t>() + n + 1, d_cur, 8, cudaMemcpyDeviceToHost, B.st));
        CK(cudaMemcpyAsync(B.h_res.p, B.res.p, n * sizeof(ndk_res), cudaMemcpyDeviceToHost, B.st));
        B.busy = true;
        return 0;
    }

    // wait for the kernels of a launch: calls that found the arena full go to retry; the CIGAR gather is queued behind
    int collect_a(DpLaunch &B, std::vector<uint32_t> &retry, size_t *n_retry_out)
    {
        const size_t n = B.n;
        CK(cudaStreamSynchronize(B.st));
        { float ms = 0; cudaEventElapsedTime(&ms, B.e0, B.e1); c->aln_stat.dp_ms += ms;
          if (getenv("NDG_ALN_DEBUG")) fprintf(stderr, "[ndg::dp] launch %llu: %zu calls (%zu wide), first %dx%d w %d flag 0x%x, seq %.1f MB, kernels %.1f ms, arena used %.2f GB, ext_frac %.4f\n",
                                               (unsigned long long)c->aln_stat.n_launches, n, B.nb, B.hc[0].qlen, B.hc[0].tlen, B.hc[0].w, B.hc[0].flag, B.seqb / 1e6, ms, B.h_off.as<uint64_t>()[n + 1] / 1e9, ext_frac); }
        const uint64_t *koff = B.h_off.as<uint64_t>(); const uint64_t tot = koff[n], used = koff[n + 1];
        const ndk_res *hr = B.h_res.as<ndk_res>();
        size_t n_retry = 0;
        for (size_t i = 0; i < n; ++i) if (hr[i].pad) { retry.push_back(B.ids[i]); ++n_retry; }
        if (B.ext_need > 0 && n_retry == 0) {       // what the optimistic calls really wrote
            const double u = used > B.fixed ? (double)(used - B.fixed) : 0.0;
            ext_frac = std::min(1.0, std::max(1.0 / 256, 0.5 * ext_frac + 0.5 * 1.5 * u / (double)B.ext_need));
        } else if (n_retry) ext_frac = std::min(1.0, ext_frac * 2);
        c->aln_stat.n_retries += n_retry; *n_retry_out = n_retry; c->aln_stat.p_bytes += std::min<uint64_t>(used, c->ksw_p.cap);
        if (tot) {
            RSV(B.cout, tot * 4);
            if (B.h_cig.reserve(tot * 4)) return fail(c, NDG_ERR_NOMEM, "pinned cigar staging");
            k_ksw_cigar_gather<<<nblk((uint64_t)n * 32, 256), 256, 0, B.st>>>(B.calls.as<ndk_call>(), B.res.as<ndk_res>(), (const unsigned long long *)B.coff.as<nd_u64>(), (int)n,
                                                                                 B.cig.as<uint32_t>(), B.cout.as<uint32_t>());
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(B.h_cig.p, B.cout.p, tot * 4, cudaMemcpyDeviceToHost, B.st));
        }
        return 0;
    }

    // CIGARs of a launch have arrived: hand the results over (overlaps the next launch's kernels)
    int collect_b(DpLaunch &B, std::vector<ndal::DpRes> &res, std::vector<uint32_t> &cigars)
    {
        const size_t n = B.n;
        CK(cudaStreamSynchronize(B.st));
        B.busy = false;
        const uint64_t *koff = B.h_off.as<uint64_t>(); const uint64_t tot = koff[n];
        const ndk_res *hr = B.h_res.as<ndk_res>();
        const size_t base = cigars.size();
        if (base + tot >= (1ull << 32)) return fail(c, NDG_ERR_OVERFLOW, "round produces more than 2^32 CIGAR operations; align fewer reads per batch");
        cigars.resize(base + tot);
        if (tot) memcpy(cigars.data() + base, B.h_cig.p, tot * 4);
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)n; ++i) {
            if (hr[i].pad) continue;
            ndal::DpRes &z = res[B.ids[(size_t)i]];
            z.max = hr[i].max; z.zdropped = hr[i].zdropped; z.max_q = hr[i].max_q; z.max_t = hr[i].max_t; z.mqe = hr[i].mqe; z.mqe_t = hr[i].mqe_t;
            z.mte = hr[i].mte; z.mte_q = hr[i].mte_q; z.score = hr[i].score; z.n_cigar = hr[i].n_cigar; z.reach_end = hr[i].reach_end;
            z.cig_off = (uint32_t)(base + koff[i]);
        }
        ++c->aln_stat.n_launches;
        return 0;
    }

    int run_impl(const std::vector<ndal::DpReq> &reqs, const ndal::SeqStore &T, std::vector<ndal::DpRes> &res, std::vector<uint32_t> &cigars)
    {
        const size_t n = reqs.size();
        cudaSetDevice(c->device);
        {                                           // backtrack arena: a fixed slice of HBM, reused by every launch
            if (!c->arena_want && getenv("NDG_ARENA_GB")) c->arena_want = (uint64_t)(atof(getenv("NDG_ARENA_GB")) * (1ull << 30));
            size_t fr = 0, tot = 0; cudaMemGetInfo(&fr, &tot);
            const uint64_t want = c->arena_want ? c->arena_want : std::min<uint64_t>(96ull << 30, (uint64_t)((fr + c->ksw_p.cap) * 0.45));
            if (c->ksw_p.cap < want - want / 8) {   // also replaces a smaller buffer left by ndg_ksw_extd2_batch()
                c->ksw_p.release();
                if (cudaMalloc(&c->ksw_p.p, want) != cudaSuccess) { cudaGetLastError(); c->ksw_p.p = nullptr; return fail(c, NDG_ERR_NOMEM, "backtrack arena of %llu bytes", (unsigned long long)want); }
                c->ksw_p.cap = want;
            }
        }
        const double t_plan0 = omp_get_wtime();
        const uint64_t arena = c->ksw_p.cap & ~(uint64_t)255, arena_eff = arena - arena / 16;
        std::vector<uint64_t> need(n);
        std::vector<uint32_t> pending(n);
        for (size_t i = 0; i < n; ++i) { need[i] = p_need(reqs[i]); pending[i] = (uint32_t)i; if (need[i] > arena_eff) return fail(c, NDG_ERR_NOMEM, "one alignment needs a %llu-byte backtrack matrix, the arena is %llu", (unsigned long long)need[i], (unsigned long long)arena); }
        __gnu_parallel::sort(pending.begin(), pending.end(), [&](uint32_t x, uint32_t y) { return need[x] != need[y] ? need[x] > need[y] : x < y; });   // long calls first
        std::vector<uint8_t> full(n, 0);            // 1: reserve the worst case (after a failed optimistic attempt)
        std::vector<uint32_t> group, retry;
        const uint64_t SEQ_MAX = 768ull << 20; const size_t CNT_MAX = 4u << 20;
        auto est_of = [&](uint32_t id) -> uint64_t {
            // an extension usually z-drops long before its worst-case matrix is written: reserve a fraction first
            if (full[id] || !(reqs[id].flag & NDK_EXTZ_ONLY)) return need[id];
            return std::min<uint64_t>(need[id], (1ull << 20) + (uint64_t)(need[id] * ext_frac));
        };
        t_plan += omp_get_wtime() - t_plan0;
        size_t head = 0, rhead = 0; int cur = -1, nxt = 0;
        // A launch takes a run of the size-ordered list: calls of similar length share a launch, so no launch waits on one call that is
        // far longer than the rest.  While the kernels of one launch run, the next one is staged (sequence bytes, call table, copies)
        // and the previous one's results are handed over; kernels of different launches never overlap (they share the arena).
        for (;;) {
            const bool more = head < pending.size() || rhead < retry.size();
            if (more) {
                DpLaunch &B = L[nxt];
                group.clear();
                uint64_t sum = 0, seqb = 0; B.fixed = 0; B.ext_need = 0;
                auto take = [&](uint32_t id) -> bool {
                    const uint64_t e = est_of(id), sb = (uint64_t)reqs[id].qlen + reqs[id].tlen + 2;
                    if (!group.empty() && (sum + e > arena_eff || seqb + sb > SEQ_MAX || group.size() >= CNT_MAX)) return false;
                    group.push_back(id); sum += e; seqb += sb;
                    if (e == need[id]) B.fixed += e; else B.ext_need += need[id];
                    return true;
                };
                while (rhead < retry.size() && take(retry[rhead])) ++rhead;           // retried calls first (they are long)
                if (rhead >= retry.size()) while (head < pending.size() && take(pending[head])) ++head;
                const double tp = omp_get_wtime();
                if (int r = prepare(B, reqs, T, group)) return r;
                if (int r = submit_copies(B)) return r;
                t_prepare += omp_get_wtime() - tp;
            }
            if (cur >= 0) {
                const size_t r0 = retry.size(); size_t nr = 0;
                const double tw = omp_get_wtime();
                if (int r = collect_a(L[cur], retry, &nr)) return r;
                t_wait += omp_get_wtime() - tw;
                for (size_t i = r0; i < retry.size(); ++i) { if (full[retry[i]] && L[cur].n == 1) return fail(c, NDG_ERR_NOMEM, "backtrack arena exhausted by a single call"); full[retry[i]] = 1; }
            }
            if (more) { if (int r = submit_kernels(L[nxt], c->ksw_p.as<uint8_t>(), arena)) return r; L[nxt].busy = true; }
            if (cur >= 0) { const double th = omp_get_wtime(); if (int r = collect_b(L[cur], res, cigars)) return r; t_hand += omp_get_wtime() - th; }
            if (more) { cur = nxt; nxt ^= 1; }
            else if (cur >= 0) cur = -1;            // the last launch is in; retried calls may have appeared
            else break;
        }
        return 0;
    }

    void run(const std::vector<ndal::DpReq> &reqs, const ndal::SeqStore &T, std::vector<ndal::DpRes> &res, std::vector<uint32_t> &cigars) override
    {
        res.assign(reqs.size(), ndal::DpRes()); cigars.clear();
        if (rc) return;                             // after an error the tasks drain on default results; the caller reports rc
        const double t0 = omp_get_wtime();
        rc = run_impl(reqs, T, res, cigars);
        if (rc) for (auto &l : L) if (l.busy) { cudaStreamSynchronize(l.st); l.busy = false; }
        c->aln_stat.dp_wall_ms += (omp_get_wtime() - t0) * 1e3;
    }
};

// chains of the resident sub-batch -> host, alignment, final hits appended to c->aln_hits
int align_resident(ndg_ctx *c, const ndg_reads_t *Q, int r0, int r1)
{
    const uint64_t NH = c->n_hits;
    if (NH == 0) return 0;
    const double t_host0 = omp_get_wtime();
    omp_set_num_threads(c->n_threads);
    uint64_t *h = c->h_scalars.as<uint64_t>();
    { ndal::Opt &A = c->aln; const ndg_opt_t &o = c->opt; A.k = o.k; A.hpc = o.hpc; A.bw = o.bw; A.max_gap = o.max_gap; A.min_cnt = o.min_cnt; A.min_chain_score = o.min_chain_score; }
    uint64_t NCA = 0;
    if (int rc = collect_chains(c, r0, &NCA)) return rc;
    if (c->h_hits.reserve((NH + 1) * sizeof(nd_hit)) || c->h_cha.reserve((NCA + 1) * 16) || c->h_chf.reserve((NH + 1) * 4)) return fail(c, NDG_ERR_NOMEM, "pinned chain staging");
    CK(cudaMemcpyAsync(c->h_hits.p, c->hits_sorted.p, NH * sizeof(nd_hit), cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(c->h_cha.p, c->ch_a.p, NCA * 16, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(c->h_chf.p, c->ch_fend.p, NH * 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    const nd_hit *hh = c->h_hits.as<nd_hit>(); const nd_u128 *ha = c->h_cha.as<nd_u128>(); const int32_t *hf = c->h_chf.as<int32_t>();

    ndal::SeqStore ST; ST.words = c->ht->words.data(); ST.word_off = c->ht->word_off.data(); ST.len = c->h_tlen.data(); ST.nmask = c->ht->has_n ? c->ht->nmask.data() : nullptr;
    ndal::SeqStore SQ; SQ.words = Q->words; SQ.word_off = Q->word_off; SQ.len = Q->len; SQ.nmask = Q->nmask;
    GpuDp be; be.c = c; be.any_ambi = ST.nmask != nullptr || SQ.nmask != nullptr;
    be.P.m = 5; memset(be.P.mat, 0, sizeof be.P.mat); ndal::simple_mat(be.P.mat, c->aln); be.P.q = c->aln.q; be.P.e = c->aln.e; be.P.q2 = c->aln.q2; be.P.e2 = c->aln.e2;

    // reads are aligned in groups that bound the host memory held by one round of DP results
    const uint64_t GROUP_CHAINS = 400000;
    uint64_t hi = 0, aoff = 0;
    int r = r0;
    while (r < r1) {
        const int g0 = r;
        uint64_t nch = 0;
        std::vector<uint64_t> hbeg, abeg;
        while (r < r1 && (r == g0 || nch < GROUP_CHAINS)) {
            hbeg.push_back(hi); abeg.push_back(aoff);
            while (hi < NH && hh[hi].qidx == r) { aoff += (uint64_t)hh[hi].cnt; ++hi; ++nch; }
            ++r;
        }
        hbeg.push_back(hi); abeg.push_back(aoff);
        std::vector<ndal::ReadJob> jobs((size_t)(r - g0));
#pragma omp parallel for schedule(dynamic, 1)
        for (int j = 0; j < r - g0; ++j) {
            ndal::ReadJob &J = jobs[(size_t)j];
            const int qi = g0 + j;
            J.qid = qi; J.qlen = (int)Q->len[qi];
            const uint64_t b = hbeg[(size_t)j], e = hbeg[(size_t)j + 1];
            if (b == e) continue;
            J.strands.resize((size_t)2 * J.qlen);
            ndal::encode_strands(SQ, qi, J.qlen, J.strands.data(), J.strands.data() + J.qlen);
            std::vector<ndal::ChainIn> ch((size_t)(e - b));
            uint64_t ao = abeg[(size_t)j];
            for (uint64_t k = b; k < e; ++k) {
                ndal::ChainIn &x = ch[(size_t)(k - b)];
                x.zkey = hh[k].zkey; x.score = hh[k].score; x.cnt = hh[k].cnt; x.fend = hf[k]; x.jend = hh[k].as; x.anchors = ha + ao;
                ao += (uint64_t)hh[k].cnt;
            }
            ndal::build_read_job(J, ch);
        }
        ndal::Stats st;
        const double tb0 = omp_get_wtime();
        ndal::align_batch(c->aln, ST, jobs, be, &st);
        if (be.rc) return be.rc;
        if (getenv("NDG_ALN_DEBUG")) fprintf(stderr, "[ndg::dp] back end host phases: order+estimate %.2f s, prepare+copies %.2f s, waiting for kernels %.2f s, hand-back %.2f s\n", be.t_plan, be.t_prepare, be.t_wait, be.t_hand);
        if (getenv("NDG_ALN_DEBUG")) fprintf(stderr, "[ndg::aln] group of %zu reads: prepare %.2f s, walk %.2f s, dp (host wall) %.2f s, hand-back %.2f s, finish %.2f s, total %.2f s\n", jobs.size(),
                                             tb0 - t_host0, st.t_walk, st.t_dp, st.t_hand, st.t_finish, omp_get_wtime() - tb0);
        c->aln_stat.n_chains += st.n_tasks; c->aln_stat.n_calls += st.n_calls; c->aln_stat.n_rounds += st.n_rounds; c->aln_stat.rect_cells += st.cells;
        for (size_t j = 0; j < jobs.size(); ++j)
            for (const ndal::Reg &x : jobs[j].regs) {
                ndg_hit_t o; memset(&o, 0, sizeof o);
                o.zkey = (uint64_t)(uint32_t)(x.has_p ? x.dp_max : x.score) << 32 | x.hash;
                o.qidx = jobs[j].qid; o.rid = x.rid; o.qs = x.qs; o.qe = x.qe; o.rs = x.rs; o.re = x.re; o.mlen = x.mlen; o.blen = x.blen;
                o.score = x.has_p ? x.dp_max : x.score; o.cnt = x.cnt; o.rev = x.rev; o.as = x.as;
                c->aln_hits.push_back(o);
                if (c->paf_on) {
                    ndg_hit_extra_t e; memset(&e, 0, sizeof e);
                    e.n_span = -1; e.dp_score = x.dp_score; e.dp_max = x.dp_max; e.n_ambi = x.n_ambi; e.score0 = x.score0; e.chain_score = x.score;
                    e.flags = (x.inv ? 1u : 0u) | (x.id == x.parent ? 2u : 0u) | ((uint32_t)(x.split & 3) << 2) | (x.has_p ? 16u : 0u);
                    e.n_cigar = (uint32_t)x.cigar.size(); e.cigar_off = c->paf_cig.size();
                    c->paf_cig.insert(c->paf_cig.end(), x.cigar.begin(), x.cigar.end());
                    c->paf_he.push_back(e);
                }
            }
    }
    c->aln_stat.host_ms += (omp_get_wtime() - t_host0) * 1e3;
    return 0;
}

} // namespace

// ------------------------------------------------------------------------------------------------
// ovl_sort on the GPU (nd_ovlsort_gpu.cuh, nd_ovlsort.cuh).  ref: util/ovl_sort.c
// ------------------------------------------------------------------------------------------------
namespace {
struct DevScope {                                  // transient device buffers of one ndg_ovl_sort call
    std::vector<DevBuf *> all;
    DevBuf *get() { DevBuf *b = new DevBuf(); all.push_back(b); return b; }
    ~DevScope() { for (auto b : all) { b->release(); delete b; } }
};
}

extern "C" int ndg_ovl_sort(ndg_ctx_t *c, const uint8_t *const *streams, const uint64_t *stream_bytes, int32_t n_streams,
                            const uint32_t *seed_name, const uint32_t *seed_len, uint64_t n_seeds, const ndg_ovlsort_opt_t *opt,
                            const uint8_t **ovl, uint64_t *n_bytes, const char **bl_text, uint64_t *bl_bytes, uint64_t *n_rec_in, uint64_t *n_rec_out)
{
    if (!c || !streams || !stream_bytes || n_streams < 0 || !seed_name || !seed_len || !opt || !ovl || !n_bytes) return NDG_ERR_ARG;
    cudaSetDevice(c->device);
    RSV(c->scalars, 1024);
    uint64_t *h = c->h_scalars.as<uint64_t>();
    DevScope S;
    // seeds: sorted by name for the device lookups; the shortest seed feeds the clipping rule (ref: util/ovl_sort.c:106-133,538)
    ndos_par P; P.max_bin_cov = opt->max_bin_cov; P.max_flank_len = opt->max_flank_len; P.is_hq = opt->is_hq; P.with_bl = opt->max_flank_len != 0; P.min_seed_len = 0;
    std::vector<std::pair<uint32_t, uint32_t>> sd(n_seeds);
    for (uint64_t i = 0; i < n_seeds; ++i) sd[i] = { seed_name[i], seed_len[i] };
    std::stable_sort(sd.begin(), sd.end(), [](const std::pair<uint32_t, uint32_t> &a, const std::pair<uint32_t, uint32_t> &b) { return a.first < b.first; });
    { size_t k = 0; for (size_t i = 0; i < sd.size(); ++i) if (i == 0 || sd[i].first != sd[i - 1].first) sd[k++] = sd[i]; sd.resize(k); }   // the first entry of a name wins (kh_put absent test)
    {   // min_seed_len follows the file order of the first occurrences (:126); the minimum over distinct names is the same number
        for (auto &x : sd) if ((int)x.second < P.min_seed_len || !P.min_seed_len) P.min_seed_len = (int)x.second;
    }
    const uint64_t NS = sd.size();
    std::vector<uint32_t> hn(NS), hl(NS);
    for (uint64_t i = 0; i < NS; ++i) { hn[i] = sd[i].first; hl[i] = sd[i].second; }
    DevBuf &d_sn = *S.get(), &d_sl = *S.get();
    RSV(d_sn, (NS + 1) * 4); RSV(d_sl, (NS + 1) * 4);
    if (NS) { CK(cudaMemcpyAsync(d_sn.p, hn.data(), NS * 4, cudaMemcpyHostToDevice, c->st)); CK(cudaMemcpyAsync(d_sl.p, hl.data(), NS * 4, cudaMemcpyHostToDevice, c->st)); }

    // ---- 1. decode + reader filter, stream by stream; the qualifying records are appended in stream order ----------------------------------
    DevBuf &R = *S.get();
    uint64_t NR = 0, n_in = 0;
    DevBuf &d_b = *S.get(), &flag = *S.get(), &vidx = *S.get(), &vals = *S.get(), &dq = *S.get(), &dt = *S.get(), &qn = *S.get(), &tn = *S.get(), &ok = *S.get(),
           &f1 = *S.get(), &f5 = *S.get(), &nf1 = *S.get(), &nf5 = *S.get(), &keep = *S.get(), &pos = *S.get();
    for (int s = 0; s < n_streams; ++s) {
        const uint64_t nb = stream_bytes[s];
        if (!nb) continue;
        RSV(d_b, nb + 16); RSV(flag, (nb + 1) * 4); RSV(vidx, (nb + 1) * 8);
        CK(cudaMemcpyAsync(d_b.p, streams[s], nb, cudaMemcpyHostToDevice, c->st));
        KLAUNCH("ovl_sort", (k_os_term<<<nblk(nb + 1, 256), 256, 0, c->st>>>(d_b.as<uint8_t>(), nb, flag.as<uint32_t>())));
        if (int rc = scan_u32_to_u64(c, flag.as<uint32_t>(), vidx.as<nd_u64>(), nb)) return rc;
        CK(cudaMemcpyAsync(h, vidx.as<nd_u64>() + nb, 8, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        const uint64_t nrec = h[0] / 8;
        if (!nrec) continue;
        n_in += nrec;
        RSV(vals, (h[0] + 8) * 4); RSV(dq, (nrec + 1) * 4); RSV(dt, (nrec + 1) * 4); RSV(qn, (nrec + 1) * 4); RSV(tn, (nrec + 1) * 4);
        RSV(ok, nrec + 16); RSV(f1, (nrec + 1) * 4); RSV(f5, (nrec + 1) * 4); RSV(nf1, (nrec + 1) * 8); RSV(nf5, (nrec + 1) * 8); RSV(keep, (2 * nrec + 2) * 4); RSV(pos, (2 * nrec + 2) * 8);
        KLAUNCH("ovl_sort", (k_os_values<<<nblk(nb, 256), 256, 0, c->st>>>(d_b.as<uint8_t>(), nb, vidx.as<nd_u64>(), vals.as<uint32_t>())));
        KLAUNCH("ovl_sort", (k_os_deltas<<<nblk(nrec, 256), 256, 0, c->st>>>(vals.as<uint32_t>(), nrec, dq.as<uint32_t>(), dt.as<uint32_t>())));
        {
            size_t tmp = 0;
            CK(cub::DeviceScan::InclusiveSum(nullptr, tmp, dq.as<uint32_t>(), qn.as<uint32_t>(), (long long)nrec, c->st));
            RSV(c->cubtmp, tmp);
            KLAUNCH("ovl_sort", CK(cub::DeviceScan::InclusiveSum(c->cubtmp.p, tmp, dq.as<uint32_t>(), qn.as<uint32_t>(), (long long)nrec, c->st)));
            KLAUNCH("ovl_sort", CK(cub::DeviceScan::InclusiveSum(c->cubtmp.p, tmp, dt.as<uint32_t>(), tn.as<uint32_t>(), (long long)nrec, c->st)));
        }
        KLAUNCH("ovl_sort", (k_os_cand<<<nblk(nrec + 1, 256), 256, 0, c->st>>>(vals.as<uint32_t>(), qn.as<uint32_t>(), tn.as<uint32_t>(), nrec, d_sn.as<uint32_t>(), d_sl.as<uint32_t>(), NS,
                             ok.as<uint8_t>(), f1.as<uint32_t>(), f5.as<uint32_t>())));
        if (int rc = scan_u32_to_u64(c, f1.as<uint32_t>(), nf1.as<nd_u64>(), nrec)) return rc;
        if (int rc = scan_u32_to_u64(c, f5.as<uint32_t>(), nf5.as<nd_u64>(), nrec)) return rc;
        KLAUNCH("ovl_sort", (k_os_keep<<<nblk(nrec + 1, 256), 256, 0, c->st>>>(ok.as<uint8_t>(), nf1.as<nd_u64>(), nf5.as<nd_u64>(), nrec, keep.as<uint32_t>())));
        CK(cudaMemsetAsync(keep.as<uint32_t>() + 2 * nrec, 0, 8, c->st));
        if (int rc = scan_u32_to_u64(c, keep.as<uint32_t>(), pos.as<nd_u64>(), 2 * nrec)) return rc;
        CK(cudaMemcpyAsync(h, pos.as<nd_u64>() + 2 * nrec, 8, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        const uint64_t add = h[0];
        if (R.grow_keep((NR + add + 1) * sizeof(ndos_rec), NR * sizeof(ndos_rec), c->st)) return fail(c, NDG_ERR_NOMEM, "ovl_sort records");
        KLAUNCH("ovl_sort", (k_os_scatter<<<nblk(nrec, 256), 256, 0, c->st>>>(vals.as<uint32_t>(), qn.as<uint32_t>(), tn.as<uint32_t>(), nrec, keep.as<uint32_t>(), pos.as<nd_u64>(), NR, R.as<ndos_rec>())));
        CK(cudaGetLastError());
        NR += add;
    }
    if (n_rec_in) *n_rec_in = n_in;
    c->os_bl.clear();
    *ovl = nullptr; *n_bytes = 0;
    if (bl_text) *bl_text = c->os_bl.c_str();
    if (bl_bytes) *bl_bytes = 0;
    if (n_rec_out) *n_rec_out = 0;
    if (NR == 0) { CK(cudaStreamSynchronize(c->st)); return NDG_OK; }
    if (NR >= (1ULL << 32) - 16) return fail(c, NDG_ERR_OVERFLOW, "ovl_sort: %llu records in one call", (unsigned long long)NR);

    // ---- 2. order: (qname, match descending, span ascending), ties in input order (two stable radix passes on an index) --------------------
    DevBuf &k32 = *S.get(), &k32b = *S.get(), &ix = *S.get(), &ixb = *S.get(), &k64 = *S.get(), &k64b = *S.get();
    RSV(k32, NR * 4); RSV(k32b, NR * 4); RSV(ix, NR * 4); RSV(ixb, NR * 4); RSV(k64, NR * 8); RSV(k64b, NR * 8);
    KLAUNCH("ovl_sort", (k_os_key_span<<<nblk(NR, 256), 256, 0, c->st>>>(R.as<ndos_rec>(), NR, k32.as<uint32_t>(), ix.as<uint32_t>())));
    {
        size_t t1 = 0, t2 = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, t1, k32.as<uint32_t>(), k32b.as<uint32_t>(), ix.as<uint32_t>(), ixb.as<uint32_t>(), (long long)NR, 0, 32, c->st));
        CK(cub::DeviceRadixSort::SortPairs(nullptr, t2, k64.as<nd_u64>(), k64b.as<nd_u64>(), ixb.as<uint32_t>(), ix.as<uint32_t>(), (long long)NR, 0, 64, c->st));
        RSV(c->cubtmp, std::max(t1, t2));
        KLAUNCH("ovl_sort", CK(cub::DeviceRadixSort::SortPairs(c->cubtmp.p, t1, k32.as<uint32_t>(), k32b.as<uint32_t>(), ix.as<uint32_t>(), ixb.as<uint32_t>(), (long long)NR, 0, 32, c->st)));
        KLAUNCH("ovl_sort", (k_os_key_main<<<nblk(NR, 256), 256, 0, c->st>>>(R.as<ndos_rec>(), ixb.as<uint32_t>(), NR, k64.as<nd_u64>())));
        KLAUNCH("ovl_sort", CK(cub::DeviceRadixSort::SortPairs(c->cubtmp.p, t2, k64.as<nd_u64>(), k64b.as<nd_u64>(), ixb.as<uint32_t>(), ix.as<uint32_t>(), (long long)NR, 0, 64, c->st)));
    }
    // ---- 3. reads: boundaries, the read's own record in front, per-read scratch ---------------------------------------------------------
    DevBuf &start = *S.get(), &nstart = *S.get();
    RSV(start, (NR + 1) * 4); RSV(nstart, (NR + 1) * 8);
    KLAUNCH("ovl_sort", (k_os_starts<<<nblk(NR + 1, 256), 256, 0, c->st>>>(R.as<ndos_rec>(), ix.as<uint32_t>(), NR, start.as<uint32_t>())));
    if (int rc = scan_u32_to_u64(c, start.as<uint32_t>(), nstart.as<nd_u64>(), NR)) return rc;
    CK(cudaMemcpyAsync(h, nstart.as<nd_u64>() + NR, 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    const uint64_t NQ = h[0], NT = NR + NQ;
    DevBuf &L = *S.get(), &q_off = *S.get(), &q_len = *S.get(), &binl = *S.get(), &bin_off = *S.get(), &acc = *S.get(), &dead = *S.get(), &label = *S.get(), &qname = *S.get();
    RSV(L, (NT + 1) * sizeof(ndos_rec)); RSV(q_off, (NQ + 1) * 8); RSV(q_len, (NQ + 1) * 4); RSV(binl, (NQ + 1) * 4); RSV(bin_off, (NQ + 1) * 8);
    RSV(acc, (NT + 1) * 4); RSV(dead, NT + 16); RSV(label, NQ + 16); RSV(qname, (NQ + 1) * 4);
    KLAUNCH("ovl_sort", (k_os_layout<<<nblk(NR, 256), 256, 0, c->st>>>(R.as<ndos_rec>(), ix.as<uint32_t>(), start.as<uint32_t>(), nstart.as<nd_u64>(), NR, d_sn.as<uint32_t>(), d_sl.as<uint32_t>(), NS,
                         L.as<ndos_rec>(), q_off.as<nd_u64>(), q_len.as<uint32_t>())));
    KLAUNCH("ovl_sort", (k_os_binlen<<<nblk(NQ + 1, 256), 256, 0, c->st>>>(q_len.as<uint32_t>(), NQ, binl.as<uint32_t>())));
    if (int rc = scan_u32_to_u64(c, binl.as<uint32_t>(), bin_off.as<nd_u64>(), NQ)) return rc;
    CK(cudaMemcpyAsync(h, bin_off.as<nd_u64>() + NQ, 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    DevBuf &bins = *S.get();
    RSV(bins, (h[0] + 16) * 2);
    // ---- 4. the per-read filter ----------------------------------------------------------------------------------------------------------
    KLAUNCH("ovl_sort_reads", (k_os_reads<<<nblk(NQ, 64), 64, 0, c->st>>>(P, L.as<ndos_rec>(), q_off.as<nd_u64>(), q_len.as<uint32_t>(), NQ, NT, bin_off.as<nd_u64>(), bins.as<uint16_t>(),
                               acc.as<uint32_t>(), dead.as<uint8_t>(), label.as<char>())));
    CK(cudaGetLastError());
    // ---- 5. survivors -> 8-field stream ----------------------------------------------------------------------------------------------------
    DevBuf &live = *S.get(), &kidx = *S.get(), &len = *S.get(), &off = *S.get(), &outb = *S.get();
    RSV(live, NT + 16); RSV(kidx, (NT + 1) * 4);
    KLAUNCH("ovl_sort", (k_os_live<<<nblk(NT, 256), 256, 0, c->st>>>(L.as<ndos_rec>(), dead.as<uint8_t>(), NT, live.as<uint8_t>())));
    uint32_t *d_nk = c->scalars.as<uint32_t>() + 200;
    {
        size_t tmp = 0;
        cub::CountingInputIterator<uint32_t> it(0);
        CK(cub::DeviceSelect::Flagged(nullptr, tmp, it, live.as<uint8_t>(), kidx.as<uint32_t>(), d_nk, (long long)NT, c->st));
        RSV(c->cubtmp, tmp);
        KLAUNCH("ovl_sort", CK(cub::DeviceSelect::Flagged(c->cubtmp.p, tmp, it, live.as<uint8_t>(), kidx.as<uint32_t>(), d_nk, (long long)NT, c->st)));
    }
    CK(cudaMemcpyAsync(h, d_nk, 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    const uint64_t NK = (uint32_t)h[0];
    uint64_t out_bytes = 0;
    if (NK) {
        RSV(len, (NK + 1) * 4); RSV(off, (NK + 1) * 8);
        KLAUNCH("ovl_sort", (k_os_len<<<nblk(NK + 1, 256), 256, 0, c->st>>>(L.as<ndos_rec>(), kidx.as<uint32_t>(), NK, len.as<uint32_t>())));
        if (int rc = scan_u32_to_u64(c, len.as<uint32_t>(), off.as<nd_u64>(), NK)) return rc;
        CK(cudaMemcpyAsync(h, off.as<nd_u64>() + NK, 8, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        out_bytes = h[0];
        RSV(outb, out_bytes + 64);
        KLAUNCH("ovl_sort", (k_os_write<<<nblk(NK, 256), 256, 0, c->st>>>(L.as<ndos_rec>(), kidx.as<uint32_t>(), NK, off.as<nd_u64>(), outb.as<uint8_t>())));
        CK(cudaGetLastError());
        if (c->os_out.reserve(out_bytes + 64)) return fail(c, NDG_ERR_NOMEM, "pinned ovl_sort output");
        CK(cudaMemcpyAsync(c->os_out.p, outb.p, out_bytes, cudaMemcpyDeviceToHost, c->st));
    }
    // labels -> .bl text (ref: util/ovl_sort.c:384-387: "<qname> c" contained, "<qname> k" chimeric)
    KLAUNCH("ovl_sort", (k_os_qnames<<<nblk(NQ, 256), 256, 0, c->st>>>(L.as<ndos_rec>(), q_off.as<nd_u64>(), NQ, qname.as<uint32_t>())));
    std::vector<char> hlab(NQ); std::vector<uint32_t> hqn(NQ);
    CK(cudaMemcpyAsync(hlab.data(), label.p, NQ, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(hqn.data(), qname.p, NQ * 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (P.with_bl) { char line[32]; for (uint64_t q = 0; q < NQ; ++q) if (hlab[q]) { snprintf(line, sizeof line, "%u %c\n", hqn[q], hlab[q]); c->os_bl += line; } }
    *ovl = c->os_out.as<uint8_t>(); *n_bytes = out_bytes;
    if (bl_text) *bl_text = c->os_bl.c_str();
    if (bl_bytes) *bl_bytes = c->os_bl.size();
    if (n_rec_out) *n_rec_out = NK;
    return NDG_OK;
}

extern "C" int ndg_set_paf(ndg_ctx_t *c, int on)
{
    if (!c) return NDG_ERR_ARG;
    c->paf_on = on != 0;
    return NDG_OK;
}

extern "C" int ndg_map_extras(ndg_ctx_t *c, const ndg_hit_extra_t **he, const ndg_query_extra_t **qe, const uint32_t **cig)
{
    if (!c) return NDG_ERR_ARG;
    if (!c->paf_on) return fail(c, NDG_ERR_STATE, "ndg_set_paf() was not enabled for the last ndg_map_batch()");
    if (he) *he = c->paf_he.data();
    if (qe) *qe = c->paf_qe.data();
    if (cig) *cig = c->paf_cig.data();
    return NDG_OK;
}

extern "C" int ndg_set_aln_opt(ndg_ctx_t *c, const ndg_aln_opt_t *o)
{
    if (!c) return NDG_ERR_ARG;
    if (!o) { c->aln_on = false; return NDG_OK; }
    if (o->q == o->q2 && o->e == o->e2) return fail(c, NDG_ERR_ARG, "single-affine gap costs (q==q2, e==e2) select ksw_extz2 in the reference; only the dual-affine kernel is built");
    if (o->a <= 0 || o->b < 0 || o->q <= 0 || o->e <= 0 || o->q2 <= 0 || o->e2 <= 0 || (o->q + o->e) + (o->q2 + o->e2) > 127 || o->a + o->b > 127 - 0)
        return fail(c, NDG_ERR_ARG, "scoring out of the int8 range the reference accepts (ref: minimap2/options.c:199-207)");
    ndal::Opt &A = c->aln;
    A.a = o->a; A.b = o->b; A.q = o->q; A.e = o->e; A.q2 = o->q2; A.e2 = o->e2; A.sc_ambi = o->sc_ambi; A.zdrop = o->zdrop; A.zdrop_inv = o->zdrop_inv;
    A.end_bonus = o->end_bonus; A.min_dp_max = o->min_dp_max; A.min_ksw_len = o->min_ksw_len; A.no_end_flt = o->no_end_flt; A.max_clip_ratio = o->max_clip_ratio;
    c->aln_on = true;
    return NDG_OK;
}

extern "C" int ndg_set_threads(ndg_ctx_t *c, int n)
{
    if (!c) return NDG_ERR_ARG;
    c->n_threads = n > 0 ? n : usable_cpus();
    return c->n_threads;
}

extern "C" int ndg_aln_stats(ndg_ctx_t *c, ndg_aln_stat_t *out)
{
    if (!c || !out) return NDG_ERR_ARG;
    *out = c->aln_stat;
    return NDG_OK;
}


// ------------------------------------------------------------------------------------------------
// --step 2.  Host logic in nd_step2.hpp; every mapping runs through the GPU pipeline above.
// ------------------------------------------------------------------------------------------------
namespace {

struct ReadGather {                                // a read batch assembled from pieces of other batches (2-bit words verbatim)
    std::vector<uint32_t> words, len, nmask, hash, id, tag; std::vector<uint64_t> word_off, key; bool has_n = false;
    // n_to_a: ambiguous bases become 'A'.  The reference hands a re-mapped target over as "ACGT"[code] characters, code 4 gives the
    // string terminator, and seq_nt4_table[0] is 0 = A (ref: minimap2/map.c:1041-1045, minimap2/sketch.c:8).
    void add(const uint32_t *w, const uint64_t off, uint32_t L, const uint32_t *nm, uint32_t tg, bool n_to_a = false)
    {
        const size_t nw = (L + 15) / 16, w0 = words.size();
        word_off.push_back(w0);
        words.insert(words.end(), w + off, w + off + nw);
        if (words.size() & 1) words.push_back(0);
        if (has_n) nmask.resize(words.size() / 2 + 1, 0);
        if (nm && n_to_a) {
            for (uint32_t i = 0; i < L; ++i) if (nm[off / 2 + (i >> 5)] >> (i & 31) & 1) words[w0 + (i >> 4)] &= ~(3u << (30 - 2 * (i & 15)));
        } else if (has_n && nm) memcpy(nmask.data() + w0 / 2, nm + off / 2, ((nw + 1) / 2) * 4);
        len.push_back(L); hash.push_back(0); id.push_back((uint32_t)len.size() - 1); key.push_back(len.size()); tag.push_back(tg);
    }
    ndg_reads_t view()
    {
        ndg_reads_t v; memset(&v, 0, sizeof v);
        word_off.push_back(words.size());
        v.n_reads = (int32_t)len.size(); v.words = words.data(); v.word_off = word_off.data(); v.len = len.data(); v.nmask = has_n ? nmask.data() : nullptr;
        v.name_key = key.data(); v.name_hash = hash.data(); v.name_id = id.data(); v.tag = tag.data();
        return v;
    }
};

// all re-mappings of a query batch in one tagged index + one mapping pass of the auxiliary context
int step2_remap(ndg_ctx *c, const ndg_reads_t *Q, std::vector<nds2::QueryJob> &jobs, std::vector<nds2::Remap> &reqs)
{
    if (reqs.empty()) return 0;
    const double t0 = omp_get_wtime();
    if (!c->remap) { if (int rc = ndg_init(c->device, &c->remap)) return fail(c, rc, "cannot create the re-mapping context"); }
    ndg_ctx *R = c->remap;
    const uint64_t launches0 = ndg_kernel_launches(R);
    ndg_opt_t o2 = c->opt; o2.k = c->s2.kn; o2.w = c->s2.wn; o2.flag_no_diag = 0; o2.flag_no_dual = 0; o2.mid_occ = c->mid_occ(); o2.min_mid_occ = 0;
    if (int rc = ndg_set_opt(R, &o2)) return fail(c, rc, "%s", ndg_last_error(R));
    ndg_set_limits(R, c->sub_batch_bases, c->max_anchors);
    const bool any_n = Q->nmask || c->ht->has_n;
    ReadGather IX, MP; IX.has_n = MP.has_n = any_n;
    const uint32_t *tn = c->ht->has_n ? c->ht->nmask.data() : nullptr;
    std::vector<uint32_t> qtag(jobs.size(), 0), qslot(jobs.size(), 0), base(reqs.size(), 0);
    uint32_t next_tag = 1;
    for (size_t x = 0; x < reqs.size(); ++x) {
        nds2::Remap &m = reqs[x];
        if (m.query_is_index) {
            if (!qtag[(size_t)m.query]) {           // the query becomes a one-read mini index (ref: map.c:1047)
                qtag[(size_t)m.query] = next_tag++;
                qslot[(size_t)m.query] = (uint32_t)IX.len.size();
                IX.add(Q->words, Q->word_off[m.query], Q->len[m.query], Q->nmask, qtag[(size_t)m.query]);
            }
            base[x] = qslot[(size_t)m.query];
            MP.add(c->ht->words.data(), c->ht->word_off[(size_t)m.mapped_rid], c->h_tlen[(size_t)m.mapped_rid], tn, qtag[(size_t)m.query], true);
        } else {                                    // up to cn targets form the mini index, the query is mapped (ref: map.c:1085-1125)
            const uint32_t tg = next_tag++;
            base[x] = (uint32_t)IX.len.size();
            for (int32_t rid : m.group_rids) IX.add(c->ht->words.data(), c->ht->word_off[(size_t)rid], c->h_tlen[(size_t)rid], tn, tg);
            MP.add(Q->words, Q->word_off[m.query], Q->len[m.query], Q->nmask, tg);
            ++c->s2_groups;
        }
    }
    ndg_reads_t iv = IX.view(), mv = MP.view();
    ndg_index_stat_t st;
    if (ndg_index_begin(R) || ndg_index_add(R, &iv) || ndg_index_finalize(R, &st)) return fail(c, NDG_ERR_STATE, "re-mapping index: %s", ndg_last_error(R));
    ndg_map_result_t res;
    if (int rc = ndg_map_batch(R, &mv, &res)) return fail(c, rc, "re-mapping: %s", ndg_last_error(R));
    size_t h = 0;
    for (size_t x = 0; x < reqs.size(); ++x) {
        nds2::Remap &m = reqs[x];
        m.result.clear();
        while (h < (size_t)res.n_hits && res.hits[h].qidx == (int32_t)x) {
            const ndg_hit_t &g = res.hits[h++];
            nds2::Reg z; z.rid = g.rid - (int32_t)base[x]; z.rev = g.rev; z.qs = g.qs; z.qe = g.qe; z.rs = g.rs; z.re = g.re; z.mlen = g.mlen; z.blen = g.blen; z.score = g.score; z.cnt = g.cnt;
            m.result.push_back(z);
        }
    }
    c->s2_remaps += reqs.size(); c->launches += ndg_kernel_launches(R) - launches0; c->s2_remap_ms += (omp_get_wtime() - t0) * 1e3;
    return 0;
}

} // namespace

extern "C" int ndg_step2_begin(ndg_ctx_t *c, const ndg_step2_opt_t *o)
{
    if (!c) return NDG_ERR_ARG;
    if (!o) { c->s2_on = false; return NDG_OK; }
    if (o->mode != 0 && o->mode != 2) return fail(c, NDG_ERR_ARG, "--mode %d is not built (0 and 2 are)", o->mode);
    if (o->kn < 1 || o->kn > 28 || o->wn < 1 || o->wn > 255 || o->cn < 1) return fail(c, NDG_ERR_ARG, "kn/wn/cn out of range");
    nds2::Opt &S = c->s2;
    S.minlen = o->minlen; S.maxhan1 = o->maxhan1; S.maxhan2 = o->maxhan2; S.minmatch = o->minmatch; S.mode = o->mode; S.cn = o->cn; S.kn = o->kn; S.wn = o->wn; S.outctn = o->outctn; S.minide = o->minide;
    c->s2w = nds2::Writer(); c->s2w.begin(); c->s2_emitted = 0; c->s2_on = true; c->s2_remaps = c->s2_groups = 0; c->s2_remap_ms = 0;
    return NDG_OK;
}

extern "C" int ndg_step2_map_batch(ndg_ctx_t *c, const ndg_reads_t *Q, const uint8_t **ovl, uint64_t *n_bytes, uint64_t *n_rec, uint64_t *sum_qspan,
                                   ndg_map_result_t *work)
{
    if (!c || !Q || !ovl || !n_bytes || Q->n_reads <= 0) return NDG_ERR_ARG;
    if (!c->s2_on) return fail(c, NDG_ERR_STATE, "ndg_step2_begin() was not called");
    if (c->aln_on) return fail(c, NDG_ERR_STATE, "--step 2 with base-level alignment is not built");
    if (!Q->name_id || !c->t_has_id) return fail(c, NDG_ERR_ARG, "step 2 needs name_id for targets and queries");
    if (c->ht->word_off.size() != c->n_targets) return fail(c, NDG_ERR_STATE, "step 2 needs ndg_step2_begin() before ndg_index_add() (target bases are kept on the host)");
    omp_set_num_threads(c->n_threads);
    // the bytes handed over by the previous call are the caller's now: drop them (the reference flushes every 100 MB, lib/ovl.c:247-250)
    if (c->s2_emitted) { c->s2w.out.erase(c->s2w.out.begin(), c->s2w.out.begin() + (std::ptrdiff_t)c->s2_emitted); c->s2_emitted = 0; }
    ndg_map_result_t R;
    if (int rc = ndg_map_batch(c, Q, &R)) return rc;
    const int nq = Q->n_reads;
    std::vector<nds2::QueryJob> jobs((size_t)nq);
    {
        int64_t h = 0;
        for (int q = 0; q < nq; ++q) {
            nds2::QueryJob &J = jobs[(size_t)q];
            J.qlen = Q->len[q]; J.qkey = Q->name_key ? Q->name_key[q] : (uint64_t)q;
            while (h < R.n_hits && R.hits[h].qidx == q) {
                const ndg_hit_t &g = R.hits[h++];
                nds2::Reg z; z.rid = g.rid; z.rev = g.rev; z.qs = g.qs; z.qe = g.qe; z.rs = g.rs; z.re = g.re; z.mlen = g.mlen; z.blen = g.blen; z.score = g.score; z.cnt = g.cnt;
                J.regs.push_back(z);
            }
        }
    }
    std::vector<nds2::Remap> reqs; std::vector<size_t> rbeg((size_t)nq + 1, 0);
    for (int q = 0; q < nq; ++q) {
        nds2::mark_candidates(c->s2, jobs[(size_t)q], c->h_tkey.data(), c->h_tlen.data());
        rbeg[(size_t)q] = reqs.size();
        nds2::plan_remaps(c->s2, q, jobs[(size_t)q], reqs);
    }
    rbeg[(size_t)nq] = reqs.size();
    if (int rc = step2_remap(c, Q, jobs, reqs)) return rc;
#pragma omp parallel for schedule(dynamic, 8)
    for (int q = 0; q < nq; ++q) nds2::apply_remaps(c->s2, jobs[(size_t)q], reqs, rbeg[(size_t)q], rbeg[(size_t)q + 1], c->h_tlen.data());
    uint64_t rec = 0, sq = 0;
    for (int q = 0; q < nq; ++q)
        c->s2w.emit_query(c->s2, jobs[(size_t)q].regs, Q->name_id[q], jobs[(size_t)q].qkey, Q->len[q], c->h_tid.data(), c->h_tkey.data(), c->h_tlen.data(), &rec, &sq);
    *ovl = c->s2w.out.data() + c->s2_emitted; *n_bytes = c->s2w.out.size() - c->s2_emitted; c->s2_emitted = c->s2w.out.size();
    if (n_rec) *n_rec += rec;
    if (sum_qspan) *sum_qspan += sq;
    if (work) { *work = R; work->hits = nullptr; }
    return NDG_OK;
}

extern "C" int ndg_step2_finish(ndg_ctx_t *c, const char **bl, uint64_t *n)
{
    if (!c || !bl || !n) return NDG_ERR_ARG;
    if (!c->s2_on) return fail(c, NDG_ERR_STATE, "ndg_step2_begin() was not called");
    c->s2_bl = c->s2w.bl_text();
    *bl = c->s2_bl.c_str(); *n = c->s2_bl.size();
    return NDG_OK;
}

// ------------------------------------------------------------------------------------------------
// introspection
// ------------------------------------------------------------------------------------------------
extern "C" int64_t ndg_debug_fetch(ndg_ctx_t *c, const char *what, void *dst, uint64_t cap)
{
    if (!c || !what) return NDG_ERR_ARG;
    cudaSetDevice(c->device);
    const void *src = nullptr; uint64_t bytes = 0;
    std::string w = what;
    if (w == "mz") { src = c->mz.p; bytes = c->n_mz * 16; }
    else if (w == "mz_off") { src = c->mz_off.p; bytes = (uint64_t)(c->q.n_reads + 1) * 8; }
    else if (w == "anchors") { src = c->a.p; bytes = c->n_a * 16; }
    else if (w == "a_off") { src = c->a_off.p; bytes = (uint64_t)(c->q.n_reads + 1) * 8; }
    else if (w == "f") { src = c->f.p; bytes = c->n_a * 4; }
    else if (w == "p") { src = c->p.p; bytes = c->n_a * 4; }
    else if (w == "v") { src = c->v.p; bytes = c->n_a * 4; }
    else if (w == "seg_start") { src = c->seg_start.p; bytes = c->n_seg * 4; }
    else if (w == "ksw_p") { src = c->ksw_p.p; bytes = c->ksw_p.cap < (64u << 20) ? c->ksw_p.cap : (64u << 20); }
    else return fail(c, NDG_ERR_ARG, "unknown stage buffer '%s'", what);
    if (!dst) return (int64_t)bytes;
    if (bytes > cap) return fail(c, NDG_ERR_OVERFLOW, "buffer too small");
    if (bytes) { cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->st); if (cudaStreamSynchronize(c->st) != cudaSuccess) return fail(c, NDG_ERR_CUDA, "fetch"); }
    return (int64_t)bytes;
}

// ------------------------------------------------------------------------------------------------
// .ovl encoder, host side, table-free.  ref: lib/ovl.c:10-29 (varint table) and :109-150 (encode_ovl)
// ------------------------------------------------------------------------------------------------
static inline int ndg_put_varint(uint8_t *dst, uint32_t v)
{
    if (v < 0x80) { dst[0] = (uint8_t)v; return 1; }
    int ng = v < (1u << 14) ? 2 : v < (1u << 21) ? 3 : v < (1u << 28) ? 4 : 5;
    for (int g = 0; g < ng; ++g) dst[g] = (uint8_t)((v >> (7 * (ng - 1 - g)) & 0x7f) | (g + 1 < ng ? 0x80 : 0));
    return ng;
}

extern "C" int ndg_ovl_encode8(uint8_t *dst, uint32_t state[2], int rev, uint32_t qname, uint32_t qs, uint32_t qe,
                               uint32_t tname, uint32_t ts, uint32_t te, uint32_t match)
{
    uint32_t fld[8]; uint32_t ql = qe - qs, tl = te - ts; uint32_t flags = (uint32_t)rev & 0xff;
    if (qname >= state[0]) fld[0] = qname - state[0]; else { flags |= 2; fld[0] = state[0] - qname; }
    if (tname >= state[1]) fld[4] = tname - state[1]; else { flags |= 4; fld[4] = state[1] - tname; }
    if (ql >= tl) fld[6] = ql - tl; else { flags |= 8; fld[6] = tl - ql; }
    state[0] = qname; state[1] = tname;
    fld[1] = flags; fld[2] = qs; fld[3] = ql; fld[5] = ts; fld[7] = match;
    int n = 0;
    for (int i = 0; i < 8; ++i) n += ndg_put_varint(dst + n, fld[i]);
    return n;
}

// (no counterpart in the reference: stitches streams that were encoded independently; the delta/flag rules are encode_ovl's, lib/ovl.c:109-150)
extern "C" int ndg_ovl_rebase_first(const uint8_t *ovl, uint64_t n, const uint32_t prev[2], uint8_t *dst, uint32_t *skip)
{
    if (!ovl || !prev || !dst || !skip) return NDG_ERR_ARG;
    uint32_t v[8]; uint64_t p = 0;
    for (int f = 0; f < 8; ++f) {                    // big-endian base-128 groups, continuation bit on all but the last byte
        uint64_t x = 0; int nb = 0;
        for (;;) { if (p >= n || ++nb > 5) return NDG_ERR_ARG; const uint8_t b = ovl[p++]; x = x << 7 | (b & 127); if (b < 128) break; }
        v[f] = (uint32_t)x;
    }
    *skip = (uint32_t)p;
    // encoded from state (0,0): the name fields are absolute and their "decreased" flags are clear
    const uint32_t qname = v[0], tname = v[4];
    uint32_t flags = v[1] & ~6u;
    if (qname >= prev[0]) v[0] = qname - prev[0]; else { flags |= 2; v[0] = prev[0] - qname; }
    if (tname >= prev[1]) v[4] = tname - prev[1]; else { flags |= 4; v[4] = prev[1] - tname; }
    v[1] = flags;
    int m = 0;
    for (int f = 0; f < 8; ++f) m += ndg_put_varint(dst + m, v[f]);
    return m;
}

// ref: minimap2/map.c:804-821 (check_realign_nextdenovo), used by --dvt at map.c:1302
static inline int ndg_check_dovetail(int rev, uint32_t qs, uint32_t qe, uint32_t qlen, uint32_t ts, uint32_t te, uint32_t tlen,
                                     int32_t maxhan1, int32_t maxhan2)
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

extern "C" int64_t ndg_ovl_emit_step1(const ndg_hit_t *hits, int64_t n_hits,
                                      const uint32_t *q_id, const uint64_t *q_key, const uint32_t *q_len,
                                      const uint32_t *t_id, const uint64_t *t_key, const uint32_t *t_len,
                                      int32_t minlen, int32_t dvt, int32_t maxhan1, int32_t maxhan2,
                                      uint32_t state[2], uint8_t *dst, uint64_t cap, uint64_t *n_rec, uint64_t *sum_qspan)
{
    uint64_t n = 0, rec = 0, sq = 0;
    for (int64_t i = 0; i < n_hits; ++i) {
        const ndg_hit_t &h = hits[i];
        if (q_key[h.qidx] == t_key[h.rid]) continue;
        if (h.qe - h.qs < minlen) continue;
        if (dvt && !ndg_check_dovetail(h.rev, h.qs, h.qe, q_len[h.qidx], h.rs, h.re, t_len[h.rid], maxhan1, maxhan2)) continue;
        if (n + 40 > cap) return NDG_ERR_OVERFLOW;
        n += ndg_ovl_encode8(dst + n, state, h.rev, q_id[h.qidx], (uint32_t)h.qs, (uint32_t)h.qe, t_id[h.rid], (uint32_t)h.rs, (uint32_t)h.re, (uint32_t)h.mlen);
        ++rec; sq += (uint64_t)(h.qe - h.qs);
    }
    if (n_rec) *n_rec += rec;
    if (sum_qspan) *sum_qspan += sq;
    return (int64_t)n;
}

// ------------------------------------------------------------------------------------------------
// ksw_extd2 batch.  ref: minimap2/ksw2_extd2_sse.c (one call per alignment), here one warp each
// ------------------------------------------------------------------------------------------------
static_assert(sizeof(ndg_ksw_result_t) == sizeof(ndk_res), "ksw result layout");

// variant 3: every call the register-resident kernel takes goes to it (grouped by ring size and maximum mode), the rest to the
// memory-resident kernels.  Sequences are re-materialised in the padded layout nd_kswrx.cuh expects.
static int ksw_batch_rx(ndg_ctx *c, const ndg_ksw_call_t *calls, int32_t n, const uint8_t *seqs, uint64_t seq_bytes, const ndk_par &P,
                        ndg_ksw_result_t *res, uint32_t *cigar, const uint64_t *cigar_off)
{
    struct Grp { int cfg; bool exact; std::vector<int> ids; };
    std::vector<Grp> groups;
    std::vector<int> legacy;
    for (int i = 0; i < n; ++i) {
        const ndg_ksw_call_t &k = calls[i];
        const int cls = rx_class(k.qlen, k.tlen, k.w, k.flag);
        if (cls < 0) { legacy.push_back(i); continue; }
        const bool ex = !(k.flag & NDK_APPROX_MAX);
        size_t g = 0;
        for (; g < groups.size(); ++g) if (groups[g].cfg == cls && groups[g].exact == ex) break;
        if (g == groups.size()) groups.push_back(Grp{ cls, ex, {} });
        groups[g].ids.push_back(i);
    }
    // padded sequences + call table in group order
    std::vector<ndk_call> hc; hc.reserve((size_t)n);
    std::vector<int> order; order.reserve((size_t)n);
    uint64_t sb = 0, p_total = 0;
    for (auto &g : groups) for (int i : g.ids) { order.push_back(i); sb += rx_seq_bytes(calls[i].qlen, calls[i].tlen); }
    std::vector<uint8_t> hs(sb + 64, 0);
    uint64_t at = 0;
    for (int i : order) {
        const ndg_ksw_call_t &k = calls[i];
        ndk_call d; memset(&d, 0, sizeof d);
        uint64_t qo, to; rx_seq_offsets(at, k.qlen, &qo, &to);
        if (k.qlen > 0) memcpy(hs.data() + qo, seqs + k.q_off, (size_t)k.qlen);
        if (k.tlen > 0) memcpy(hs.data() + to, seqs + k.t_off, (size_t)k.tlen);
        bool ambi = false;
        for (int j = 0; j < k.qlen && !ambi; ++j) ambi = seqs[k.q_off + j] == (uint8_t)(P.m - 1);
        for (int j = 0; j < k.tlen && !ambi; ++j) ambi = seqs[k.t_off + j] == (uint8_t)(P.m - 1);
        d.qlen = k.qlen; d.tlen = k.tlen; d.q_off = qo; d.t_off = to; d.w = k.w; d.zdrop = k.zdrop; d.end_bonus = k.end_bonus; d.flag = k.flag | (ambi ? NDX_FLAG_AMBI : 0);
        d.cigar_off = cigar_off[i]; d.p_off = p_total;
        if (k.qlen > 0 && k.tlen > 0 && !(k.flag & NDK_SCORE_ONLY)) {
            int w = k.w < 0 ? std::max(k.qlen, k.tlen) : k.w;
            int ncol = std::min(k.qlen, k.tlen); ncol = (std::min(ncol, w + 1) + 15) / 16 + 1;
            p_total += ((uint64_t)(k.qlen + k.tlen - 1) * ncol + 1) * 16;
        }
        at += rx_seq_bytes(k.qlen, k.tlen);
        hc.push_back(d);
    }
    const size_t nr = order.size();
    const uint64_t cig_total = cigar_off[n];
    if (nr) {
        RSV(c->ksw_calls, nr * sizeof(ndk_call)); RSV(c->ksw_seqs, sb + 64); RSV(c->ksw_res, nr * sizeof(ndk_res));
        RSV(c->ksw_p, p_total + 64); RSV(c->ksw_cig, (cig_total + 4) * 4); RSV(c->scalars, 1024);
        CK(cudaMemcpyAsync(c->ksw_calls.p, hc.data(), nr * sizeof(ndk_call), cudaMemcpyHostToDevice, c->st));
        CK(cudaMemcpyAsync(c->ksw_seqs.p, hs.data(), sb + 64, cudaMemcpyHostToDevice, c->st));
        uint32_t *d_work = c->scalars.as<uint32_t>() + 128;
        CK(cudaMemsetAsync(d_work, 0, 64, c->st));
        size_t off = 0; int gi = 0;
        for (auto &g : groups) {
            KLAUNCH("ksw_extd2", CK(rx_launch(g.cfg, g.exact, 0, c->n_sm, c->st, c->ksw_calls.as<ndk_call>() + off, (int)g.ids.size(), c->ksw_seqs.as<uint8_t>(), P,
                                              c->ksw_res.as<ndk_res>() + off, c->ksw_p.as<uint8_t>(), c->ksw_cig.as<uint32_t>(), nullptr, 0, d_work + gi, nullptr, 0)));
            off += g.ids.size(); ++gi;
            if (gi >= 16) return fail(c, NDG_ERR_ARG, "too many kernel groups");
        }
        std::vector<ndk_res> hr(nr);
        CK(cudaMemcpyAsync(hr.data(), c->ksw_res.p, nr * sizeof(ndk_res), cudaMemcpyDeviceToHost, c->st));
        if (cigar && cig_total) CK(cudaMemcpyAsync(cigar, c->ksw_cig.p, cig_total * 4, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        for (size_t x = 0; x < nr; ++x) memcpy(&res[order[x]], &hr[x], sizeof(ndk_res));
    }
    if (!legacy.empty()) {                               // generic scoring / bands wider than every ring: the memory-resident kernels
        std::vector<ndg_ksw_call_t> lc; std::vector<uint64_t> lo(1, 0);
        for (int i : legacy) { lc.push_back(calls[i]); lo.push_back(lo.back() + (uint64_t)calls[i].qlen + calls[i].tlen + 2); }
        std::vector<ndg_ksw_result_t> lr(lc.size()); std::vector<uint32_t> lcg(lo.back() + 4);
        const int keep = c->ksw_variant; c->ksw_variant = 1;
        int8_t mat[25]; memcpy(mat, P.mat, 25);
        int rc = ndg_ksw_extd2_batch(c, lc.data(), (int32_t)lc.size(), seqs, seq_bytes, P.m, mat, P.q, P.e, P.q2, P.e2, lr.data(), lcg.data(), lo.data());
        c->ksw_variant = keep;
        if (rc) return rc;
        for (size_t x = 0; x < legacy.size(); ++x) {
            res[legacy[x]] = lr[x];
            if (cigar) memcpy(cigar + cigar_off[legacy[x]], lcg.data() + lo[x], (size_t)lr[x].n_cigar * 4);
        }
    }
    return NDG_OK;
}

extern "C" int ndg_ksw_extd2_batch(ndg_ctx_t *c, const ndg_ksw_call_t *calls, int32_t n, const uint8_t *seqs, uint64_t seq_bytes,
                                   int32_t m, const int8_t *mat, int32_t q, int32_t e, int32_t q2, int32_t e2,
                                   ndg_ksw_result_t *res, uint32_t *cigar, const uint64_t *cigar_off)
{
    if (!c || !calls || n <= 0 || !seqs || !mat || !res || !cigar_off || m < 2 || m > 5) return NDG_ERR_ARG;
    cudaSetDevice(c->device);
    if (c->ksw_variant == 3) {
        ndk_par P; P.m = m; memset(P.mat, 0, sizeof P.mat); memcpy(P.mat, mat, (size_t)m * m); P.q = q; P.e = e; P.q2 = q2; P.e2 = e2;
        for (int i = 0; i < n; ++i) if (calls[i].qlen < 0 || calls[i].tlen < 0 || calls[i].q_off + (uint64_t)calls[i].qlen > seq_bytes || calls[i].t_off + (uint64_t)calls[i].tlen > seq_bytes)
            return fail(c, NDG_ERR_ARG, "ksw call %d out of range", i);
        return ksw_batch_rx(c, calls, n, seqs, seq_bytes, P, res, cigar, cigar_off);
    }
    std::vector<ndk_call> hc((size_t)n);
    uint64_t p_total = 0; int Tmax = 16, rmax = 1, qmax = 1;
    for (int i = 0; i < n; ++i) {
        const ndg_ksw_call_t &k = calls[i];
        if (k.qlen < 0 || k.tlen < 0 || k.q_off + (uint64_t)k.qlen > seq_bytes || k.t_off + (uint64_t)k.tlen > seq_bytes) return fail(c, NDG_ERR_ARG, "ksw call %d out of range", i);
        ndk_call &d = hc[i];
        d.qlen = k.qlen; d.tlen = k.tlen; d.q_off = k.q_off; d.t_off = k.t_off; d.w = k.w; d.zdrop = k.zdrop; d.end_bonus = k.end_bonus; d.flag = k.flag;
        d.cigar_off = cigar_off[i]; d.p_off = p_total;
        if (k.qlen > 0 && k.tlen > 0) {
            int w = k.w < 0 ? std::max(k.qlen, k.tlen) : k.w;
            int ncol = std::min(k.qlen, k.tlen); ncol = (std::min(ncol, w + 1) + 15) / 16 + 1;
            if (!(k.flag & NDK_SCORE_ONLY)) p_total += ((uint64_t)(k.qlen + k.tlen - 1) * ncol + 1) * 16;
            Tmax = std::max(Tmax, (k.tlen + 15) / 16 * 16); rmax = std::max(rmax, k.qlen + k.tlen); qmax = std::max(qmax, k.qlen);
        }
    }
    const size_t state_bytes = ((size_t)17 * Tmax + ((size_t)(qmax + 15) / 16 + 1) * 16 + 96 + 255) / 256 * 256;   // cells 8T, sf T, qr
    const size_t h_bytes = ((size_t)4 * Tmax + 255) / 256 * 256, off_bytes = ((size_t)8 * rmax + 255) / 256 * 256;
    const size_t slot_bytes = state_bytes + h_bytes + off_bytes;
    const int blocks = std::min<int>(c->n_sm * 16, (n + NDK_WARPS - 1) / NDK_WARPS);
    const size_t slots = (size_t)blocks * NDK_WARPS;
    const uint64_t cig_total = cigar_off[n];
    RSV(c->ksw_calls, (size_t)n * sizeof(ndk_call)); RSV(c->ksw_seqs, seq_bytes + 64); RSV(c->ksw_res, (size_t)n * sizeof(ndk_res));
    RSV(c->ksw_p, p_total + 64); RSV(c->ksw_cig, (cig_total + 4) * 4); RSV(c->ksw_scratch, slots * slot_bytes); RSV(c->scalars, 1024);
    CK(cudaMemcpyAsync(c->ksw_calls.p, hc.data(), (size_t)n * sizeof(ndk_call), cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(c->ksw_seqs.p, seqs, seq_bytes, cudaMemcpyHostToDevice, c->st));
    uint32_t *d_work = c->scalars.as<uint32_t>() + 128;
    CK(cudaMemsetAsync(d_work, 0, 4, c->st));
    ndk_par P; P.m = m; memset(P.mat, 0, sizeof P.mat); memcpy(P.mat, mat, (size_t)m * m); P.q = q; P.e = e; P.q2 = q2; P.e2 = e2;
    if (c->ksw_variant == 2) {                       // block-per-call generation (wide bands)
        const int cb = std::min<int>(c->n_sm * (16 / NDK_COOP_WARPS), n);
        RSV(c->ksw_scratch, (size_t)cb * slot_bytes);
        KLAUNCH("ksw_extd2", (k_ksw_extd2_coop<<<cb, NDK_COOP_WARPS * 32, 0, c->st>>>(c->ksw_calls.as<ndk_call>(), n, c->ksw_seqs.as<uint8_t>(), P, c->ksw_res.as<ndk_res>(),
                              c->ksw_p.as<uint8_t>(), c->ksw_cig.as<uint32_t>(), c->ksw_scratch.as<uint8_t>(), slot_bytes, state_bytes, h_bytes, d_work)));
    } else if (c->ksw_variant == 1)
        KLAUNCH("ksw_extd2", (k_ksw_extd2_px<<<blocks, NDK_WARPS * 32, 0, c->st>>>(c->ksw_calls.as<ndk_call>(), n, c->ksw_seqs.as<uint8_t>(), P, c->ksw_res.as<ndk_res>(),
                              c->ksw_p.as<uint8_t>(), c->ksw_cig.as<uint32_t>(), c->ksw_scratch.as<uint8_t>(), slot_bytes, state_bytes, h_bytes, d_work)));
    else
    KLAUNCH("ksw_extd2", (k_ksw_extd2<<<blocks, NDK_WARPS * 32, 0, c->st>>>(c->ksw_calls.as<ndk_call>(), n, c->ksw_seqs.as<uint8_t>(), P, c->ksw_res.as<ndk_res>(),
                          c->ksw_p.as<uint8_t>(), c->ksw_cig.as<uint32_t>(), c->ksw_scratch.as<uint8_t>(), slot_bytes, state_bytes, h_bytes, d_work)));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(res, c->ksw_res.p, (size_t)n * sizeof(ndk_res), cudaMemcpyDeviceToHost, c->st));
    if (cigar && cig_total) CK(cudaMemcpyAsync(cigar, c->ksw_cig.p, cig_total * 4, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return NDG_OK;
}
