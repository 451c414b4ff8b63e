/*
 * ndgpu.h -- C ABI of the B200-native read-overlap engine (libndgpu.so).
 *
 * The reference (Nextomics/NextDenovo, minimap2-nd) has no plugin/FFI interface for this path: its
 * contract is the process boundary (argv in, .ovl bytes out).  This header is the seam *inside*
 * that process where "host C calls CUDA": each entry point below replaces one stage of the
 * reference's in-process pipeline and is exactly what a maintainer would bind from
 * minimap2/main.c / minimap2/map.c (see INTEGRATION.md for the stub).  All "ref:" citations are
 * relative to the reference tree.
 *
 * Conventions: plain C, caller-owned host buffers unless stated, int return codes (0 = success,
 * negative = failure; ndg_last_error() gives the message), no exceptions cross the boundary, one
 * context per GPU and per host thread.  There is no CPU fallback: every compute entry point fails
 * with NDG_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef NDGPU_H
#define NDGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDG_OK            0
#define NDG_ERR_ARG      (-1)
#define NDG_ERR_CUDA     (-2)
#define NDG_ERR_NOMEM    (-3)
#define NDG_ERR_OVERFLOW (-4)
#define NDG_ERR_STATE    (-5)

typedef struct ndg_ctx ndg_ctx_t;

/* Effective options of the hot path.
 * ref: mm_idxopt_t / mm_mapopt_t, minimap2/minimap.h:102-170; defaults minimap2/options.c:4-64;
 *      presets minimap2/options.c:89-97. */
typedef struct {
    int32_t k, w, hpc;                    /* -k -w -H */
    int32_t flag_no_diag, flag_no_dual;   /* MM_F_NO_DIAG, MM_F_NO_DUAL (minimap.h:8-9) */
    int32_t min_cnt, min_chain_score;     /* -n -m */
    int32_t bw, max_gap;                  /* -r -g */
    int32_t max_chain_skip, max_chain_iter;
    int32_t seed;                         /* --seed (options.c:16) */
    float   mid_occ_frac;                 /* -f (<1) */
    int32_t mid_occ;                      /* -f (>=1); <=0: derive from the first index part */
    int32_t min_mid_occ;                  /* --min-occ-floor */
} ndg_opt_t;

/* A batch of reads, 2-bit packed exactly as the words of a .2bit record
 * (ref: lib/bseq.c:114-139: 16 bases per little-endian u32, first base in the top two bits).
 * name_key: any u64 whose unsigned order equals strcmp() order of the read names and whose
 * equality equals string equality, consistently between the targets of the current index part and
 * the queries (ref: minimap2/map.c:126-148 uses strcmp on names).  name_hash: __ac_X31_hash_string
 * of the name (ref: minimap2/khash.h:383-388, used at minimap2/map.c:525). */
typedef struct {
    int32_t n_reads;
    const uint32_t *words;      /* all reads, concatenated */
    const uint64_t *word_off;   /* n_reads+1 offsets into words[] (even when nmask != NULL) */
    const uint32_t *len;        /* bases per read */
    const uint32_t *nmask;      /* optional: 1 bit per base, bit (i&31) of nmask[word_off/2 + (i>>5)] = ambiguous */
    const uint64_t *name_key;
    const uint32_t *name_hash;
    const uint32_t *name_id;    /* optional: strtoul(name) (ref: minimap2/map.c:1300-1301); needed by ndg_map_batch_ovl */
    const uint32_t *tag;        /* optional: group tag (> 0) per read.  Index reads and query reads only see each other when their tags
                                 * are equal: one launch then serves many throw-away mini indexes (the per-query indexes of --step 2,
                                 * ref: minimap2/map.c:1046-1125).  NULL = one shared index. */
} ndg_reads_t;

/* One chain-derived hit: the fields of mm_reg1_t (ref: minimap2/minimap.h:84-99) that leave the
 * hot path (.ovl writer minimap2/map.c:1299-1310, PAF writer minimap2/format.c:304-318). */
typedef struct {
    uint64_t zkey;              /* (score<<32|cnt) ^ hash: order key of mm_gen_regs (hit.c:62-71) */
    int32_t  qidx;              /* query index within the batch */
    int32_t  rid;               /* target index within the index part */
    int32_t  qs, qe, rs, re;
    int32_t  mlen, blen;
    int32_t  score, cnt;
    int32_t  rev;
    int32_t  as;
} ndg_hit_t;

typedef struct {
    int64_t n_hits;
    const ndg_hit_t *hits;      /* pinned host memory owned by the context; sorted by qidx, then in the
                                   reference's per-query output order; valid until the next map call */
    uint64_t n_minimizers, n_anchors, n_segments;   /* work counters of this call */
    double device_ms;           /* CUDA-event time from the first to the last kernel of the call (ndg_map_resident) */
} ndg_map_result_t;

typedef struct {
    uint64_t n_seq, n_minimizers, n_distinct, n_singletons, sum_len;
    int32_t mid_occ;
} ndg_index_stat_t;

/* ---- lifecycle ---------------------------------------------------------------------------- */
int  ndg_init(int device, ndg_ctx_t **ctx);
void ndg_destroy(ndg_ctx_t *ctx);
const char *ndg_last_error(const ndg_ctx_t *ctx);
const char *ndg_version(void);
int  ndg_set_opt(ndg_ctx_t *ctx, const ndg_opt_t *opt);
/* tuning: bases per device sub-batch, anchors per sub-batch (0 = default) */
int  ndg_set_limits(ndg_ctx_t *ctx, uint64_t sub_batch_bases, uint64_t max_anchors);
/* sort: 0 = one thread per read (baseline), 1 = block sort for reads without equal keys + exact warp-level replay of the reference's
 * unstable sort for the others, 2 = exact replay for every read.  chain: 0 = one thread per segment (baseline), 1 = warp per segment */
int  ndg_set_variant(ndg_ctx_t *ctx, int sort_variant, int chain_variant);
/* ksw_extd2 kernel generation: 0 = one cell per lane, 1 = two cells per lane in s16x2 (DPX) registers */
int  ndg_set_ksw_variant(ndg_ctx_t *ctx, int variant);

/* ---- regime B: base-level alignment of every chain (`-c`, MM_F_CIGAR) ---------------------------
 * Replaces align_regs() -> mm_align_skeleton() (ref: minimap2/map.c:484-505, minimap2/align.c:857-913):
 * once enabled, ndg_map_batch / ndg_map_batch_ovl return hits whose qs/qe/rs/re/mlen/blen are the
 * ones of the gapped alignment (mm_update_extra, align.c:240-286), filtered by mm_filter_regs and ordered
 * by mm_hit_sort (dp_max, ref: hit.c:169-201,257-276); ndg_hit_t.score then carries dp_max.
 * The DP runs on the GPU (ksw_extd2 kernel, one launch per round of calls); which calls to make and the
 * CIGAR stitching is host code (csrc/nd_align.hpp).  Must be set before ndg_index_add(): the library keeps a
 * host copy of the 2-bit target reads (what mm_idx_t::S is for, ref: minimap2/index.c:152-162).
 * Fields follow mm_mapopt_t (ref: minimap2/options.c:37-44): a,b match/mismatch; q,e,q2,e2 gap costs;
 * sc_ambi; zdrop, zdrop_inv; end_bonus; min_dp_max; min_ksw_len; max_clip_ratio; no_end_flt = MM_F_NO_END_FLT.
 * Requires q+e < q2+e2 or different (q,e)/(q2,e2) (the dual-affine kernel; q==q2&&e==e2 is rejected). */
typedef struct {
    int32_t a, b, q, e, q2, e2, sc_ambi, zdrop, zdrop_inv, end_bonus, min_dp_max, min_ksw_len, no_end_flt;
    float max_clip_ratio;
} ndg_aln_opt_t;
int  ndg_set_aln_opt(ndg_ctx_t *ctx, const ndg_aln_opt_t *opt);      /* NULL switches regime B off */
/* host threads of the alignment stage (the reference's -t, ref: minimap2/main.c:230); n <= 0: every CPU the process may use
 * (affinity mask capped by the cgroup CPU quota).  Returns the number in effect. */
int  ndg_set_threads(ndg_ctx_t *ctx, int n);
/* counters of the alignment stage since the last ndg_prof_reset(): chains, DP calls, GPU rounds/launches, sum(qlen*tlen);
 * p_bytes = backtrack bytes written = DP cells computed (1 byte per cell, 64-row granularity);
 * dp_ms = sum of per-launch kernel times (two launches overlap, so it can exceed dp_wall_ms = wall time spent in the DP backend);
 * host_ms = wall time of the whole alignment stage */
typedef struct { uint64_t n_chains, n_calls, n_rounds, n_launches, n_retries, rect_cells, p_bytes; double dp_ms, host_ms, dp_wall_ms; } ndg_aln_stat_t;
int  ndg_aln_stats(ndg_ctx_t *ctx, ndg_aln_stat_t *out);

/* page-locked host memory for batch inputs (optional; lets the H2D copy run at full PCIe speed) */
void *ndg_host_alloc(uint64_t bytes);
void  ndg_host_free(void *p);

/* ---- index: replaces mm_idx_gen()/mm_idx_post() (ref: minimap2/index.c:353-372,191-248) and
 *      mm_idx_cal_max_occ() (index.c:164-185) for one index part ------------------------------ */
int ndg_index_begin(ndg_ctx_t *ctx);
/* register target metadata (len, name_key) and sketch the targets on this GPU; rid continues */
int ndg_index_add(ndg_ctx_t *ctx, const ndg_reads_t *targets);
/* register metadata only (targets sketched on another GPU, see ndg_index_shard_*) */
int ndg_index_add_meta(ndg_ctx_t *ctx, const ndg_reads_t *targets);
int ndg_index_finalize(ndg_ctx_t *ctx, ndg_index_stat_t *stat);
/* N-GPU build of one index part (no counterpart in the reference, whose index is built by one process; the collectives are the
 * caller's: NCCL from one process per GPU, or peer copies between contexts of one process).  The table is range-partitioned by
 * minimizer: rank r builds slice r from the minimizers of its key range and the finished slices are all-gathered, so the sort is
 * sharded N ways instead of replicated.  Call order on every rank r (target order = rank order: ndg_index_add_meta() for the reads
 * of ranks < r, ndg_index_add() for the own ones, ndg_index_add_meta() for ranks > r):
 *   split(N, counts)          own minimizers sorted and cut into the N key ranges; counts[j] pairs go to rank j
 *   bucket(j, &k, &v, &n)     device pointers of those pairs (two u64 arrays)          -> all-to-all, received in rank order
 *   recv_reserve(total,&k,&v) device buffers for everything destined for this rank
 *   sort(&n_distinct)         received pairs sorted, run lengths taken
 *   place(r, max_nd, hist)    max_nd = max over ranks of n_distinct: common slice size; slice r built; hist[65536] = run-length histogram
 *   table_slice(j, &p, &b)    where slice j lives (equal sizes)                        -> all-gather in place
 *   pos_reserve(r, n[N], seg) n[j] = pairs sorted on rank j; position list allocated; seg[j] = device pointer of segment j -> broadcast each
 *   finish(hist_sum, nd_sum)  global histogram / distinct count (mid_occ, ref: minimap2/index.c:164-185 needs the whole part) */
/* device-to-device copy between buffers owned by two contexts (same GPU or peers), synchronous: the building block of the collectives
 * when all N contexts live in one process (bin/minimap2-nd --gpus N) */
int ndg_copy_between(ndg_ctx_t *dst_ctx, void *dst, ndg_ctx_t *src_ctx, const void *src, uint64_t bytes);
/* host helper: the slice (= rank) the minimizer key (hash, < 4^k) belongs to in an n_slices-way build */
int32_t ndg_index_key_slice(uint64_t key, int32_t k, int32_t n_slices);
int ndg_index_shard_split(ndg_ctx_t *ctx, int32_t n_slices, uint64_t *counts);
int ndg_index_shard_bucket(ndg_ctx_t *ctx, int32_t j, const void **dev_keys, const void **dev_vals, uint64_t *n);
int ndg_index_shard_recv_reserve(ndg_ctx_t *ctx, uint64_t total, void **dev_keys, void **dev_vals);
int ndg_index_shard_sort(ndg_ctx_t *ctx, uint64_t *n_distinct);
int ndg_index_shard_place(ndg_ctx_t *ctx, int32_t my_slice, uint64_t max_distinct_per_slice, uint32_t *hist65536);
int ndg_index_shard_table_slice(ndg_ctx_t *ctx, int32_t j, void **dev, uint64_t *bytes);
int ndg_index_shard_pos_reserve(ndg_ctx_t *ctx, int32_t my_slice, const uint64_t *n_pos_per_slice, void **dev_segments);
int ndg_index_shard_finish(ndg_ctx_t *ctx, const uint32_t *hist65536, uint64_t n_distinct, ndg_index_stat_t *stat);
/* Let ctx borrow the finalized index of src (same device; no counterpart in the reference, where worker threads share `mi`):
 * several contexts, each with its own stream and workspace, can then map batches concurrently against one copy of the index.
 * src must stay alive and keep its index while ctx uses it; ndg_index_begin()/ndg_destroy() on ctx drop the loan. */
int ndg_index_share(ndg_ctx_t *ctx, ndg_ctx_t *src);

/* point query, for tests: replaces mm_idx_get (ref: minimap2/index.c:81-98). positions copied to host */
int ndg_index_get(ndg_ctx_t *ctx, uint64_t minier, uint64_t *pos, int cap, int *n);

/* ---- mapping: replaces kt_for(worker_for) -> mm_map_frag() for a batch of queries
 *      (ref: minimap2/map.c:1241-1244, 879-906, 507-623), chain-only regime ------------------ */
int ndg_map_batch(ndg_ctx_t *ctx, const ndg_reads_t *queries, ndg_map_result_t *out);
/* Whole step 1 for a batch: mapping plus the reference's output stage (filter + delta/varint records,
 * ref: minimap2/map.c:1297-1304, lib/ovl.c:109-150) done on the GPU; returns the .ovl bytes of the batch
 * (pinned host memory owned by the context, valid until the next call).  state[2] = the encoder's
 * running (prev_qname, prev_tname) (ref: minimap2/main.c:29), updated in place.  Needs name_id on both the
 * targets (at ndg_index_add) and the queries. */
typedef struct { int32_t minlen, dvt, maxhan1, maxhan2; } ndg_emit_opt_t;
int ndg_map_batch_ovl(ndg_ctx_t *ctx, const ndg_reads_t *queries, const ndg_emit_opt_t *eo, uint32_t state[2],
                      const uint8_t **ovl, uint64_t *n_bytes, uint64_t *n_rec, uint64_t *sum_qspan, ndg_map_result_t *work);
/* Start the host-to-device copy of the NEXT query batch on a copy stream and return at once (what the reference's pipeline step 0
 * does for the next batch while step 1 maps the current one, ref: minimap2/map.c:1389-1390).  The following ndg_map_batch_ovl() on
 * the same buffers finds the batch resident.  A batch larger than one device sub-batch is left to the mapping call.  The host
 * buffers must stay untouched until that call returns. */
int ndg_map_prefetch(ndg_ctx_t *ctx, const ndg_reads_t *queries);
/* the same in three steps, for measurement with inputs resident in HBM */
int ndg_map_upload(ndg_ctx_t *ctx, const ndg_reads_t *queries);      /* H2D only (single sub-batch) */
int ndg_map_resident(ndg_ctx_t *ctx, ndg_map_result_t *out);          /* kernels only; hits stay on device */
int ndg_map_download(ndg_ctx_t *ctx, ndg_map_result_t *out);          /* D2H of the hits */

/* ---- PAF extras: what mm_write_paf3 prints beyond the coordinates (ref: minimap2/format.c:282-318) and what mm_set_mapq / mm_est_err
 * need (ref: minimap2/hit.c:446-491, minimap2/esterr.c:30-64, minimap2/map.c:91-124 rep_len).  Off by default: nothing of this
 * reaches .ovl.  After ndg_set_paf(ctx, 1) every ndg_map_batch() also fills, parallel to its hit list:
 *   per hit  : regime A: the raw counts of mm_est_err (n_match, n_span = en - st + 1 or -1 when the chain start is not found);
 *              regime B: dp_score, dp_max, n_ambi of mm_extra_t, the CIGAR (BAM encoding, into the pool), the chain scores;
 *              flags bit0 inv, bit1 id == parent (only split-off and inversion hits in all-chain mode), bits 2-3 split, bit4 has a CIGAR
 *   per query: rep_len, the number of minimizers that passed the occurrence filter and the sum of their spans (avg_k = sum/n).
 * The float arithmetic (logf) is the writer's, on the host, as in the reference. */
typedef struct { int32_t n_match, n_span, dp_score, dp_max, n_ambi, score0, chain_score; uint32_t flags, n_cigar; uint64_t cigar_off; } ndg_hit_extra_t;
typedef struct { int32_t rep_len, n_mini; uint64_t sum_span; } ndg_query_extra_t;
int ndg_set_paf(ndg_ctx_t *ctx, int on);
int ndg_map_extras(ndg_ctx_t *ctx, const ndg_hit_extra_t **hit_extra, const ndg_query_extra_t **query_extra, const uint32_t **cigar_pool);

/* ---- `--step 2` (cns_align): first-pass mapping, per-query re-mapping of the dovetail candidates against throw-away mini
 * indexes, 10-field records after the contained/dovetail bookkeeping, `.bl` side file.
 * Replaces the step-2 branch of worker_for() and of the writer (ref: minimap2/map.c:988-1128,1305-1328, lib/ovl.c:205-253,339-361,
 * 449-563).  Every mapping (first pass and the re-mappings: each candidate target against a (kn,wn) index of the query, or the
 * query against groups of cn targets) runs through the same GPU pipeline; the re-mappings of a whole batch share one tagged index.
 * Call order: ndg_step2_begin() once, before the first ndg_index_add(); ndg_step2_map_batch() per query batch (records appended
 * to the stream, the first call also returns the 2-byte header of the 10-field flavour); ndg_step2_finish() for the .bl text.
 * mode: 0 = no re-mapping, 2 = the default re-mapping (modes 1 and 3 are rejected). */
typedef struct { int32_t minlen, maxhan1, maxhan2, minmatch, mode, cn, kn, wn, outctn; float minide; } ndg_step2_opt_t;
int ndg_step2_begin(ndg_ctx_t *ctx, const ndg_step2_opt_t *opt);
int ndg_step2_map_batch(ndg_ctx_t *ctx, const ndg_reads_t *queries, const uint8_t **ovl, uint64_t *n_bytes, uint64_t *n_rec, uint64_t *sum_qspan,
                        ndg_map_result_t *work);
int ndg_step2_finish(ndg_ctx_t *ctx, const char **bl_text, uint64_t *bl_bytes);

/* ---- `ovl_sort`: the step that consumes the .ovl files of the overlap path (ref: util/ovl_sort.c:1-1143, nextDenovo:344-354):
 * decode every 8-field stream, drop self hits and spans < 500, emit each record from both reads' point of view when the read is a
 * seed (seed_name/seed_len, what `-i` lists), sort by (qname, match descending, query span ascending; ties in input order),
 * then per query read: the read's own record first, the depth-capped admission of its overlaps on a 64-base histogram, the
 * chimera / low-coverage / contained verdict (labels `c` and `k` of the .bl file), and the surviving records as one 8-field
 * stream.  Everything between the input bytes and the output bytes runs on the GPU.  Equal to the reference binary whenever that
 * holds its input in memory (one read thread, no spill files: the case of a pipeline block; with spill files the reference
 * filters every chunk separately and once more while merging).  opt: -k (40), -l (300; 0 = no .bl), -H.
 * ovl / bl_text: owned by the context, valid until the next call. */
typedef struct { int32_t max_bin_cov, max_flank_len, is_hq; } ndg_ovlsort_opt_t;
int ndg_ovl_sort(ndg_ctx_t *ctx, const uint8_t *const *streams, const uint64_t *stream_bytes, int32_t n_streams,
                 const uint32_t *seed_name, const uint32_t *seed_len, uint64_t n_seeds, const ndg_ovlsort_opt_t *opt,
                 const uint8_t **ovl, uint64_t *n_bytes, const char **bl_text, uint64_t *bl_bytes, uint64_t *n_rec_in, uint64_t *n_rec_out);

/* ---- sketch only, for tests: replaces mm_sketch (ref: minimap2/sketch.c:358-361) ----------- */
/* returns minimizers of all reads, concatenated; off has n_reads+1 entries */
int ndg_sketch(ndg_ctx_t *ctx, const ndg_reads_t *reads, uint32_t rid0, uint64_t *xy, uint64_t cap, uint64_t *off);

/* ---- introspection ------------------------------------------------------------------------- */
/* copy a stage buffer of the last ndg_map_resident()/ndg_map_batch() sub-batch to the host.
 * what: "mz" (u128), "mz_off" (u64), "anchors" (u128, sorted), "a_off" (u64), "f","p","v" (i32) */
int64_t ndg_debug_fetch(ndg_ctx_t *ctx, const char *what, void *dst, uint64_t cap_bytes);
/* accumulated device time per kernel family since the last reset: names/ms/launches/bytes arrays */
int ndg_prof_get(ndg_ctx_t *ctx, int max, const char **names, double *ms, uint64_t *launches);
int ndg_prof_reset(ndg_ctx_t *ctx);
uint64_t ndg_kernel_launches(const ndg_ctx_t *ctx);

/* ---- .ovl codec (host side, table-free; byte-identical to lib/ovl.c:10-29,109-150) --------- */
int ndg_ovl_encode8(uint8_t *dst, uint32_t state[2], int rev, uint32_t qname, uint32_t qs, uint32_t qe,
                    uint32_t tname, uint32_t ts, uint32_t te, uint32_t match);

/* ---- base-level extension: replaces ksw_extd2_sse (ref: minimap2/ksw2_extd2_sse.c:1-394; called from
 *      mm_align_pair, minimap2/align.c:313-339) for a batch of independent alignments ---------------
 * seqs: concatenated residue codes 0..m-1 (as mm_align1 prepares them); call i aligns
 * query seqs[q_off .. q_off+qlen) to target seqs[t_off .. t_off+tlen).  Results are bit-identical to
 * the reference function (score, end points, z-drop, CIGAR in BAM encoding).  cigar_off: n+1 offsets
 * into cigar[] with cigar_off[i+1]-cigar_off[i] >= qlen+tlen+2. */
typedef struct { int32_t qlen, tlen; uint64_t q_off, t_off; int32_t w, zdrop, end_bonus, flag; } ndg_ksw_call_t;
typedef struct { int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar, reach_end, reserved; } ndg_ksw_result_t;
int ndg_ksw_extd2_batch(ndg_ctx_t *ctx, const ndg_ksw_call_t *calls, int32_t n_calls, const uint8_t *seqs, uint64_t seq_bytes,
                        int32_t m, const int8_t *mat, int32_t q, int32_t e, int32_t q2, int32_t e2,
                        ndg_ksw_result_t *res, uint32_t *cigar, const uint64_t *cigar_off);

/* Stitching batches that were encoded independently (e.g. by several contexts, see ndg_index_share): a batch encoded from
 * state (0,0) starts with absolute names; this rewrites its FIRST 8-field record onto the true previous (qname, tname) of the
 * stream (ref: lib/ovl.c:109-150, the delta/flag rules of encode_ovl).  dst receives the replacement bytes (at most 48),
 * *skip the length of the original first record; returns the replacement length, or <0 on a truncated stream. */
int ndg_ovl_rebase_first(const uint8_t *ovl, uint64_t n_bytes, const uint32_t prev[2], uint8_t *dst, uint32_t *skip);

/* Step-1 emitter: applies the reference's output filter (ref: minimap2/map.c:1297-1304: drop hits whose
 * target is the query itself, keep qe-qs >= minlen, optional --dvt test :804-821) to a sorted hit list
 * and appends 8-field records.  q_id/q_key/q_len are indexed by hit.qidx, t_id/t_key/t_len by hit.rid
 * (ids = strtoul(name)).  Returns bytes written or <0; *n_rec and *sum_qspan are incremented. */
int64_t ndg_ovl_emit_step1(const ndg_hit_t *hits, int64_t n_hits,
                           const uint32_t *q_id, const uint64_t *q_key, const uint32_t *q_len,
                           const uint32_t *t_id, const uint64_t *t_key, const uint32_t *t_len,
                           int32_t minlen, int32_t dvt, int32_t maxhan1, int32_t maxhan2,
                           uint32_t state[2], uint8_t *dst, uint64_t cap, uint64_t *n_rec, uint64_t *sum_qspan);

#ifdef __cplusplus
}
#endif
#endif
