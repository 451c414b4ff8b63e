/*
 * nd_oracle.h -- CPU restatement of NextDenovo's read-overlap hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (nextdenovo_b200/, the C-ABI
 * library, the minimap2-nd-compatible CLI) may include, link or call this.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it, as the
 * checker.
 *
 * Parity pin: this restatement is checked (tests/test_oracle_pin.py) against
 *   (1) the unmodified reference compiled from /root/reference into oracle/_ref
 *       (stage probes through libminimap2_ref.so, whole-file .ovl md5s), and
 *   (2) the golden .ovl fixtures minted from that reference (tests/golden/).
 * The reference itself ships no golden vectors for this path.
 *
 * All "ref:" citations are relative to /root/reference/.
 */
#ifndef ND_ORACLE_H
#define ND_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t x, y; } ndo_u128;            /* ref: minimap2/minimap.h:53 */

typedef struct {                                        /* effective options of the path */
    int k, w, hpc;                                      /* ref: minimap2/options.c:7,90,95 */
    int flag_no_diag, flag_no_dual;                     /* ref: minimap2/minimap.h:8-9 */
    int min_cnt, min_chain_score, bw, max_gap;          /* ref: minimap2/options.c:20-23,92-93 */
    int max_chain_skip, max_chain_iter;                 /* ref: minimap2/options.c:25-26 */
    int seed;                                           /* ref: minimap2/options.c:16 */
    float mid_occ_frac;                                 /* ref: minimap2/options.c:17 */
    int mid_occ;                                        /* 0 = derive from index */
    int minlen;                                         /* ref: minimap2/main.c:193 */
} ndo_opt;

typedef struct {                                        /* subset of mm_reg1_t that reaches .ovl/PAF */
    int32_t rid, rev, qs, qe, rs, re, mlen, blen, score, cnt;
    uint32_t hash;
} ndo_reg;

typedef struct ndo_index ndo_index;

void     ndo_opt_preset(ndo_opt *o, const char *preset); /* NULL, "ava-ont", "ava-pb" */
uint64_t ndo_hash64(uint64_t key, uint64_t mask);

/* seq: ASCII bases. Appends minimizers to out (cap must be >= len+1); returns new count. */
size_t ndo_sketch(const char *seq, int len, int w, int k, uint32_t rid, int hpc, ndo_u128 *out, size_t n0);

void ndo_rs_sort_128x(ndo_u128 *beg, ndo_u128 *end);
void ndo_rs_sort_64(uint64_t *beg, uint64_t *end);

ndo_index *ndo_index_build(int n_seq, const char *const *seqs, const int *lens, const char *const *names,
                           int w, int k, int hpc);
void  ndo_index_free(ndo_index *idx);
const uint64_t *ndo_index_get(const ndo_index *idx, uint64_t minier, int *n);
int   ndo_index_mid_occ(const ndo_index *idx, float frac);
size_t ndo_index_n_keys(const ndo_index *idx);

/* seed collection: returns malloc'd sorted anchors, *n_a set */
ndo_u128 *ndo_collect_seeds(const ndo_index *idx, const ndo_opt *o, int mid_occ, const char *qname, int qlen,
                            const ndo_u128 *mv, size_t n_mv, int64_t *n_a, int sorted);

/* chaining: consumes a (freed), returns compacted anchors + u[] */
ndo_u128 *ndo_chain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
                       int64_t n, ndo_u128 *a, int *n_u, uint64_t **u);

ndo_reg *ndo_gen_regs(uint32_t hash, int qlen, int n_u, const uint64_t *u, const ndo_u128 *a);
uint32_t ndo_query_hash(const char *qname, int qlen, int seed);

/* whole read: returns malloc'd regs in reference output order */
ndo_reg *ndo_map_read(const ndo_index *idx, const ndo_opt *o, int mid_occ, const char *qname, const char *seq, int qlen,
                      int *n_regs);

/* .ovl 8-field encoder (table-free); state = {prev_qname, prev_tname}. returns bytes written */
int ndo_ovl_encode8(uint8_t *dst, uint32_t state[2], int rev, uint32_t qname, uint32_t qs, uint32_t qe,
                    uint32_t tname, uint32_t ts, uint32_t te, uint32_t match);

/* file level: self/cross overlap of .2bit files, step-1 semantics; writes .ovl; returns #records or <0 */
long ndo_step1_files(const char *target_2bit, const char *query_2bit, const ndo_opt *o, const char *out_ovl,
                     uint64_t *sum_qspan_out);

/* .2bit helpers */
typedef struct { int n; char **seq; int *len; char **name; uint32_t *id; } ndo_reads;
ndo_reads *ndo_read_2bit(const char *fn);
void ndo_reads_free(ndo_reads *r);

#ifdef __cplusplus
}
#endif
#endif
