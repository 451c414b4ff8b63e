/*
 * nd_oracle.c -- CPU restatement of NextDenovo's read-overlap hot path (step 1, chain-only regime).
 *
 * TEST INFRASTRUCTURE ONLY (see nd_oracle.h).  Plain C, single-threaded, written for clarity.
 * Every function cites the reference code whose behaviour it restates ("ref:" paths are relative
 * to /root/reference/).  Pinned against the compiled reference by tests/test_oracle_pin.py.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include "nd_oracle.h"

/* ------------------------------------------------------------------------------------------ */
/* options                                                                                     */
/* ------------------------------------------------------------------------------------------ */

/* ref: minimap2/options.c:4-64 (defaults), :89-97 (ava-ont / ava-pb), minimap2/main.c:190-194 (--step 1) */
void ndo_opt_preset(ndo_opt *o, const char *preset)
{
    if (preset == NULL) {
        memset(o, 0, sizeof(*o));
        o->k = 15; o->w = 10; o->hpc = 0;
        o->min_cnt = 3; o->min_chain_score = 40; o->bw = 500; o->max_gap = 5000;
        o->max_chain_skip = 25; o->max_chain_iter = 5000;
        o->seed = 11; o->mid_occ_frac = 2e-4f; o->mid_occ = 0; o->minlen = 500;
        return;
    }
    if (strcmp(preset, "ava-ont") == 0) {
        o->hpc = 0; o->k = 15; o->w = 5;
        o->flag_no_diag = o->flag_no_dual = 1;
        o->min_chain_score = 100; o->max_gap = 10000; o->max_chain_skip = 25; o->bw = 2000;
    } else if (strcmp(preset, "ava-pb") == 0) {
        o->hpc = 1; o->k = 19; o->w = 5;
        o->flag_no_diag = o->flag_no_dual = 1;
        o->min_chain_score = 100; o->max_gap = 10000; o->max_chain_skip = 25;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* minimizer sketch                                                                            */
/* ------------------------------------------------------------------------------------------ */

/* ref: minimap2/sketch.c:28-38 -- invertible integer hash on 2k bits */
uint64_t ndo_hash64(uint64_t key, uint64_t mask)
{
    key = (~key + (key << 21)) & mask;
    key ^= key >> 24;
    key = (key + (key << 3) + (key << 8)) & mask;
    key ^= key >> 14;
    key = (key + (key << 2) + (key << 4)) & mask;
    key ^= key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

static inline int base_code(char ch) /* ref: minimap2/sketch.c:9-26 (seq_nt4_table) */
{
    switch (ch) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    case 0: return 0; case 1: return 1; case 2: return 2; case 3: return 3; /* raw codes */
    default: return 4;
    }
}

/*
 * ref: minimap2/sketch.c:77-143 (mm_sketch_shortkmer).  State machine over "slots": every position
 * except those producing a strand-symmetric k-mer occupies one slot of a w-entry ring.  `cur` is
 * always the right-most minimum of the ring.  Emission rules (H4 of SURVEY.md):
 *   - a candidate that is <= cur replaces it; the displaced cur is emitted once l >= w+k;
 *   - when cur falls out of the window it is emitted (l >= w+k-1), the ring is rescanned with >=
 *     (right-most minimum wins) and all other ring entries equal to the new minimum are emitted;
 *   - the first full window (l == w+k-1) emits ring entries equal to cur;
 *   - the final cur is always flushed.
 */
/* long k-mers (k > 28, the ava-hifi preset: k = 51).  ref: minimap2/sketch.c:215-281: the k-mer is 2k bits in four 64-bit words (word 0
 * = the newest bases), strands compared as 256-bit integers, hash = hash64 of the top (partial) word under its mask plus the unmasked
 * hash of every lower word that is not zero. */
typedef struct { uint64_t u[4]; } ndo_u256;
static uint64_t ndo_hash64_nomask(uint64_t key)
{
    key = (~key + (key << 21));
    key ^= key >> 24;
    key = (key + (key << 3) + (key << 8));
    key ^= key >> 14;
    key = (key + (key << 2) + (key << 4));
    key ^= key >> 28;
    key = (key + (key << 31));
    return key;
}

size_t ndo_sketch(const char *seq, int len, int w, int k, uint32_t rid, int hpc, ndo_u128 *out, size_t n)
{
    const int longk = k > 28;
    const int k_idx = (k - 1) / 32, lshift = ((k - 1) & 31) << 1, top_bits = (((k - 1) & 31) + 1) << 1;
    /* k a multiple of 32: the reference shifts 1ULL by 64, which x86 executes as a shift by 0: the mask is 0 and the top word is wiped (kept) */
    const uint64_t mask = longk ? (1ULL << (top_bits & 63)) - 1 : (1ULL << 2 * k) - 1, rc_shift = longk ? (uint64_t)lshift : 2 * (k - 1);
    const ndo_u128 none = { UINT64_MAX, UINT64_MAX };
    uint64_t fw = 0, rv = 0;
    ndo_u256 F = { { 0, 0, 0, 0 } }, R = { { 0, 0, 0, 0 } };
    ndo_u128 ring[256], cur = none;
    int runq[32], rq_head = 0, rq_n = 0;       /* homopolymer run lengths: a 32-entry ring, as the reference's tiny_queue_t -- for k > 31 it
                                                * wraps before an entry is popped, so `span` is NOT the sum of the last k runs there (kept) */
    int slot = 0, cur_slot = 0, l = 0, span = 0, i, j;

    assert(len > 0 && w > 0 && w < 256 && k > 0 && k <= 128);
    for (j = 0; j < w; ++j) ring[j] = none;

    for (i = 0; i < len; ++i) {
        int c = base_code(seq[i]);
        ndo_u128 cand = none;
        if (c < 4) {
            int strand;
            if (hpc) {
                int run = 1;
                while (i + run < len && base_code(seq[i + run]) == c) ++run;
                i += run - 1;                               /* i now = last base of the run */
                runq[(rq_head + rq_n++) & 31] = run;
                span += run;
                if (rq_n > k) { span -= runq[rq_head]; rq_head = (rq_head + 1) & 31; --rq_n; }
            } else span = l + 1 < k ? l + 1 : k;
            if (!longk) {
                fw = (fw << 2 | (uint64_t)c) & mask;
                rv = (rv >> 2) | (uint64_t)(3 ^ c) << rc_shift;
                if (fw == rv) continue;                     /* symmetric k-mer: no slot consumed */
                strand = fw < rv ? 0 : 1;
            } else {
                int cmp = 0, q;
                F.u[3] = F.u[3] << 2 | F.u[2] >> 62; F.u[2] = F.u[2] << 2 | F.u[1] >> 62; F.u[1] = F.u[1] << 2 | F.u[0] >> 62; F.u[0] = F.u[0] << 2 | (uint64_t)c;
                F.u[k_idx] &= mask;
                R.u[0] = R.u[0] >> 2 | R.u[1] << 62; R.u[1] = R.u[1] >> 2 | R.u[2] << 62; R.u[2] = R.u[2] >> 2 | R.u[3] << 62; R.u[3] >>= 2;
                R.u[k_idx] |= (uint64_t)(3 ^ c) << rc_shift;
                for (q = 3; q >= 0 && !cmp; --q) cmp = F.u[q] < R.u[q] ? -1 : F.u[q] > R.u[q] ? 1 : 0;
                if (cmp == 0) continue;
                strand = cmp < 0 ? 0 : 1;
            }
            ++l;
            if (l >= k && span < 256) {
                uint64_t h;
                if (!longk) h = ndo_hash64(strand ? rv : fw, mask);
                else {
                    const ndo_u256 *K = strand ? &R : &F;
                    h = ndo_hash64(K->u[k_idx], mask);
                    if (k_idx > 2 && K->u[2]) h += ndo_hash64_nomask(K->u[2]);
                    if (k_idx > 1 && K->u[1]) h += ndo_hash64_nomask(K->u[1]);
                    if (k_idx > 0 && K->u[0]) h += ndo_hash64_nomask(K->u[0]);
                }
                cand.x = h << 8 | (uint64_t)span;
                cand.y = (uint64_t)rid << 32 | (uint32_t)i << 1 | (uint64_t)strand;
            }
        } else { l = 0; rq_n = rq_head = 0; span = 0; }
        ring[slot] = cand;
        if (l == w + k - 1 && cur.x != UINT64_MAX) {        /* first full window: ties of cur */
            for (j = slot + 1; j < w; ++j) if (ring[j].x == cur.x && ring[j].y != cur.y) out[n++] = ring[j];
            for (j = 0; j < slot; ++j)     if (ring[j].x == cur.x && ring[j].y != cur.y) out[n++] = ring[j];
        }
        if (cand.x <= cur.x) {
            if (l >= w + k && cur.x != UINT64_MAX) out[n++] = cur;
            cur = cand; cur_slot = slot;
        } else if (slot == cur_slot) {
            if (l >= w + k - 1 && cur.x != UINT64_MAX) out[n++] = cur;
            cur.x = UINT64_MAX;
            for (j = slot + 1; j < w; ++j) if (cur.x >= ring[j].x) { cur = ring[j]; cur_slot = j; }
            for (j = 0; j <= slot; ++j)    if (cur.x >= ring[j].x) { cur = ring[j]; cur_slot = j; }
            if (l >= w + k - 1 && cur.x != UINT64_MAX) {
                for (j = slot + 1; j < w; ++j) if (ring[j].x == cur.x && ring[j].y != cur.y) out[n++] = ring[j];
                for (j = 0; j <= slot; ++j)    if (ring[j].x == cur.x && ring[j].y != cur.y) out[n++] = ring[j];
            }
        }
        if (++slot == w) slot = 0;
    }
    if (cur.x != UINT64_MAX) out[n++] = cur;
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* in-place MSD radix sort ("American flag"), unstable -- tie order is observable    */
/* ------------------------------------------------------------------------------------------ */

/* ref: minimap2/ksort.h:98-151 (KRADIX_SORT_INIT), instantiated minimap2/misc.c:155-159 */
#define NDO_RS_SMALL 64

static void ins_sort_128x(ndo_u128 *beg, ndo_u128 *end)
{
    ndo_u128 *p, *q;
    for (p = beg + 1; p < end; ++p) {
        if (p->x < (p - 1)->x) {
            ndo_u128 t = *p;
            for (q = p; q > beg && t.x < (q - 1)->x; --q) *q = *(q - 1);
            *q = t;
        }
    }
}

static void flag_sort_128x(ndo_u128 *beg, ndo_u128 *end, int shift)
{
    ndo_u128 *head[256], *tail[256], *p;
    size_t cnt[256];
    int d;
    memset(cnt, 0, sizeof(cnt));
    for (p = beg; p != end; ++p) ++cnt[p->x >> shift & 255];
    for (d = 0, p = beg; d < 256; ++d) { head[d] = p; p += cnt[d]; tail[d] = p; }
    for (d = 0; d < 256;) {                                  /* cycle-leader permutation */
        int t;
        if (head[d] == tail[d]) { ++d; continue; }
        t = (int)(head[d]->x >> shift & 255);
        if (t == d) { ++head[d]; continue; }
        {
            ndo_u128 carry = *head[d];
            do {
                ndo_u128 evicted = *head[t];
                *head[t]++ = carry;
                carry = evicted;
                t = (int)(carry.x >> shift & 255);
            } while (t != d);
            *head[d]++ = carry;
        }
    }
    if (shift) {
        int next = shift > 8 ? shift - 8 : 0;
        for (d = 0; d < 256; ++d) {
            ndo_u128 *b = d ? tail[d - 1] : beg, *e = tail[d];
            if (e - b > NDO_RS_SMALL) flag_sort_128x(b, e, next);
            else if (e - b > 1) ins_sort_128x(b, e);
        }
    }
}

void ndo_rs_sort_128x(ndo_u128 *beg, ndo_u128 *end)
{
    if (end - beg <= NDO_RS_SMALL) ins_sort_128x(beg, end);
    else flag_sort_128x(beg, end, 56);
}

static void ins_sort_64(uint64_t *beg, uint64_t *end)
{
    uint64_t *p, *q;
    for (p = beg + 1; p < end; ++p)
        if (*p < *(p - 1)) {
            uint64_t t = *p;
            for (q = p; q > beg && t < *(q - 1); --q) *q = *(q - 1);
            *q = t;
        }
}

static void flag_sort_64(uint64_t *beg, uint64_t *end, int shift)
{
    uint64_t *head[256], *tail[256], *p;
    size_t cnt[256];
    int d;
    memset(cnt, 0, sizeof(cnt));
    for (p = beg; p != end; ++p) ++cnt[*p >> shift & 255];
    for (d = 0, p = beg; d < 256; ++d) { head[d] = p; p += cnt[d]; tail[d] = p; }
    for (d = 0; d < 256;) {
        int t;
        if (head[d] == tail[d]) { ++d; continue; }
        t = (int)(*head[d] >> shift & 255);
        if (t == d) { ++head[d]; continue; }
        {
            uint64_t carry = *head[d];
            do {
                uint64_t evicted = *head[t];
                *head[t]++ = carry;
                carry = evicted;
                t = (int)(carry >> shift & 255);
            } while (t != d);
            *head[d]++ = carry;
        }
    }
    if (shift) {
        int next = shift > 8 ? shift - 8 : 0;
        for (d = 0; d < 256; ++d) {
            uint64_t *b = d ? tail[d - 1] : beg, *e = tail[d];
            if (e - b > NDO_RS_SMALL) flag_sort_64(b, e, next);
            else if (e - b > 1) ins_sort_64(b, e);
        }
    }
}

void ndo_rs_sort_64(uint64_t *beg, uint64_t *end)
{
    if (end - beg <= NDO_RS_SMALL) ins_sort_64(beg, end);
    else flag_sort_64(beg, end, 56);
}

/* ------------------------------------------------------------------------------------------ */
/* index: the reference's 2^14 buckets + khash hold exactly a map  hash -> ascending positions  */
/* ------------------------------------------------------------------------------------------ */

struct ndo_index {
    int w, k, hpc, n_seq;
    char **name;
    int *len;
    size_t n_keys, n_pos;
    uint64_t *key;      /* distinct minimizer hashes, ascending */
    uint64_t *off;      /* n_keys+1 offsets into pos[] */
    uint64_t *pos;      /* rid<<32 | lastPos<<1 | strand, ascending within a key (ref: index.c:230) */
};

static int cmp_u128_xy(const void *a, const void *b)
{
    const ndo_u128 *p = (const ndo_u128*)a, *q = (const ndo_u128*)b;
    uint64_t kp = p->x >> 8, kq = q->x >> 8;
    if (kp != kq) return kp < kq ? -1 : 1;
    return p->y < q->y ? -1 : p->y > q->y;
}

/* ref: minimap2/index.c:280-372 (mm_idx_gen: sketch every target, bucket), :191-243 (worker_post:
 * group equal minimizers, positions sorted ascending) */
ndo_index *ndo_index_build(int n_seq, const char *const *seqs, const int *lens, const char *const *names,
                           int w, int k, int hpc)
{
    ndo_index *idx = (ndo_index*)calloc(1, sizeof(*idx));
    size_t cap = 0, n = 0, i, j;
    ndo_u128 *mz;
    int s;
    idx->w = w; idx->k = k; idx->hpc = hpc; idx->n_seq = n_seq;
    idx->name = (char**)calloc(n_seq, sizeof(char*));
    idx->len = (int*)calloc(n_seq, sizeof(int));
    for (s = 0; s < n_seq; ++s) cap += (size_t)lens[s] + 1;
    mz = (ndo_u128*)malloc(cap * sizeof(ndo_u128));
    for (s = 0; s < n_seq; ++s) {
        idx->name[s] = strdup(names[s]);
        idx->len[s] = lens[s];
        if (lens[s] > 0) n = ndo_sketch(seqs[s], lens[s], w, k, (uint32_t)s, hpc, mz, n);
    }
    qsort(mz, n, sizeof(ndo_u128), cmp_u128_xy);
    idx->pos = (uint64_t*)malloc((n ? n : 1) * 8);
    for (i = 0, idx->n_keys = 0; i < n; ++i)
        if (i == 0 || mz[i].x >> 8 != mz[i - 1].x >> 8) ++idx->n_keys;
    idx->key = (uint64_t*)malloc((idx->n_keys ? idx->n_keys : 1) * 8);
    idx->off = (uint64_t*)malloc((idx->n_keys + 1) * 8);
    for (i = 0, j = 0; i < n; ++i) {
        if (i == 0 || mz[i].x >> 8 != mz[i - 1].x >> 8) { idx->key[j] = mz[i].x >> 8; idx->off[j] = i; ++j; }
        idx->pos[i] = mz[i].y;
    }
    idx->off[idx->n_keys] = n;
    idx->n_pos = n;
    free(mz);
    return idx;
}

void ndo_index_free(ndo_index *idx)
{
    int s;
    if (!idx) return;
    for (s = 0; s < idx->n_seq; ++s) free(idx->name[s]);
    free(idx->name); free(idx->len); free(idx->key); free(idx->off); free(idx->pos); free(idx);
}

size_t ndo_index_n_keys(const ndo_index *idx) { return idx->n_keys; }

/* ref: minimap2/index.c:81-98 (mm_idx_get) */
const uint64_t *ndo_index_get(const ndo_index *idx, uint64_t minier, int *n)
{
    size_t lo = 0, hi = idx->n_keys;
    *n = 0;
    while (lo < hi) {
        size_t mid = lo + (hi - lo) / 2;
        if (idx->key[mid] < minier) lo = mid + 1; else hi = mid;
    }
    if (lo == idx->n_keys || idx->key[lo] != minier) return NULL;
    *n = (int)(idx->off[lo + 1] - idx->off[lo]);
    return idx->pos + idx->off[lo];
}

static int cmp_u32(const void *a, const void *b)
{
    uint32_t p = *(const uint32_t*)a, q = *(const uint32_t*)b;
    return p < q ? -1 : p > q;
}

/* ref: minimap2/index.c:164-185 (mm_idx_cal_max_occ): order statistic of per-minimizer counts */
int ndo_index_mid_occ(const ndo_index *idx, float frac)
{
    size_t n = idx->n_keys, i;
    uint32_t *cnt, kth;
    int r;
    if (frac <= 0.) return INT32_MAX;
    cnt = (uint32_t*)malloc((n ? n : 1) * 4);
    for (i = 0; i < n; ++i) cnt[i] = (uint32_t)(idx->off[i + 1] - idx->off[i]);
    qsort(cnt, n, 4, cmp_u32);
    kth = (uint32_t)((1. - frac) * n);
    r = (int)(cnt[kth] + 1);
    free(cnt);
    return r;
}

/* ------------------------------------------------------------------------------------------ */
/* seed collection                                                                              */
/* ------------------------------------------------------------------------------------------ */

#define NDO_SEED_TANDEM (1ULL << 42)   /* ref: minimap2/mmpriv.h:20 */
#define NDO_SEED_SELF   (1ULL << 43)   /* ref: minimap2/mmpriv.h:21 */

/* ref: minimap2/map.c:91-124 (collect_matches), :126-148 (skip_seed), :216-248 (collect_seed_hits) */
ndo_u128 *ndo_collect_seeds(const ndo_index *idx, const ndo_opt *o, int mid_occ, const char *qname, int qlen,
                            const ndo_u128 *mv, size_t n_mv, int64_t *n_a, int sorted)
{
    size_t i, cap = 0;
    int64_t n = 0;
    ndo_u128 *a;
    for (i = 0; i < n_mv; ++i) {
        int t; ndo_index_get(idx, mv[i].x >> 8, &t);
        if (t < mid_occ) cap += t;
    }
    a = (ndo_u128*)malloc((cap ? cap : 1) * sizeof(ndo_u128));
    for (i = 0; i < n_mv; ++i) {
        int t, kk, tandem = 0;
        const uint64_t *cr = ndo_index_get(idx, mv[i].x >> 8, &t);
        uint32_t q_pos = (uint32_t)mv[i].y, q_span = (uint32_t)(mv[i].x & 0xff);
        if (t >= mid_occ) continue;                        /* repetitive minimizer (rep_len not needed for .ovl) */
        if (i > 0 && mv[i].x >> 8 == mv[i - 1].x >> 8) tandem = 1;
        if (i + 1 < n_mv && mv[i].x >> 8 == mv[i + 1].x >> 8) tandem = 1;
        for (kk = 0; kk < t; ++kk) {
            uint64_t r = cr[kk], x, y;
            int is_self = 0;
            int32_t rpos = (int32_t)((uint32_t)r >> 1);
            if (qname && (o->flag_no_diag || o->flag_no_dual)) {
                int rid = (int)(r >> 32);
                int cmp = strcmp(qname, idx->name[rid]);
                if (o->flag_no_diag && cmp == 0 && idx->len[rid] == qlen) {
                    if ((uint32_t)r >> 1 == q_pos >> 1) continue;          /* exact diagonal */
                    if ((r & 1) == (q_pos & 1)) is_self = 1;
                }
                if (o->flag_no_dual && cmp > 0) continue;
            }
            if ((r & 1) == (q_pos & 1)) {                  /* same strand */
                x = (r & 0xffffffff00000000ULL) | (uint32_t)rpos;
                y = (uint64_t)q_span << 32 | q_pos >> 1;
            } else {
                x = 1ULL << 63 | (r & 0xffffffff00000000ULL) | (uint32_t)rpos;
                y = (uint64_t)q_span << 32 | (uint32_t)(qlen - (int32_t)((q_pos >> 1) + 1 - q_span) - 1);
            }
            if (tandem) y |= NDO_SEED_TANDEM;
            if (is_self) y |= NDO_SEED_SELF;
            a[n].x = x; a[n].y = y; ++n;
        }
    }
    if (sorted) ndo_rs_sort_128x(a, a + n);
    *n_a = n;
    return a;
}

/* ------------------------------------------------------------------------------------------ */
/* chaining                                                                                     */
/* ------------------------------------------------------------------------------------------ */

static inline int ilog2_u32(uint32_t v) /* ref: minimap2/chain.c:8-20; v > 0 */
{
    int r = 0;
    while (v >>= 1) ++r;
    return r;
}

/*
 * ref: minimap2/chain.c:22-162 (mm_chain_dp), single-segment non-cDNA case (the only one reachable
 * from --step 1 ava presets).  Consumes `a`; returns the compacted anchors of all kept chains and
 * *u (score<<32 | count per chain), both ordered by the first anchor's x (the final re-sort).
 */
ndo_u128 *ndo_chain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
                       int64_t n, ndo_u128 *a, int *n_u_, uint64_t **u_)
{
    int32_t *f, *p, *t, *v;
    int64_t i, j, st = 0;
    uint64_t sum_span = 0, *u, *u2;
    float avg_span;
    int32_t n_u, n_v, kept;
    ndo_u128 *b, *wv;

    *u_ = NULL; *n_u_ = 0;
    if (n == 0 || a == NULL) { free(a); return NULL; }
    f = (int32_t*)malloc(n * 4); p = (int32_t*)malloc(n * 4);
    t = (int32_t*)calloc(n, 4);  v = (int32_t*)malloc(n * 4);
    for (i = 0; i < n; ++i) sum_span += a[i].y >> 32 & 0xff;
    avg_span = (float)sum_span / n;

    for (i = 0; i < n; ++i) {                               /* forward DP (chain.c:45-86) */
        uint64_t ri = a[i].x;
        int32_t qi = (int32_t)a[i].y, span = (int32_t)(a[i].y >> 32 & 0xff);
        int32_t best = span, n_skip = 0;
        int64_t best_j = -1;
        while (st < i && ri > a[st].x + (uint64_t)max_dist_x) ++st;
        if (i - st > max_iter) st = i - max_iter;
        for (j = i - 1; j >= st; --j) {
            int64_t dr = (int64_t)(ri - a[j].x);
            int32_t dq = qi - (int32_t)a[j].y, dd, sc, lg;
            if (dr == 0 || dq <= 0) continue;
            if (dq > max_dist_y || dq > max_dist_x) continue;
            dd = (int32_t)(dr > dq ? dr - dq : dq - dr);
            if (dd > bw) continue;
            sc = (dq < dr ? dq : (int32_t)dr);
            if (sc > span) sc = span;
            lg = dd ? ilog2_u32((uint32_t)dd) : 0;
            sc -= (int)(dd * .01 * avg_span) + (lg >> 1);
            sc += f[j];
            if (sc > best) {
                best = sc; best_j = j;
                if (n_skip > 0) --n_skip;
            } else if (t[j] == (int32_t)i) {
                if (++n_skip > max_skip) break;
            }
            if (p[j] >= 0) t[p[j]] = (int32_t)i;
        }
        f[i] = best; p[i] = (int32_t)best_j;
        v[i] = best_j >= 0 && v[best_j] > best ? v[best_j] : best;
    }

    memset(t, 0, n * 4);                                    /* chain ends (chain.c:88-106) */
    for (i = 0; i < n; ++i) if (p[i] >= 0) t[p[i]] = 1;
    for (i = 0, n_u = 0; i < n; ++i) if (t[i] == 0 && v[i] >= min_sc) ++n_u;
    if (n_u == 0) { free(a); free(f); free(p); free(t); free(v); return NULL; }
    u = (uint64_t*)malloc(n_u * 8);
    for (i = 0, n_u = 0; i < n; ++i)
        if (t[i] == 0 && v[i] >= min_sc) {
            j = i;
            while (j >= 0 && f[j] < v[j]) j = p[j];
            if (j < 0) j = i;
            u[n_u++] = (uint64_t)f[j] << 32 | (uint64_t)j;
        }
    ndo_rs_sort_64(u, u + n_u);
    for (i = 0; i < n_u >> 1; ++i) { uint64_t z = u[i]; u[i] = u[n_u - 1 - i]; u[n_u - 1 - i] = z; }

    memset(t, 0, n * 4);                                    /* backtrack (chain.c:113-130) */
    for (i = 0, n_v = 0, kept = 0; i < n_u; ++i) {
        int32_t n_v0 = n_v, kept0 = kept;
        j = (int32_t)u[i];
        do { v[n_v++] = (int32_t)j; t[j] = 1; j = p[j]; } while (j >= 0 && t[j] == 0);
        if (j < 0) {
            if (n_v - n_v0 >= min_cnt) u[kept++] = u[i] >> 32 << 32 | (uint32_t)(n_v - n_v0);
        } else if ((int32_t)(u[i] >> 32) - f[j] >= min_sc) {
            if (n_v - n_v0 >= min_cnt) u[kept++] = ((u[i] >> 32) - (uint64_t)f[j]) << 32 | (uint32_t)(n_v - n_v0);
        }
        if (kept == kept0) n_v = n_v0;
    }
    n_u = kept;
    free(f); free(p); free(t);

    b = (ndo_u128*)malloc((n_v ? n_v : 1) * sizeof(ndo_u128)); /* chain.c:137-142 */
    for (i = 0, kept = 0; i < n_u; ++i) {
        int32_t k0 = kept, ni = (int32_t)u[i];
        for (j = 0; j < ni; ++j) b[kept++] = a[v[k0 + (ni - j - 1)]];
    }
    free(v);

    wv = (ndo_u128*)malloc((n_u ? n_u : 1) * sizeof(ndo_u128)); /* chain.c:145-160: order chains by target */
    for (i = 0, kept = 0; i < n_u; ++i) { wv[i].x = b[kept].x; wv[i].y = (uint64_t)kept << 32 | (uint64_t)i; kept += (int32_t)u[i]; }
    ndo_rs_sort_128x(wv, wv + n_u);
    u2 = (uint64_t*)malloc((n_u ? n_u : 1) * 8);
    for (i = 0, kept = 0; i < n_u; ++i) {
        int32_t src = (int32_t)wv[i].y, cnt = (int32_t)u[src];
        u2[i] = u[src];
        memcpy(&a[kept], &b[wv[i].y >> 32], cnt * sizeof(ndo_u128));
        kept += cnt;
    }
    memcpy(b, a, kept * sizeof(ndo_u128));
    free(a); free(wv); free(u);
    *n_u_ = n_u; *u_ = u2;
    return b;
}

/* ------------------------------------------------------------------------------------------ */
/* chains -> hits                                                                               */
/* ------------------------------------------------------------------------------------------ */

static inline uint64_t mix64(uint64_t key) /* ref: minimap2/hit.c:40-50 */
{
    key = ~key + (key << 21);
    key ^= key >> 24;
    key = key + (key << 3) + (key << 8);
    key ^= key >> 14;
    key = key + (key << 2) + (key << 4);
    key ^= key >> 28;
    key = key + (key << 31);
    return key;
}

static inline uint32_t wang32(uint32_t key) /* ref: minimap2/khash.h:400-409 */
{
    key += ~(key << 15); key ^= (key >> 10); key += (key << 3);
    key ^= (key >> 6);   key += ~(key << 11); key ^= (key >> 16);
    return key;
}

/* ref: minimap2/map.c:525-527, minimap2/khash.h:383-388 */
uint32_t ndo_query_hash(const char *qname, int qlen, int seed)
{
    uint32_t h = 0;
    if (qname) {
        const char *s = qname;
        h = (uint32_t)*s;
        if (h) for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)*s;
    }
    h ^= wang32((uint32_t)qlen) + wang32((uint32_t)seed);
    return wang32(h);
}

/* ref: minimap2/hit.c:52-88 (mm_gen_regs), :23-38 (mm_reg_set_coor), :8-21 (mm_cal_fuzzy_len) */
ndo_reg *ndo_gen_regs(uint32_t hash, int qlen, int n_u, const uint64_t *u, const ndo_u128 *a)
{
    ndo_u128 *z;
    ndo_reg *r;
    int i, k;
    if (n_u == 0) return NULL;
    z = (ndo_u128*)malloc(n_u * sizeof(ndo_u128));
    for (i = 0, k = 0; i < n_u; ++i) {
        uint32_t h = (uint32_t)mix64((mix64(a[k].x) + mix64(a[k].y)) ^ hash);
        z[i].x = u[i] ^ h;
        z[i].y = (uint64_t)k << 32 | (uint32_t)(int32_t)u[i];
        k += (int32_t)u[i];
    }
    ndo_rs_sort_128x(z, z + n_u);
    for (i = 0; i < n_u >> 1; ++i) { ndo_u128 t = z[i]; z[i] = z[n_u - 1 - i]; z[n_u - 1 - i] = t; }
    r = (ndo_reg*)calloc(n_u, sizeof(ndo_reg));
    for (i = 0; i < n_u; ++i) {
        ndo_reg *ri = &r[i];
        int cnt = (int32_t)z[i].y, as = (int)(z[i].y >> 32), j;
        int32_t span0 = (int32_t)(a[as].y >> 32 & 0xff);
        ri->score = (int32_t)(z[i].x >> 32);
        ri->hash = (uint32_t)z[i].x;
        ri->cnt = cnt;
        ri->rev = (int)(a[as].x >> 63);
        ri->rid = (int32_t)(a[as].x << 1 >> 33);
        ri->rs = (int32_t)a[as].x + 1 > span0 ? (int32_t)a[as].x + 1 - span0 : 0;
        ri->re = (int32_t)a[as + cnt - 1].x + 1;
        if (!ri->rev) {
            ri->qs = (int32_t)a[as].y + 1 - span0;
            ri->qe = (int32_t)a[as + cnt - 1].y + 1;
        } else {
            ri->qs = qlen - ((int32_t)a[as + cnt - 1].y + 1);
            ri->qe = qlen - ((int32_t)a[as].y + 1 - span0);
        }
        ri->mlen = ri->blen = span0;
        for (j = as + 1; j < as + cnt; ++j) {
            int span = (int)(a[j].y >> 32 & 0xff);
            int tl = (int32_t)a[j].x - (int32_t)a[j - 1].x;
            int ql = (int32_t)a[j].y - (int32_t)a[j - 1].y;
            ri->blen += tl > ql ? tl : ql;
            ri->mlen += tl > span && ql > span ? span : tl < ql ? tl : ql;
        }
    }
    free(z);
    return r;
}

/* ------------------------------------------------------------------------------------------ */
/* one read end to end                                                                          */
/* ------------------------------------------------------------------------------------------ */

/* ref: minimap2/map.c:507-623 (mm_map_frag), n_segs == 1, no CIGAR; est_err/mapq skipped (not emitted) */
ndo_reg *ndo_map_read(const ndo_index *idx, const ndo_opt *o, int mid_occ, const char *qname, const char *seq, int qlen,
                      int *n_regs)
{
    ndo_u128 *mv, *a;
    size_t n_mv;
    int64_t n_a;
    uint64_t *u = NULL;
    int n_u = 0;
    ndo_reg *r;
    *n_regs = 0;
    if (qlen == 0) return NULL;
    mv = (ndo_u128*)malloc(((size_t)qlen + 1) * sizeof(ndo_u128));
    n_mv = ndo_sketch(seq, qlen, idx->w, idx->k, 0, idx->hpc, mv, 0);
    a = ndo_collect_seeds(idx, o, mid_occ, qname, qlen, mv, n_mv, &n_a, 1);
    a = ndo_chain_dp(o->max_gap, o->max_gap, o->bw, o->max_chain_skip, o->max_chain_iter, o->min_cnt,
                     o->min_chain_score, n_a, a, &n_u, &u);
    r = ndo_gen_regs(ndo_query_hash(qname, qlen, o->seed), qlen, n_u, u, a);
    *n_regs = n_u;
    free(mv); free(a); free(u);
    return r;
}

/* ------------------------------------------------------------------------------------------ */
/* .ovl encoder                                                                                 */
/* ------------------------------------------------------------------------------------------ */

static inline int put_varint(uint8_t *dst, uint32_t v) /* ref: lib/ovl.c:10-29 (table) == :133-143 (direct) */
{
    int m = 0, sh;
    if (v <= 127) { dst[0] = (uint8_t)v; return 1; }
    for (sh = 28; sh >= 0; sh -= 7) {
        uint32_t g = v >> sh & 127;
        if (g > 0 || m > 0) dst[m++] = (uint8_t)(g | 128);
    }
    dst[m - 1] &= 127;
    return m;
}

/* ref: lib/ovl.c:109-150 (encode_ovl) */
int ndo_ovl_encode8(uint8_t *dst, uint32_t state[2], int rev, uint32_t qname, uint32_t qs, uint32_t qe,
                    uint32_t tname, uint32_t ts, uint32_t te, uint32_t match)
{
    uint32_t f[8], qspan = qe - qs, tspan = te - ts;
    uint8_t flags = (uint8_t)rev;
    int i, n = 0;
    if (qname >= state[0]) f[0] = qname - state[0]; else { flags |= 2; f[0] = state[0] - qname; }
    state[0] = qname;
    if (tname >= state[1]) f[4] = tname - state[1]; else { flags |= 4; f[4] = state[1] - tname; }
    state[1] = tname;
    if (qspan >= tspan) f[6] = qspan - tspan; else { flags |= 8; f[6] = tspan - qspan; }
    f[1] = flags; f[2] = qs; f[3] = qspan; f[5] = ts; f[7] = match;
    for (i = 0; i < 8; ++i) n += put_varint(dst + n, f[i]);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* .2bit reader + file-level step 1                                                             */
/* ------------------------------------------------------------------------------------------ */

/* ref: lib/bseq.c:93-97 (header), :114-139 (seq2bit), :257-302 (kbit_read) */
ndo_reads *ndo_read_2bit(const char *fn)
{
    FILE *fp = fopen(fn, "rb");
    ndo_reads *r;
    uint8_t hdr[2];
    int cap = 0;
    if (!fp) return NULL;
    if (fread(hdr, 1, 2, fp) != 2 || !(hdr[0] == 0 || hdr[0] == 254)) { fclose(fp); return NULL; }
    r = (ndo_reads*)calloc(1, sizeof(*r));
    for (;;) {
        uint32_t name, len, nw, i, *wbuf;
        char *s, nb[16];
        if (fread(&name, 4, 1, fp) != 1) break;
        if (fread(&len, 4, 1, fp) != 1) break;
        nw = len ? ((len - 1) >> 4) + 1 : 0;
        wbuf = (uint32_t*)malloc((nw ? nw : 1) * 4);
        if (fread(wbuf, 4, nw, fp) != nw) { free(wbuf); break; }
        s = (char*)malloc((size_t)nw * 16 + 1);
        for (i = 0; i < nw * 16; ++i) s[i] = "ACGT"[wbuf[i >> 4] >> (30 - 2 * (i & 15)) & 3];
        s[len] = 0;
        free(wbuf);
        if (r->n == cap) {
            cap = cap ? cap * 2 : 1024;
            r->seq = (char**)realloc(r->seq, cap * sizeof(char*));
            r->len = (int*)realloc(r->len, cap * sizeof(int));
            r->name = (char**)realloc(r->name, cap * sizeof(char*));
            r->id = (uint32_t*)realloc(r->id, cap * sizeof(uint32_t));
        }
        snprintf(nb, sizeof nb, "%u", name);
        r->seq[r->n] = s; r->len[r->n] = (int)len; r->name[r->n] = strdup(nb); r->id[r->n] = name;
        ++r->n;
    }
    fclose(fp);
    return r;
}

void ndo_reads_free(ndo_reads *r)
{
    int i;
    if (!r) return;
    for (i = 0; i < r->n; ++i) { free(r->seq[i]); free(r->name[i]); }
    free(r->seq); free(r->len); free(r->name); free(r->id); free(r);
}

/* ref: minimap2/main.c:477-509 (single index part), minimap2/map.c:1297-1304 (step-1 filter + emit) */
long ndo_step1_files(const char *target_2bit, const char *query_2bit, const ndo_opt *o, const char *out_ovl,
                     uint64_t *sum_qspan_out)
{
    ndo_reads *T = ndo_read_2bit(target_2bit), *Q;
    ndo_index *idx;
    FILE *fo;
    int mid_occ, i, j;
    long n_rec = 0;
    uint64_t sum_qspan = 0;
    uint32_t st[2] = { 0, 0 };
    if (!T) return -1;
    Q = strcmp(target_2bit, query_2bit) == 0 ? T : ndo_read_2bit(query_2bit);
    if (!Q) { ndo_reads_free(T); return -1; }
    idx = ndo_index_build(T->n, (const char* const*)T->seq, T->len, (const char* const*)T->name, o->w, o->k, o->hpc);
    mid_occ = o->mid_occ > 0 ? o->mid_occ : ndo_index_mid_occ(idx, o->mid_occ_frac);
    fo = out_ovl ? fopen(out_ovl, "wb") : NULL;
    for (i = 0; i < Q->n; ++i) {
        int n_regs;
        ndo_reg *r = ndo_map_read(idx, o, mid_occ, Q->name[i], Q->seq[i], Q->len[i], &n_regs);
        for (j = 0; j < n_regs; ++j) {
            uint8_t buf[48];
            int nb;
            if (strcmp(Q->name[i], idx->name[r[j].rid]) == 0) continue;
            if (r[j].qe - r[j].qs < o->minlen) continue;
            nb = ndo_ovl_encode8(buf, st, r[j].rev, (uint32_t)strtoul(Q->name[i], NULL, 10), r[j].qs, r[j].qe,
                                 (uint32_t)strtoul(idx->name[r[j].rid], NULL, 10), r[j].rs, r[j].re, r[j].mlen);
            if (fo) fwrite(buf, 1, nb, fo);
            ++n_rec; sum_qspan += (uint64_t)(r[j].qe - r[j].qs);
        }
        free(r);
    }
    if (fo) fclose(fo);
    if (sum_qspan_out) *sum_qspan_out = sum_qspan;
    ndo_index_free(idx);
    if (Q != T) ndo_reads_free(Q);
    ndo_reads_free(T);
    return n_rec;
}

#ifdef NDO_MAIN
/* usage: nd_oracle [-x preset] [-k K] [-w W] target.2bit query.2bit out.ovl */
int main(int argc, char **argv)
{
    ndo_opt o;
    int i;
    uint64_t sq;
    long n;
    ndo_opt_preset(&o, NULL);
    for (i = 1; i + 1 < argc && argv[i][0] == '-'; i += 2) {
        if (!strcmp(argv[i], "-x")) ndo_opt_preset(&o, argv[i + 1]);
        else if (!strcmp(argv[i], "-k")) o.k = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "-w")) o.w = atoi(argv[i + 1]);
        else { fprintf(stderr, "unknown option %s\n", argv[i]); return 1; }
    }
    if (argc - i < 3) { fprintf(stderr, "usage: nd_oracle [-x preset] [-k K] [-w W] target.2bit query.2bit out.ovl\n"); return 1; }
    n = ndo_step1_files(argv[i], argv[i + 1], &o, argv[i + 2], &sq);
    fprintf(stderr, "[nd_oracle] records=%ld sum_qspan=%llu\n", n, (unsigned long long)sq);
    return n < 0;
}
#endif
