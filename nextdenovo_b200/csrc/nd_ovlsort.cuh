// nd_ovlsort.cuh -- per-query stage of `ovl_sort`: the step that follows the overlap path in the pipeline
// (ref: nextDenovo:344-354 -> util/ovl_sort.c).  After the records of all input .ovl files were decoded, mirrored and sorted by
// (qname, match descending, query span ascending) (ref: util/ovl_sort.c:246-262,930-975), every query read is an independent
// sequential problem over its own records: a depth-capped admission test on a 64-base coverage histogram
// (`encode_ovl_filter`, :668-741, HiFi variant :627-665), then a per-read verdict -- chimera / low-coverage clipping / contained --
// that can retire admitted records again (`ovl_filter`, :446-565, with `check_chimer` :287-336, `check_chimer2` :339-382,
// `check_chimer_hq` :262-285, `del_repeat_alns` :393-443).  One GPU thread runs one read; everything here is __host__ __device__ so
// the same code is checked on a CPU-only box against the reference binary (tests/ovlsort_harness.cc, test infrastructure).
//
// The reference's arithmetic is kept to the type: 16-bit saturating-ish bins, the 8-bit `pcount` that wraps, unsigned/signed
// comparisons as C promotes them, float/double mixes in the depth test, and the in-place reuse of the bin array as the list of
// low-coverage runs.
#pragma once
#include <stdint.h>
#include "nd_core.cuh"

#define NDOS_BIN_OFFSET 6                  // ref: util/ovl_sort.h:3
#define NDOS_MAX_OVL_COV 150               // :4
#define NDOS_TOL_EDGE 50                   // :10
#define NDOS_TOL_COUNT 5                   // :11
#define NDOS_MIN_CONTAINED 2               // :12

struct ndos_rec { uint32_t qname, rev, qs, qe, tname, ts, te, match; };      // `overlap`, ref: lib/ovl.h:20-25; qe/te 0-based inclusive here

struct ndos_par { int max_bin_cov, max_flank_len, is_hq, min_seed_len, with_bl; };   // -k, -l, -H, shortest seed (ovl_sort.c:122), -l != 0

// order of the sort (ref: util/ovl_sort.c:246-262 cmp_ovl); ties keep their input order (glibc qsort is a merge sort here)
ND_HD bool ndos_less(const ndos_rec &a, const ndos_rec &b)
{
    if (a.qname != b.qname) return a.qname < b.qname;
    if (a.match != b.match) return a.match > b.match;
    return (int)(a.qe - a.qs) < (int)(b.qe - b.qs);
}

struct ndos_cov {                          // `cov_data`, ref: util/ovl_sort.h:57-66
    uint8_t pcount;
    uint32_t ovl_i;                        // uint16_t in the reference; the admission test keeps it below 64536
    uint32_t qs, qe, qlen, qmaxlen, qcov, binlen, bincount, contained, chimera, len;
};

#define NDOS_MAX(x, y) ((x) > (y) ? (x) : (y))
#define NDOS_MIN(x, y) ((x) > (y) ? (y) : (x))

// ---- admission of one record (`encode_ovl_filter`); first = the read's own pseudo record that opens the read (:945-953) -------------
ND_HD bool ndos_admit(const ndos_par &P, ndos_cov &c, uint16_t *bins, const ndos_rec &o, bool first)
{
    int lable = 1;
    if (first) {
        c.len = (c.qlen >> NDOS_BIN_OFFSET) + (P.is_hq ? 2u : 1u);
        for (uint32_t i = 0; i < c.len; ++i) bins[i] = 0;
        c.pcount = 1;
        c.qcov = c.bincount = c.binlen = c.contained = c.chimera = c.ovl_i = 0;
        c.qmaxlen = c.qlen * NDOS_MAX_OVL_COV * (P.is_hq ? 6u : 1u);
    } else if (c.qcov > c.qmaxlen || c.ovl_i > 65535u - 1000u) {
        lable = 0;
    } else if (P.is_hq) {                                                      // :647-655
        const int offset = 1 + (int)(c.qlen >> (NDOS_BIN_OFFSET + 1));
        const int s = (int)((o.qs + 10) >> (NDOS_BIN_OFFSET + 1)), e = (int)((o.qe - 10) >> (NDOS_BIN_OFFSET + 1));
        bins[s]++; bins[e + offset]++;
    } else {                                                                   // :691-727
        int m = 0, n = 0, mincov = 200;
        const int j = (int)((o.qs + 10) >> NDOS_BIN_OFFSET), k = (int)((o.qe - 10) >> NDOS_BIN_OFFSET);
        int d1 = (int)(o.qs - c.qs), d2 = (int)(o.qe - c.qe);
        if (d1 < 0) d1 = -d1; if (d2 < 0) d2 = -d2;
        if ((j > 15 || (uint32_t)k < c.len - 16u) && d1 < NDOS_TOL_EDGE && d2 < NDOS_TOL_EDGE)
            lable = c.pcount++ < NDOS_TOL_COUNT ? 2 : 0;
        if (lable) {
            for (int i = j + 1; i <= k; ++i) {
                if (!bins[i]) ++n;
                if (++bins[i] < mincov) mincov = bins[i];
                if (bins[i] > 65535 - 1000) bins[i]--;
                m += bins[i];
            }
            const float avg = (float)m / (k - j);
            const float dens = (float)c.bincount / c.binlen;
            const float capped = NDOS_MIN(NDOS_MAX(dens, 10), P.max_bin_cov);
            if ((mincov > P.max_bin_cov || avg > 1.3 * capped) && (o.qe - o.qs <= c.qlen * 0.8)) {
                for (int i = j + 1; i <= k; ++i) bins[i]--;
                lable = 0;
            } else {
                if (lable != 2) c.pcount = 1;
                lable = 1;
                c.binlen += (uint32_t)n;
                c.bincount += (uint32_t)(k - j);
            }
        }
    }
    if (lable) {
        c.qs = o.qs; c.qe = o.qe;
        c.qcov += c.qe - c.qs + 1;
        if (o.qname != o.tname && o.qs <= (uint32_t)P.max_flank_len && o.qe + (uint32_t)P.max_flank_len >= c.qlen) c.contained++;
        c.ovl_i++;
    }
    return lable != 0;
}

// ---- chimera tests on the coverage histogram ---------------------------------------------------------------------------------------
// acc[0..ovl_i): indices (into recs) of the admitted records in admission order, acc[0] = the pseudo record; dead[x] = 1 once the
// reference would have zeroed the record's qe
ND_HD int ndos_check_chimer_hq(const ndos_cov &c, const uint16_t *bins, const ndos_rec *recs, const uint32_t *acc, const uint8_t *dead)   // :262-285
{
    int l = 0, r;
    while (l < (int)c.len && bins[l] < 2) ++l;
    r = (int)c.len;
    while (r > 0 && bins[r - 1] < 2) --r;
    const int flank = 15;
    for (int i = l + 1; i < r - 1; ++i)
        if (bins[i] <= 1) {
            const int s = i > l + flank ? (i - flank) << NDOS_BIN_OFFSET : l << NDOS_BIN_OFFSET;
            const int e = i + flank < r ? (i + flank) << NDOS_BIN_OFFSET : r << NDOS_BIN_OFFSET;
            uint32_t j;
            for (j = 1; j < c.ovl_i; ++j) {
                const ndos_rec &o = recs[acc[j]];
                const uint32_t qe = dead[acc[j]] ? 0u : o.qe;
                if (o.qs < (uint32_t)s && qe > (uint32_t)e) break;
            }
            if (j >= c.ovl_i) return i;
        }
    return 0;
}

ND_HD int ndos_check_chimer(const ndos_par &P, const ndos_cov &c, const uint16_t *bins, const ndos_rec *recs, const uint32_t *acc, const uint8_t *dead)   // :287-310
{
    if (P.is_hq) return ndos_check_chimer_hq(c, bins, recs, acc, dead);
    int lable = 0, llable = 0, rlable = 0;
    for (int i = 1; (uint32_t)i < c.len - 1u; ++i) {
        if (bins[i] > 20 && ++llable) {
            if (lable && ++rlable >= 5) break;
        } else {
            const int l = NDOS_MAX(i - 5, 0);
            const int r = (int)NDOS_MIN((uint32_t)(i + 5), c.len - 1u);
            if (llable > 5 && (bins[l] > 20 || bins[r] > 20) && bins[i] <= NDOS_MAX(3, NDOS_MIN(bins[l], bins[r]) / 5)) lable = i;
        }
    }
    if (rlable < 5) lable = 0;
    return lable;
}

ND_HD int ndos_check_chimer2(const ndos_cov &c, uint16_t *bins, const ndos_rec *recs, const uint32_t *acc, const uint8_t *dead)   // hot break ends, :339-382
{
    const int sh = NDOS_BIN_OFFSET + 1;
    for (uint32_t i = 0; i < c.len / 2 + 1; ++i) bins[i] = 0;
    int s = (int)c.len, e = 0, cnt = 0, t;
    for (uint32_t i = 1; i < c.ovl_i; ++i) {
        if (dead[acc[i]]) continue;
        const ndos_rec &o = recs[acc[i]];
        if (!o.qe) continue;
        ++cnt;
        t = (int)((o.qs + 10) >> sh); if (t < s) s = t; bins[t]++;
        t = (int)((o.qe - 10) >> sh); if (t > e) e = t; bins[t]++;
    }
    t = 0;
    if (cnt > 20) {
        while (s < e && bins[s] < 4) ++s;
        while (e > s && bins[e] < 4) --e;
        int m = 0, ms = bins[s], me = bins[e];
        for (int i = s; i < e + 1; ++i) {
            if (i < s + 5 && bins[i] > ms) ms = bins[i];
            if (i > e - 5 && bins[i] > me) me = bins[i];
            if (bins[i] > bins[m]) m = i;
        }
        const float fra = 1;
        if (m > s + 5 && m < e - 5 && bins[m] > fra * NDOS_MAX(ms, me) && ((cnt > 75 && m > cnt / 5) || (cnt < 75 && m > cnt / 2))) t = m << sh;
    }
    return t;
}

ND_HD void ndos_del_repeat_alns(const ndos_par &P, const ndos_cov &c, uint16_t *bins, const ndos_rec *recs, const uint32_t *acc, uint8_t *dead)   // HiFi, :393-443
{
    const int offset = 1 + (int)(c.qlen >> (NDOS_BIN_OFFSET + 1));
    const int median = 5;
    const int flank_len = P.max_flank_len > 100 ? P.max_flank_len * 3 : 300;
    for (uint32_t i = 1; i < c.ovl_i; ++i) {
        const ndos_rec &o = recs[acc[i]];
        if (o.qs <= (uint32_t)flank_len && o.qe + (uint32_t)flank_len >= c.qlen) continue;
        const int s = (int)((o.qs + 10) >> (NDOS_BIN_OFFSET + 1)), e = (int)((o.qe - 10) >> (NDOS_BIN_OFFSET + 1));
        if (bins[s] >= median && bins[e + offset] >= median) dead[acc[i]] = 1;
    }
    for (uint32_t i = 0; i < c.len; ++i) bins[i] = 0;
    for (uint32_t i = 1; i < c.ovl_i; ++i) {
        if (dead[acc[i]]) continue;
        const ndos_rec &o = recs[acc[i]];
        if (!o.qe) continue;
        const int j = (int)((o.qs + 10) >> NDOS_BIN_OFFSET), k = (int)((o.qe - 10) >> NDOS_BIN_OFFSET);
        int mincov = 65535;
        for (int t = j + 1; t <= k; ++t) {
            if (++bins[t] > 65535 - 1000) bins[t]--;
            if (bins[t] < mincov) mincov = bins[t];
        }
        if (mincov > 2 * P.max_bin_cov) {
            for (int t = j + 1; t <= k; ++t) bins[t]--;
            dead[acc[i]] = 1;
        }
    }
}

// ---- the read's verdict (`ovl_filter`): retires records (dead[]), recounts `contained`, returns the .bl label: 'c', 'k' or 0 ------------------
ND_HD char ndos_finish(const ndos_par &P, ndos_cov &c, uint16_t *bins, const ndos_rec *recs, const uint32_t *acc, uint8_t *dead)
{
    if (P.is_hq) ndos_del_repeat_alns(P, c, bins, recs, acc, dead);
    if (P.with_bl) {
        int s = 0, e = 0;
        c.chimera = (uint32_t)ndos_check_chimer(P, c, bins, recs, acc, dead);
        if (c.chimera || !(c.contained || P.is_hq)) {
            int j = 0, k, m;
            // the runs of bins below the floor, with the chimera bin merged in, are written over the histogram itself, as the reference does
#define NDOS_PUSH_CHI() do { if (c.chimera && c.chimera < (uint32_t)s && ((!j) || c.chimera > (uint32_t)bins[j - 1])) { bins[j++] = (uint16_t)c.chimera; bins[j++] = (uint16_t)c.chimera; } } while (0)
            if (c.qcov > c.qlen * 10u) {
                const int floor_ = NDOS_MIN(4, P.max_bin_cov / 10);
                for (int i = 1; (uint32_t)i < c.len - 1u; ++i) {
                    if (bins[i] < floor_) { if (s == 0) s = i; e = i; }
                    else if (s) {
                        NDOS_PUSH_CHI();
                        bins[j++] = (uint16_t)s; bins[j++] = (uint16_t)e;
                        s = e = 0;
                    }
                }
                if (s) { NDOS_PUSH_CHI(); bins[j++] = (uint16_t)s; bins[j++] = (uint16_t)e; }
                if (c.chimera && (j == 0 || c.chimera > (uint32_t)bins[j - 1])) { bins[j++] = (uint16_t)c.chimera; bins[j++] = (uint16_t)c.chimera; }
            } else if (c.chimera) { bins[j++] = (uint16_t)c.chimera; bins[j++] = (uint16_t)c.chimera; }
#undef NDOS_PUSH_CHI
            if (j) {
                m = j;
                if (bins[0] < 5) m -= 2;
                if ((uint32_t)bins[j - 1] > c.len - 5u) m -= 2;
                if (m > 0) {
                    int i;
                    k = 0; m = bins[0];
                    for (i = 2; i < j; i += 2)
                        if (bins[i] - bins[i - 1] > m) { m = bins[i] - bins[i - 1]; k = i; }
                    if (c.len - (uint32_t)bins[i - 1] > (uint32_t)m) { m = (int)(c.len - (uint32_t)bins[i - 1]); s = bins[i - 1]; e = (int)c.len; }
                    else if ((uint32_t)bins[k + 1] > c.len - 5u) { s = k > 0 ? bins[k - 1] : 0; e = (int)c.len; }     // k == 0 reads before the array in the reference
                    else if (k == 0 || bins[k - 2] < 5) { s = 0; e = bins[k]; }
                    else { s = bins[k - 1]; e = bins[k]; }
                    const int flank = 5;
                    s = s > flank ? (s - flank) << NDOS_BIN_OFFSET : 0;
                    e = (e + flank) << NDOS_BIN_OFFSET;
                    if (m > (P.min_seed_len >> NDOS_BIN_OFFSET) * 2 / 3) {
                        c.chimera = 0;
                        for (uint32_t x = 1; x < c.ovl_i; ++x) {
                            const ndos_rec &o = recs[acc[x]];
                            const uint32_t qe = dead[acc[x]] ? 0u : o.qe;
                            if (o.qs < (uint32_t)s || qe > (uint32_t)e) dead[acc[x]] = 1;
                        }
                    } else c.chimera = 1;
                } else s = e = 0;
            }
        }
        if (!P.is_hq && c.qcov > c.qlen * 20u && !c.chimera && c.contained < NDOS_MIN_CONTAINED) {
            c.chimera = (uint32_t)ndos_check_chimer2(c, bins, recs, acc, dead);
            if (!e) e = (int)c.qlen;
            if (c.chimera <= (uint32_t)(s + (15 << NDOS_BIN_OFFSET)) || c.chimera + (uint32_t)(15 << NDOS_BIN_OFFSET) >= (uint32_t)e) c.chimera = 0;
        }
    }
    c.contained = 0;
    for (uint32_t x = 0; x < c.ovl_i; ++x) {
        if (dead[acc[x]]) continue;
        const ndos_rec &o = recs[acc[x]];
        if (!o.qe) continue;
        if (o.qname != o.tname && o.qs <= (uint32_t)P.max_flank_len && o.qe + (uint32_t)P.max_flank_len >= c.qlen) {
            if (!P.is_hq) c.contained++;
            else if (o.match >= (o.qe - o.qs + 1) * 0.9) c.contained++;
        }
    }
    if (!P.with_bl) return 0;
    return c.contained >= NDOS_MIN_CONTAINED ? 'c' : c.chimera ? 'k' : 0;
}

// one read, start to end.  recs[0..n): the read's records in sorted order, recs[0] the pseudo record (qname == tname, [0, qlen-1]);
// acc[0..n) and dead[0..n) are per-record scratch/output (dead[x] stays 1 for records that were never admitted); bins: >= qlen/64 + 2
// entries.  Returns the label; *n_out = records that go to the output.
ND_HD char ndos_one_read(const ndos_par &P, uint32_t qlen, const ndos_rec *recs, uint32_t n, uint16_t *bins, uint32_t *acc, uint8_t *dead, uint32_t *n_out)
{
    ndos_cov c; c.qlen = qlen; c.pcount = 1; c.ovl_i = 0; c.qs = c.qe = 0; c.len = 0;
    c.qmaxlen = c.qcov = c.binlen = c.bincount = c.contained = c.chimera = 0;
    for (uint32_t x = 0; x < n; ++x) {
        const uint32_t at = c.ovl_i;
        if (ndos_admit(P, c, bins, recs[x], x == 0)) { acc[at] = x; dead[x] = 0; } else dead[x] = 1;
    }
    // dead[] is indexed by record, acc[] by admission order
    const char lab = ndos_finish(P, c, bins, recs, acc, dead);
    uint32_t k = 0;
    for (uint32_t x = 0; x < n; ++x) if (!dead[x] && recs[x].qe) ++k;
    *n_out = k;
    return lab;
}
