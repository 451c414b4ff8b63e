// nd_step2.hpp -- host side of `--step 2` (cns_align): which overlaps of a query get re-mapped against a
// throw-away mini index, how the re-mapping results replace the first-pass coordinates, and the 10-field
// writer with its order-dependent contained/dovetail bookkeeping and the `.bl` side file.  Header-only,
// no CUDA: the mapping itself (first pass and every re-mapping) goes through a backend -- the GPU pipeline
// in the product; tests may plug the CPU checker in.
//
// What it replaces in the reference (all "ref:" relative to /root/reference/):
//   worker_for(), step-2 branch                         minimap2/map.c:988-1128
//   check_realign_nextdenovo, update_reg_nextdenovo     minimap2/map.c:804-877
//   pipeline step 2 writer, step-2 branch               minimap2/map.c:1305-1328
//   filter_ovl, fill_aln*, merge_aln, out_bl            lib/ovl.c:254-361,449-563
//   encode_ovl_i, init_ovl_mode                         lib/ovl.c:70-75,205-253
//   khash (bucket order decides the line order of .bl)  util/khash.h:244-352
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include <string>
#include <algorithm>
#include <unordered_map>

namespace nds2 {

enum { MAX_CON = 2, INIT_ALNM = 5, EDGEBACKLEN = 50 };           // ref: lib/ovl.h:69-71

struct Opt {                                       // ref: minimap2/main.c:194-198, minimap2/options.c:50-62
    int minlen = 2000, maxhan1 = 5000, maxhan2 = 500, minmatch = 100, mode = 2, cn = 20, kn = 17, wn = 10, outctn = 0;
    float minide = 0.05f;
};

struct Reg {                                       // mm_reg1_t fields this path reads or rewrites
    int32_t rid = 0, rev = 0, qs = 0, qe = 0, rs = 0, re = 0, mlen = 0, blen = 0, score = 0, cnt = 0;
};

inline int check_realign(int rev, uint32_t qs, uint32_t qe, uint32_t qlen, uint32_t ts, uint32_t te, uint32_t tlen, int32_t maxhan1, int32_t maxhan2)
{                                                   // ref: minimap2/map.c:804-821 (signed/unsigned comparisons as written there)
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

// ---- open-addressing table with klib's bucket layout (insert-only use) ---------------------------------
// out_bl() walks the buckets, so the order of the lines of the .bl file is the bucket order: same hash (identity),
// same triangular probing, same growth schedule and the same in-place "kick-out" rehash.
struct KTable {
    std::vector<uint32_t> keys; std::vector<int32_t> val; std::vector<uint8_t> used;
    uint32_t n_buckets = 0, size = 0, upper = 0;
    void resize(uint32_t want)
    {
        uint32_t nb = want; --nb; nb |= nb >> 1; nb |= nb >> 2; nb |= nb >> 4; nb |= nb >> 8; nb |= nb >> 16; ++nb;
        if (nb < 4) nb = 4;
        if (size >= (uint32_t)(nb * 0.77 + 0.5)) return;
        std::vector<uint8_t> nused(nb, 0);
        if (n_buckets < nb) { keys.resize(nb); val.resize(nb); }
        const uint32_t mask = nb - 1;
        for (uint32_t j = 0; j != n_buckets; ++j) {
            if (!used[j]) continue;
            uint32_t key = keys[j]; int32_t v = val[j];
            used[j] = 0;                            // "deleted" in the old table
            for (;;) {
                uint32_t i = key & mask, step = 0;
                while (nused[i]) i = (i + (++step)) & mask;
                nused[i] = 1;
                if (i < n_buckets && used[i]) { std::swap(keys[i], key); std::swap(val[i], v); used[i] = 0; }
                else { keys[i] = key; val[i] = v; break; }
            }
        }
        if (n_buckets > nb) { keys.resize(nb); val.resize(nb); }
        used.swap(nused); n_buckets = nb; upper = (uint32_t)(nb * 0.77 + 0.5);
    }
    int32_t *get(uint32_t key)
    {
        if (!n_buckets) return 0;
        const uint32_t mask = n_buckets - 1; uint32_t i = key & mask, last = i, step = 0;
        while (used[i] && keys[i] != key) { i = (i + (++step)) & mask; if (i == last) return 0; }
        return used[i] ? &val[i] : 0;
    }
    int32_t *put(uint32_t key, bool *absent)
    {
        if (size >= upper) resize(n_buckets > (size << 1) ? n_buckets - 1 : n_buckets + 1);
        const uint32_t mask = n_buckets - 1; uint32_t i = key & mask, step = 0;
        while (used[i] && keys[i] != key) i = (i + (++step)) & mask;
        *absent = !used[i];
        if (!used[i]) { used[i] = 1; keys[i] = key; val[i] = -1; ++size; }
        return &val[i];
    }
};

// ---- per-read bookkeeping of the writer (ovlinfo_aln, ref: lib/ovl.h:74-89) ------------------------------
struct Aln { uint32_t s = 0, e = 0; };
struct ReadInfo {
    uint16_t lc = 0, rc = 0, alni = 0, alnm = 0;
    uint32_t con = 0, lim = 0, rim = 0, llm = 0, rlm = 0, len = 0;
    Aln alnl; std::vector<Aln> alns;
    void init_aln() { alnm = INIT_ALNM; alns.assign(alnm, Aln()); }
    void merge_aln()
    {
        for (int i = 1; i < alnm; ++i) { Aln t = alns[i]; int j; for (j = i; j > 0 && alns[j - 1].s > t.s; --j) alns[j] = alns[j - 1]; alns[j] = t; }
        int i = 0;
        while (i < alnm - 1) {
            if (!alns[i].e) { ++i; continue; }
            int j;
            for (j = i + 1; j < alnm; ++j) {
                if (alns[j].e <= alns[i].e) alns[j].e = 0;
                else if (alns[j].s <= alns[i].e && alns[j].e >= alns[i].e) { alns[i].e = alns[j].e; alns[j].e = 0; }
                else break;
            }
            i = j;
        }
    }
    uint16_t find_alni()
    {
        uint16_t i;
        if (alni != alnm - 1) for (i = 0; i < alnm; ++i) if (!alns[i].e) return i;
        merge_aln();
        for (i = 0; i < alnm; ++i) if (!alns[i].e) return i;
        alnm = (uint16_t)(alnm + INIT_ALNM); alns.resize(alnm);
        return i;
    }
    void fill_aln(uint32_t as, uint32_t ae) { if (con < MAX_CON) { alni = find_alni(); alns[alni].s = as + EDGEBACKLEN; alns[alni].e = ae - EDGEBACKLEN; } }
    void fill_alnl(uint32_t as, uint32_t ae) { if (con < MAX_CON && ae - as > alnl.e - alnl.s) { alnl.s = as; alnl.e = ae; } }
};

struct Ovl { uint8_t rev; uint32_t qname, qs, qe, qlen, tname, ts, te, tlen, identity; };   // overlap_i, ref: lib/ovl.h:27-32

struct Writer {
    KTable tab; std::vector<ReadInfo> info;
    uint32_t prev_q = 0, prev_t = 0;
    std::vector<uint8_t> out;                       // 10-field records (after the 2-byte header written by begin())
    void begin() { out.push_back(0); out.push_back(255); }       // init_ovl_mode(fp, 10), ref: lib/ovl.c:70-75

    ReadInfo &slot(uint32_t name, uint32_t len, bool *fresh)
    {
        bool absent; int32_t *v = tab.put(name, &absent);
        if (absent) { *v = (int32_t)info.size(); info.emplace_back(); info.back().len = len; info.back().init_aln(); }
        *fresh = absent;
        return info[(size_t)*v];
    }

    int filter(const Ovl &o, int32_t maxhan1, int32_t maxhan2)                   // filter_ovl, ref: lib/ovl.c:449-563
    {
        const uint32_t m1 = (uint32_t)maxhan1, m2 = (uint32_t)maxhan2;
        bool fresh;
        {
            ReadInfo &l = slot(o.qname, o.qlen, &fresh);
            if (!fresh) { if (l.con < MAX_CON) { if (o.qs <= m2 && l.lc < UINT16_MAX) l.lc++; if (o.qlen - o.qe <= m2 && l.rc < UINT16_MAX) l.rc++; } }
            else { if (o.qs <= m2) l.lc++; if (o.qlen - o.qe <= m2) l.rc++; }
        }
        {
            ReadInfo &r = slot(o.tname, o.tlen, &fresh);
            if (!fresh) { if (r.con < MAX_CON) { if (o.ts <= m2 && r.rc < UINT16_MAX) r.lc++; if (o.tlen - o.te <= m2 && r.rc < UINT16_MAX) r.rc++; } }
            else { if (o.ts <= m2) r.lc++; if (o.tlen - o.te <= m2) r.rc++; }
        }
        ReadInfo &L = info[(size_t)*tab.get(o.qname)], &R = info[(size_t)*tab.get(o.tname)];
        L.fill_aln(o.qs, o.qe);
        R.fill_aln(o.ts, o.te);
        if (L.con < MAX_CON && o.qs <= m2 && o.qe + m2 >= o.qlen) { if (++L.con >= MAX_CON) { L.alns.clear(); L.alnm = 0; } return 0; }
        if (R.con < MAX_CON && o.ts <= m2 && o.te + m2 >= o.tlen) { if (++R.con >= MAX_CON) { R.alns.clear(); R.alnm = 0; } return 0; }
        if (L.con >= MAX_CON || R.con >= MAX_CON) return 0;
        const uint32_t alnlen = std::max(o.qe - o.qs, o.te - o.ts);
        if (o.rev) {
            if (o.qs <= m1 && o.ts <= m1) {
                if (o.qs <= m2 && o.ts <= m2) { if (alnlen > L.llm) L.llm = alnlen; if (alnlen > R.llm) R.llm = alnlen; if (o.identity > L.lim) L.lim = o.identity & 0x7fff; if (o.identity > R.lim) R.lim = o.identity & 0x7fff; }
                return 1;
            } else if (o.qlen - o.qe <= m1 && o.tlen - o.te <= m1) {
                if (o.qlen - o.qe <= m2 && o.tlen - o.te <= m2) { if (alnlen > L.rlm) L.rlm = alnlen; if (alnlen > R.rlm) R.rlm = alnlen; if (o.identity > L.rim) L.rim = o.identity & 0x7fff; if (o.identity > R.rim) R.rim = o.identity & 0x7fff; }
                return 1;
            }
        } else {
            if (o.qlen - o.qe <= m1 && o.ts <= m1) {
                if (o.qlen - o.qe <= m2 && o.ts <= m2) { if (alnlen > L.rlm) L.rlm = alnlen; if (alnlen > R.llm) R.llm = alnlen; if (o.identity > L.rim) L.rim = o.identity & 0x7fff; if (o.identity > R.lim) R.lim = o.identity & 0x7fff; }
                return 1;
            } else if (o.qs <= m1 && o.tlen - o.te <= m1) {
                if (o.qs <= m2 && o.tlen - o.te <= m2) { if (alnlen > L.llm) L.llm = alnlen; if (alnlen > R.rlm) R.rlm = alnlen; if (o.identity > L.lim) L.lim = o.identity & 0x7fff; if (o.identity > R.rim) R.rim = o.identity & 0x7fff; }
                return 1;
            }
        }
        if (o.qs <= m1 && o.qe + m1 >= o.qlen) return 1;
        if (o.ts <= m1 && o.te + m1 >= o.tlen) return 1;
        L.fill_alnl(o.qs, o.qe);
        R.fill_alnl(o.ts, o.te);
        return 0;
    }

    static inline int put_varint(uint8_t *dst, uint32_t v)        // same byte strings as init_encode_table / encode_ovl_i (lib/ovl.c:10-29,236-247)
    {
        if (v <= 127) { dst[0] = (uint8_t)v; return 1; }
        int m = 0;
        for (int sh = 28; sh >= 0; sh -= 7) { uint32_t g = v >> sh & 127; if (g > 0 || m > 0) dst[m++] = (uint8_t)(g | 128); }
        dst[m - 1] &= 127;
        return m;
    }
    void encode(Ovl o)                              // encode_ovl_i, ref: lib/ovl.c:205-253
    {
        uint32_t f[10]; const uint32_t tl = o.te - o.ts;
        f[3] = o.qe - o.qs;
        if (o.qname >= prev_q) f[0] = o.qname - prev_q; else { o.rev |= 2; f[0] = prev_q - o.qname; }
        if (o.tname >= prev_t) f[4] = o.tname - prev_t; else { o.rev |= 4; f[4] = prev_t - o.tname; }
        f[7] = o.qname == prev_q ? 0 : o.qlen;
        f[8] = o.tname == prev_t ? 0 : o.tlen;
        if (f[3] >= tl) f[6] = f[3] - tl; else { o.rev |= 8; f[6] = tl - f[3]; }
        prev_q = o.qname; prev_t = o.tname;
        f[1] = o.rev; f[2] = o.qs; f[5] = o.ts; f[9] = o.identity;
        uint8_t b[64]; int n = 0;
        for (int i = 0; i < 10; ++i) n += put_varint(b + n, f[i]);
        out.insert(out.end(), b, b + n);
    }

    // one query's hits, already post-processed, in order (ref: minimap2/map.c:1290-1328)
    void emit_query(const Opt &o, const std::vector<Reg> &regs, uint32_t qname, uint64_t qkey, uint32_t qlen,
                    const uint32_t *t_id, const uint64_t *t_key, const uint32_t *t_len, uint64_t *n_rec, uint64_t *sum_qspan)
    {
        for (const Reg &r : regs) {
            if (qkey == t_key[r.rid]) continue;
            const uint32_t tlen = t_len[r.rid];
            if ((r.qe - r.qs >= o.minlen || r.mlen == r.blen) && (r.mlen == 3 || ((float)r.mlen >= (float)r.blen * o.minide && r.mlen >= o.minmatch)) &&
                r.blen >= (int32_t)qlen / 50 && (uint32_t)r.blen >= tlen / 50) {
                Ovl v = { (uint8_t)r.rev, qname, (uint32_t)r.qs, (uint32_t)r.qe, qlen, t_id[r.rid], (uint32_t)r.rs, (uint32_t)r.re, tlen,
                          (uint32_t)((uint64_t)(int64_t)r.mlen * 10000 / (uint64_t)(int64_t)r.blen) };
                const int keep = filter(v, o.maxhan1, o.maxhan2);
                bool wrote = false;
                if (keep) { encode(v); wrote = true; }
                else if (o.outctn) {
                    if (v.qs <= (uint32_t)o.maxhan2 && v.qe + o.maxhan2 >= v.qlen) { info[(size_t)*tab.get(v.qname)].con = 0; encode(v); wrote = true; }
                    else if (v.ts <= (uint32_t)o.maxhan2 && v.te + o.maxhan2 >= v.tlen) { info[(size_t)*tab.get(v.tname)].con = 0; encode(v); wrote = true; }
                }
                if (wrote) { if (n_rec) ++*n_rec; if (sum_qspan) *sum_qspan += (uint64_t)(r.qe - r.qs); }
            }
        }
    }

    std::string bl_text()                           // out_bl, ref: lib/ovl.c:339-361
    {
        std::string s; char b[256];
        for (uint32_t k = 0; k < tab.n_buckets; ++k) {
            if (!tab.used[k]) continue;
            ReadInfo &v = info[(size_t)tab.val[k]];
            if (v.con < MAX_CON) {
                snprintf(b, sizeof b, "%u\t%u\t%hu\t%hu\t%u\t%u\t%u\t%u\t%u\t%u\t%u", tab.keys[k], v.con, v.lc, v.rc, v.lim, v.rim, v.llm, v.rlm, v.len, v.alnl.s, v.alnl.e);
                s += b;
                v.merge_aln();
                for (int i = 0; i < v.alnm; ++i) if (v.alns[i].e) { snprintf(b, sizeof b, "\t%u\t%u", v.alns[i].s - EDGEBACKLEN, v.alns[i].e + EDGEBACKLEN); s += b; }
                s += "\n";
            } else { snprintf(b, sizeof b, "%u\t%u\n", tab.keys[k], v.con); s += b; }
        }
        return s;
    }
};

// ---- per-query post-processing ----------------------------------------------------------------------------

// A re-mapping request: map `mapped` against the mini index made of `group`.  In the common branch the group is the query
// itself and the mapped sequences are its candidate targets (ref: map.c:1046-1084); with >= 200 candidates the group is a
// batch of up to cn targets and the query is mapped (ref: map.c:1085-1125).
// Ambiguous bases: a target mapped against the query's mini index has them turned into A (the reference passes "ACGT"[code]
// characters, code 4 is the terminator byte, and its base table maps byte 0 to A); index reads and a mapped query keep them.
struct Remap {
    int query = 0;                                  // batch-local query index this request belongs to
    bool query_is_index = true;
    std::vector<int32_t> group_rids;                // target rids in the mini index (empty when the query is the index)
    int32_t mapped_rid = -1;                        // target rid of the mapped sequence (or -1: the query is mapped)
    int reg_k = -1;                                 // branch 1: the first-pass hit this re-mapping may replace
    int s = 0, e = 0;                               // branch 2: range of first-pass hits covered by this group
    std::vector<Reg> result;                        // filled by the backend: hits in the reference's order; rid = index into the group
};

struct QueryJob {
    uint32_t qlen = 0; uint64_t qkey = 0;
    std::vector<Reg> regs;                          // first-pass hits in mm_gen_regs order; rewritten in place
    int c = 0, seq_index = 0; bool skip = false, branch1 = true;
};

// first loop of the step-2 branch: decide which hits are dovetail candidates (mlen = 2), contained (3), duplicates (1), self (0)
inline void mark_candidates(const Opt &o, QueryJob &Q, const uint64_t *t_key, const uint32_t *t_len)
{
    std::unordered_map<int32_t, int> first;
    std::vector<Reg> &R = Q.regs;
    int c = 0, seq_index = 0;
    for (int k = 0; k < (int)R.size(); ++k) {
        Reg &r = R[k];
        if (Q.qkey == t_key[r.rid]) { r.mlen = 0; continue; }
        const uint32_t tp = (uint32_t)r.mlen;
        auto it = first.find(r.rid);
        if (it == first.end()) it = first.emplace(r.rid, k).first; else r.mlen = 1;
        const int l = it->second;
        const uint32_t alnlen = t_len[r.rid] > Q.qlen ? t_len[r.rid] : Q.qlen;
        if (l != k && (R[l].mlen == 2 || (double)r.blen < R[l].blen * 0.8 || (uint32_t)r.blen < alnlen / 3)) continue;
        else if (r.qe - r.qs >= o.minlen && (float)tp >= (float)r.blen * o.minide && tp >= (uint32_t)o.minmatch) {
            if (check_realign(r.rev, r.qs, r.qe, Q.qlen, r.rs, r.re, t_len[r.rid], o.maxhan1, 0)) {
                if (R[l].mlen == 3) c--;
                R[l].mlen = o.mode ? 2 : (int32_t)tp;
                seq_index++;
            } else if ((uint32_t)r.qs <= (uint32_t)o.maxhan2 && (uint32_t)(r.qe + o.maxhan2) >= Q.qlen) {
                R[l].mlen = 3;
                if (++c >= MAX_CON) break;
            }
        }
    }
    Q.c = c; Q.seq_index = seq_index;
    Q.skip = c >= MAX_CON || !o.mode;
    Q.branch1 = (o.mode == 2 && seq_index < 200) || seq_index < 20;
}

// all re-mapping requests of a query (issued speculatively: the reference stops early once MAX_CON containments are seen)
inline void plan_remaps(const Opt &o, int qi, QueryJob &Q, std::vector<Remap> &out)
{
    if (Q.skip) return;
    std::vector<Reg> &R = Q.regs;
    if (Q.branch1) {
        for (int k = 0; k < (int)R.size(); ++k) {
            if (R[k].mlen != 2) continue;
            Remap m; m.query = qi; m.query_is_index = true; m.mapped_rid = R[k].rid; m.reg_k = k;
            out.push_back(m);
        }
    } else {
        const int cn = (int)((float)Q.seq_index / ((Q.seq_index + o.cn - 1) / o.cn) + 0.999);
        Remap m; m.query = qi; m.query_is_index = false; m.s = 0;
        int seq_index = 0;
        for (int k = 0; k < (int)R.size(); ++k) {
            if (R[k].mlen != 2) continue;
            R[k].blen = seq_index;                  // the hit remembers its slot in the mini index (ref: map.c:1091)
            m.group_rids.push_back(R[k].rid);
            if (++seq_index >= cn) { m.e = k + 1; out.push_back(m); m.group_rids.clear(); m.s = k + 1; seq_index = 0; }
        }
        if (seq_index) { m.e = (int)R.size(); out.push_back(m); }
    }
}

// update_reg_nextdenovo, ref: minimap2/map.c:823-877
inline int update_regs(const std::vector<Reg> &nw, std::vector<Reg> &R, int s, int e, uint32_t t_l, const uint32_t *group_len, int32_t maxhan1, int32_t maxhan2)
{
    const int n = (int)nw.size();
    int i = 0, c = 0;
    for (; s < e; ++s) {
        Reg &r = R[s];
        if (r.mlen != 2) continue;
        int l = -1, pi = i, t = 0; uint32_t alnlen = 0;
        for (; i < n && t < 10; ++i) {
            const Reg &x = nw[i];
            if (x.rid == r.blen) {
                if (l == -1) { l = i; alnlen = (uint32_t)(nw[i].blen * 0.8); }
                if ((uint32_t)x.blen >= alnlen && check_realign(x.rev, x.qs, x.qe, t_l, x.rs, x.re, group_len[x.rid], maxhan1, maxhan2)) { l = i; break; }
                ++t;
                if ((uint32_t)x.qs <= (uint32_t)maxhan2 && (uint32_t)(x.qe + maxhan2) >= t_l) { l = i; ++c; break; }
            } else if (l >= 0) { --i; break; }
        }
        if (l >= 0) { const Reg &x = nw[l]; r.rev = x.rev; r.qs = x.qs; r.qe = x.qe; r.rs = x.rs; r.re = x.re; r.mlen = x.mlen; r.blen = x.blen; }
        else i = pi;
    }
    return c;
}

// apply the re-mapping results of one query in the reference's sequential order (with its early exits)
inline void apply_remaps(const Opt &o, QueryJob &Q, std::vector<Remap> &reqs, size_t r0, size_t r1, const uint32_t *t_len)
{
    if (Q.skip) return;
    std::vector<Reg> &R = Q.regs;
    int c = Q.c;
    if (Q.branch1) {
        for (size_t x = r0; x < r1; ++x) {
            Remap &m = reqs[x];
            Reg &r = R[m.reg_k];
            if (m.result.empty()) continue;
            const std::vector<Reg> &nw = m.result; const int n = (int)nw.size();
            const uint32_t alnlen = (uint32_t)(nw[0].blen * 0.8), mlen_seq = t_len[r.rid];
            int l;
            for (l = 0; l < n && l < 10; ++l)
                if ((uint32_t)nw[l].blen >= alnlen && check_realign(nw[l].rev, nw[l].qs, nw[l].qe, mlen_seq, nw[l].rs, nw[l].re, Q.qlen, o.maxhan1, o.maxhan2)) break;
            if (l == 10 || l == n) l = 0;
            Reg z = nw[l];
            std::swap(z.qs, z.rs); std::swap(z.qe, z.re);
            z.rid = r.rid;
            r = z;
            if ((uint32_t)r.qs <= (uint32_t)o.maxhan2 && (uint32_t)(r.qe + o.maxhan2) >= Q.qlen) { if (++c >= MAX_CON) break; }
        }
    } else {
        for (size_t x = r0; x < r1; ++x) {
            Remap &m = reqs[x];
            std::vector<Reg> nw = m.result;
            std::stable_sort(nw.begin(), nw.end(), [](const Reg &a, const Reg &b) { return a.rid < b.rid; });    // qsort on rid (glibc: merge sort)
            std::vector<uint32_t> glen; for (int32_t rid : m.group_rids) glen.push_back(t_len[rid]);
            c += update_regs(nw, R, m.s, m.e, Q.qlen, glen.data(), o.maxhan1, o.maxhan2);
            const bool last = x + 1 == r1;
            if (!last && c >= MAX_CON) break;
        }
    }
}

} // namespace nds2
