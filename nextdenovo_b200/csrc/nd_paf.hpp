// nd_paf.hpp -- PAF text writer of the drop-in executable (host C++): mm_write_paf3 + write_tags (ref: minimap2/format.c:282-318),
// mm_event_identity (:264-275), and the two per-hit numbers only PAF shows: `dv` from mm_est_err (ref: minimap2/esterr.c:30-64; the
// integer counts come from the GPU, the float arithmetic is done here as the reference does it on the host) and `mapq` from
// mm_set_mapq / mm_set_inv_mapq (ref: minimap2/hit.c:420-491).  In all-chain mode (`-x ava-*`, -X, -P: the only mode this build
// maps in) no hit has a parent, so every chain-level hit is type 'S' with mapq 0; only the second half of a z-drop split and an
// inversion hit carry id == parent == -1 (ref: hit.c:94, align.c:824-825) and get the primary branch of mm_set_mapq.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <string>
#include <vector>
#include <algorithm>
#include "ndgpu.h"

namespace ndpaf {

inline void put_int(std::string &s, int c)                 // mm_sprintf_lite "%d" (format.c:37-45)
{
    char buf[16]; int l = 0; unsigned x = c >= 0 ? (unsigned)c : (unsigned)-(long long)c;
    do { buf[l++] = (char)(x % 10 + '0'); x /= 10; } while (x > 0);
    if (c < 0) buf[l++] = '-';
    while (l > 0) s.push_back(buf[--l]);
}
inline void put_uint(std::string &s, uint32_t x)
{
    char buf[16]; int l = 0;
    do { buf[l++] = (char)(x % 10 + '0'); x /= 10; } while (x > 0);
    while (l > 0) s.push_back(buf[--l]);
}

struct Par { int min_chain_score, match_sc, is_sr, out_cg; };

// mapq of all hits of one query (ref: hit.c:446-491); h/e: the query's hits in output order
inline void set_mapq(const Par &P, const ndg_hit_t *h, const ndg_hit_extra_t *e, int n, int rep_len, std::vector<int> &mapq)
{
    static const float q_coef = 40.0f;
    mapq.assign((size_t)n, 0);
    if (n == 0) return;
    int64_t sum_sc = 0;
    for (int i = 0; i < n; ++i) if (e[i].flags & 2u) sum_sc += e[i].chain_score;
    const float uniq_ratio = (float)sum_sc / (sum_sc + rep_len);
    for (int i = 0; i < n; ++i) {
        const bool inv = e[i].flags & 1u, pri = e[i].flags & 2u, has_p = e[i].flags & 16u;
        if (inv || !pri) { mapq[i] = 0; continue; }
        const int score = e[i].chain_score, cnt = h[i].cnt;
        int mq;
        float pen_s1 = (score > 100 ? 1.0f : 0.01f * score) * uniq_ratio;
        float pen_cm = cnt > 10 ? 1.0f : 0.1f * cnt;
        pen_cm = pen_s1 < pen_cm ? pen_s1 : pen_cm;
        const int subsc = 0 > P.min_chain_score ? 0 : P.min_chain_score;       // mm_reg1_t::subsc stays 0 without mm_set_parent
        {                                                                       // dp_max2 stays 0 likewise: the second formula (:475-483)
            const float x = (float)subsc / e[i].score0;
            if (has_p) {
                const float identity = (float)h[i].mlen / h[i].blen;
                mq = (int)(identity * pen_cm * q_coef * (1.0f - x) * logf((float)e[i].dp_max / P.match_sc));
            } else mq = (int)(pen_cm * q_coef * (1.0f - x) * logf((float)score));
        }
        mq -= (int)(4.343f * logf(0 + 1) + .499f);
        mq = mq > 0 ? mq : 0;
        mq = mq < 60 ? mq : 60;
        if (has_p && e[i].dp_max > 0 && mq == 0) mq = 1;
        mapq[i] = mq;
    }
    // mm_set_inv_mapq (hit.c:420-444): an inversion takes the smaller mapq of its neighbours in (rid, rs) order
    if (n < 3) return;
    bool any_inv = false;
    for (int i = 0; i < n; ++i) if (e[i].flags & 1u) { any_inv = true; break; }
    if (!any_inv) return;
    std::vector<std::pair<uint64_t, int>> aux;                                  // every hit qualifies: parent < 0 in all-chain mode
    for (int i = 0; i < n; ++i) aux.push_back({ (uint64_t)(uint32_t)h[i].rid << 32 | (uint32_t)h[i].rs, i });
    std::stable_sort(aux.begin(), aux.end(), [](const std::pair<uint64_t, int> &a, const std::pair<uint64_t, int> &b) { return a.first < b.first; });
    for (size_t i = 1; i + 1 < aux.size(); ++i) {
        const int k = aux[i].second;
        if (e[k].flags & 1u) { const int l = mapq[aux[i - 1].second], r = mapq[aux[i + 1].second]; mapq[k] = l < r ? l : r; }
    }
}

// div of mm_est_err from the GPU's counts (esterr.c:36-41,56-61); -1 = unset
inline float est_div(const ndg_hit_t &h, const ndg_hit_extra_t &e, const ndg_query_extra_t &q, int qlen, int l_ref)
{
    if (q.n_mini == 0 || h.cnt == 0 || e.n_span < 0) return -1.0f;
    const float avg_k = (float)q.sum_span / q.n_mini;
    int n_tot = e.n_span;
    if (h.qs > avg_k && h.rs > avg_k) ++n_tot;
    if (qlen - h.qs > avg_k && l_ref - h.re > avg_k) ++n_tot;
    return logf((float)n_tot / e.n_match) / avg_k;
}

// one PAF line (without the trailing newline mm_err_puts adds)
inline void write_hit(std::string &s, const Par &P, uint32_t qname, int qlen, uint32_t tname, int tlen, const ndg_hit_t &h, const ndg_hit_extra_t &e,
                      const ndg_query_extra_t &q, int mapq, const uint32_t *cig_pool)
{
    const bool inv = e.flags & 1u, pri = e.flags & 2u, has_p = e.flags & 16u;
    put_uint(s, qname); s.push_back('\t'); put_int(s, qlen); s.push_back('\t'); put_int(s, h.qs); s.push_back('\t'); put_int(s, h.qe); s.push_back('\t');
    s.push_back("+-"[h.rev ? 1 : 0]); s.push_back('\t');
    put_uint(s, tname); s.push_back('\t'); put_int(s, tlen); s.push_back('\t'); put_int(s, h.rs); s.push_back('\t'); put_int(s, h.re);
    s.push_back('\t'); put_int(s, h.mlen); s.push_back('\t'); put_int(s, h.blen); s.push_back('\t'); put_int(s, mapq);
    const uint32_t *cig = cig_pool ? cig_pool + e.cigar_off : nullptr;
    if (has_p) {
        s += "\tNM:i:"; put_int(s, h.blen - h.mlen + e.n_ambi); s += "\tms:i:"; put_int(s, e.dp_max); s += "\tAS:i:"; put_int(s, e.dp_score); s += "\tnn:i:"; put_int(s, e.n_ambi);
    }
    s += "\ttp:A:"; s.push_back(pri ? (inv ? 'I' : 'P') : (inv ? 'i' : 'S'));
    s += "\tcm:i:"; put_int(s, h.cnt); s += "\ts1:i:"; put_int(s, e.chain_score);
    if (pri) { s += "\ts2:i:"; put_int(s, 0); }
    if (has_p) {
        int n_gapo = 0, n_gap = 0;
        for (uint32_t i = 0; i < e.n_cigar; ++i) { const int op = (int)(cig[i] & 0xf), len = (int)(cig[i] >> 4); if (op == 1 || op == 2) { ++n_gapo; n_gap += len; } }
        const double ide = (double)h.mlen / (h.blen - n_gap + n_gapo);
        const double div = 1.0 - ide;
        char buf[16];
        if (div == 0.0) { buf[0] = '0'; buf[1] = 0; } else snprintf(buf, 16, "%.4f", 1.0 - ide);
        s += "\tde:f:"; s += buf;
    } else {
        const float div = est_div(h, e, q, qlen, tlen);
        if (div >= 0.0f && div <= 1.0f) {
            char buf[16];
            if (div == 0.0f) { buf[0] = '0'; buf[1] = 0; } else snprintf(buf, 16, "%.4f", div);
            s += "\tdv:f:"; s += buf;
        }
    }
    const int split = (int)(e.flags >> 2 & 3u);
    if (split) { s += "\tzd:i:"; put_int(s, split); }
    if (q.rep_len >= 0) { s += "\trl:i:"; put_int(s, q.rep_len); }
    if (has_p && P.out_cg) {
        s += "\tcg:Z:";
        for (uint32_t k = 0; k < e.n_cigar; ++k) { put_int(s, (int)(cig[k] >> 4)); s.push_back("MIDNSHP=XB"[cig[k] & 0xf]); }
    }
}

} // namespace ndpaf
