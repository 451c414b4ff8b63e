// ovlsort_harness.cc -- TEST INFRASTRUCTURE.  Runs the per-read stage of the GPU `ovl_sort` (nextdenovo_b200/csrc/nd_ovlsort.cuh,
// __host__ __device__) on the CPU around a plain host front end (decode, mirror, stable sort) and back end (encode, .bl), so that the
// exact-semantics logic can be compared with the unmodified reference binary (oracle/_ref/ovl_sort) on a box without a GPU.  The
// product path does the same stages on the GPU (ndg_ovl_sort); the product never runs this file.
//
// usage: ovlsort_harness IDX OUT [-H] [-k N] [-l N] IN.ovl [IN.ovl ...]      writes OUT and OUT.bl
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <string>
#include <algorithm>
#include <unordered_map>
#include "../nextdenovo_b200/csrc/nd_ovlsort.cuh"

static int put_varint(uint8_t *d, uint32_t v)            // big-endian base-128, continuation bit on all but the last byte (ref: lib/ovl.c:10-29)
{
    uint8_t t[5]; int m = 0;
    for (int j = 28; j >= 0; j -= 7) { const uint32_t x = v >> j & 127; if (x > 0 || m > 0) t[m++] = (uint8_t)(x | 128); }
    if (m == 0) t[m++] = 128;
    t[m - 1] &= 127;
    memcpy(d, t, (size_t)m);
    return m;
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: ovlsort_harness IDX OUT [-H] [-k N] [-l N] IN.ovl ...\n"); return 2; }
    ndos_par P; P.max_bin_cov = 40; P.max_flank_len = 300; P.is_hq = 0; P.min_seed_len = 0; P.with_bl = 1;
    std::vector<const char *> files;
    for (int i = 3; i < argc; ++i) {
        if (!strcmp(argv[i], "-H")) P.is_hq = 1;
        else if (!strcmp(argv[i], "-k")) P.max_bin_cov = atoi(argv[++i]);
        else if (!strcmp(argv[i], "-l")) P.max_flank_len = atoi(argv[++i]);
        else files.push_back(argv[i]);
    }
    P.with_bl = P.max_flank_len != 0;
    std::unordered_map<uint32_t, uint32_t> seeds;          // ref: util/ovl_sort.c:106-133
    {
        FILE *f = fopen(argv[1], "r"); if (!f) return 1;
        unsigned name, len; unsigned long off;
        while (fscanf(f, "%u\t%lu\t%u", &name, &off, &len) == 3)
            if (seeds.emplace(name, len).second) if ((int)len < P.min_seed_len || !P.min_seed_len) P.min_seed_len = (int)len;
        fclose(f);
    }
    std::vector<ndos_rec> R;
    for (const char *fn : files) {                          // ref: util/ovl_sort.c:977-1046 (sort_ovl_file)
        FILE *f = fopen(fn, "rb"); if (!f) { fprintf(stderr, "cannot open %s\n", fn); return 1; }
        std::vector<uint8_t> b; { uint8_t buf[1 << 16]; size_t n; while ((n = fread(buf, 1, sizeof buf, f)) > 0) b.insert(b.end(), buf, buf + n); }
        fclose(f);
        uint32_t pq = 0, pt = 0, v[8]; int nv = 0; uint64_t cat = 0;
        int key1 = 5, key5 = 5;
        for (size_t i = 0; i < b.size(); ++i) {
            cat = cat << 7 | (b[i] & 127);
            if (b[i] < 128) { v[nv++] = (uint32_t)cat; cat = 0; }
            if (nv < 8) continue;
            nv = 0;
            v[0] = (v[1] & 2) ? pq - v[0] : pq + v[0]; pq = v[0];
            v[4] = (v[1] & 4) ? pt - v[4] : pt + v[4]; pt = v[4];
            v[6] = (v[1] & 8) ? v[5] + v[3] + v[6] : v[5] + v[3] - v[6];
            v[1] &= 1; v[3] += v[2];
            if (v[0] == v[4]) continue;
            if (v[3] - v[2] < 500 || v[6] - v[5] < 500) continue;
            auto it = seeds.find(v[0]);
            if (key1 && it != seeds.end() && it->second >= v[3]) R.push_back(ndos_rec{ v[0], v[1], v[2], v[3] - 1, v[4], v[5], v[6] - 1, v[7] });
            else if (key1) --key1;
            it = seeds.find(v[4]);
            if (key5 && it != seeds.end() && it->second >= v[6]) R.push_back(ndos_rec{ v[4], v[1], v[5], v[6] - 1, v[0], v[2], v[3] - 1, v[7] });
            else if (key5) --key5;
        }
    }
    std::stable_sort(R.begin(), R.end(), ndos_less);
    std::vector<uint8_t> out; std::string bl;
    uint32_t st[2] = { 0, 0 };
    std::vector<ndos_rec> q; std::vector<uint16_t> bins; std::vector<uint32_t> acc; std::vector<uint8_t> dead;
    for (size_t i = 0; i < R.size();) {
        size_t j = i; while (j < R.size() && R[j].qname == R[i].qname) ++j;
        const uint32_t qlen = seeds[R[i].qname];
        q.clear();
        q.push_back(ndos_rec{ R[i].qname, 0, 0, qlen - 1, R[i].qname, 0, qlen - 1, 0 });      // the read's own record opens it (ref: :945-953)
        q.insert(q.end(), R.begin() + (long)i, R.begin() + (long)j);
        bins.assign((qlen >> NDOS_BIN_OFFSET) + 4, 0); acc.assign(q.size(), 0); dead.assign(q.size(), 0);
        uint32_t n_out = 0;
        const char lab = ndos_one_read(P, qlen, q.data(), (uint32_t)q.size(), bins.data(), acc.data(), dead.data(), &n_out);
        for (size_t x = 0; x < q.size(); ++x) {
            if (dead[x] || !q[x].qe) continue;
            const ndos_rec &o = q[x];
            uint32_t f[8], flags = o.rev;
            const uint32_t ql = o.qe - o.qs, tl = o.te - o.ts;
            if (o.qname >= st[0]) f[0] = o.qname - st[0]; else { flags |= 2; f[0] = st[0] - o.qname; }
            if (o.tname >= st[1]) f[4] = o.tname - st[1]; else { flags |= 4; f[4] = st[1] - o.tname; }
            if (ql >= tl) f[6] = ql - tl; else { flags |= 8; f[6] = tl - ql; }
            st[0] = o.qname; st[1] = o.tname;
            f[1] = flags; f[2] = o.qs; f[3] = ql; f[5] = o.ts; f[7] = o.match;
            uint8_t tmp[48]; int nb = 0;
            for (int k = 0; k < 8; ++k) nb += put_varint(tmp + nb, f[k]);
            out.insert(out.end(), tmp, tmp + nb);
        }
        if (lab) { char line[32]; snprintf(line, sizeof line, "%u %c\n", R[i].qname, lab); bl += line; }
        i = j;
    }
    FILE *fo = fopen(argv[2], "wb"); if (!fo) return 1;
    fwrite(out.data(), 1, out.size(), fo); fclose(fo);
    if (P.with_bl) { std::string bn = std::string(argv[2]) + ".bl"; FILE *fb = fopen(bn.c_str(), "w"); fwrite(bl.data(), 1, bl.size(), fb); fclose(fb); }
    fprintf(stderr, "records in %zu, bytes out %zu, labels %zu\n", R.size(), out.size(), (size_t)std::count(bl.begin(), bl.end(), '\n'));
    return 0;
}
