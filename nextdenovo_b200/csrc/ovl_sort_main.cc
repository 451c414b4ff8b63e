// ovl_sort_main.cc -- `ovl_sort`-compatible executable over ndg_ovl_sort (libndgpu.so): the reference's command line
// (ref: util/ovl_sort.c:1048-1061 usage, :1063-1091 option parsing) -- `ovl_sort -i seeds.idx [-H] [-m MEM] [-t N] [-k DEPTH] [-l LEN]
// -o OUT [-d TMPDIR] input.fofn` -- with the whole job done on the GPU in one pass: no spill files, so -m, -t and -d are accepted and
// have nothing to steer.  Writes OUT and, unless -l 0, OUT.bl (ref: :199-210).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <chrono>
#include <unistd.h>
#include "ndgpu.h"

static bool slurp(const char *fn, std::vector<uint8_t> &b)
{
    FILE *f = fopen(fn, "rb"); if (!f) return false;
    uint8_t buf[1 << 16]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) b.insert(b.end(), buf, buf + n);
    fclose(f); return true;
}

static int usage()
{
    fprintf(stderr, "Usage: ovl_sort [options] input.fofn\nSort and merge ovl files (GPU)\n\nOptions:\n  -i  query idx file [required]\n  -H  high quality reads, such as HiFi [0]\n"
                    "  -m  (accepted, unused: the GPU pass holds everything in HBM)\n  -t  (accepted, unused)\n  -k  max depth of each overlap [40]\n  -l  max over hang length to filter [300]\n"
                    "  -o  output file name [required]\n  -d  (accepted, unused: no temporary files)\n  -D  GPU device [0]\n");
    return 1;
}

int main(int argc, char **argv)
{
    const auto t0 = std::chrono::steady_clock::now();
    const char *idx = nullptr, *out = nullptr; int device = 0;
    ndg_ovlsort_opt_t o = { 40, 300, 0 };
    int c;
    while ((c = getopt(argc, argv, "i:m:t:k:o:d:l:HD:")) != -1) {
        switch (c) {
            case 'i': idx = optarg; break;
            case 't': if (atoi(optarg) < 2) { fprintf(stderr, "ovl_sort: -t must be >= 2\n"); return 1; } break;   // the reference asserts this
            case 'k': o.max_bin_cov = atoi(optarg); break;
            case 'o': out = optarg; break;
            case 'l': o.max_flank_len = atoi(optarg); break;
            case 'H': o.is_hq = 1; break;
            case 'D': device = atoi(optarg); break;
            case 'm': case 'd': break;
            default: return usage();
        }
    }
    if (optind + 1 > argc || !out || !idx) return usage();
    std::vector<uint32_t> sn, sl;
    {
        FILE *f = fopen(idx, "r");
        if (!f) { fprintf(stderr, "[WARNING] Failed open index file %s!\n", idx); return 1; }
        unsigned name, len; unsigned long off;
        while (fscanf(f, "%u\t%lu\t%u", &name, &off, &len) == 3) { sn.push_back(name); sl.push_back(len); }
        fclose(f);
    }
    std::vector<std::vector<uint8_t>> files;
    {
        FILE *f = fopen(argv[argc - 1], "r");
        if (!f) { fprintf(stderr, "[WARNING] Failed open input overlap file list %s!\n", argv[argc - 1]); return 1; }
        char *line = nullptr; size_t cap = 0; ssize_t n;
        while ((n = getline(&line, &cap, f)) != -1) {                       // ref: :147-157: empty lines and '#' lines are skipped
            if (n == 1 || line[0] == '#') continue;
            if (line[n - 1] == '\n') line[n - 1] = 0;
            files.emplace_back();
            if (!slurp(line, files.back())) { fprintf(stderr, "[WARNING] Failed open overlap file %s!\n", line); return 1; }
        }
        free(line); fclose(f);
    }
    ndg_ctx_t *g = nullptr;
    if (ndg_init(device, &g) != NDG_OK) { fprintf(stderr, "[ERROR] no usable B200 (sm_100) device %d; there is no CPU fallback\n", device); return 1; }
    std::vector<const uint8_t *> ptr; std::vector<uint64_t> sz;
    for (auto &b : files) { ptr.push_back(b.data()); sz.push_back(b.size()); }
    const uint8_t *ovl = nullptr; uint64_t nb = 0, nbl = 0, n_in = 0, n_out = 0; const char *bl = nullptr;
    const auto t1 = std::chrono::steady_clock::now();
    if (ndg_ovl_sort(g, ptr.data(), sz.data(), (int32_t)ptr.size(), sn.data(), sl.data(), sn.size(), &o, &ovl, &nb, &bl, &nbl, &n_in, &n_out)) {
        fprintf(stderr, "[ERROR] %s\n", ndg_last_error(g)); return 1;
    }
    const double t_call = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    FILE *fo = fopen(out, "wb");
    if (!fo || (nb && fwrite(ovl, 1, (size_t)nb, fo) != (size_t)nb) || fclose(fo) != 0) { fprintf(stderr, "[WARNING] Failed create overlap file %s!\n", out); return 1; }
    if (o.max_flank_len) {
        const std::string bn = std::string(out) + ".bl";
        FILE *fb = fopen(bn.c_str(), "w");
        if (!fb || (nbl && fwrite(bl, 1, (size_t)nbl, fb) != (size_t)nbl) || fclose(fb) != 0) { fprintf(stderr, "[WARNING] Failed create rid file %s!\n", bn.c_str()); return 1; }
    }
    fprintf(stderr, "[ovl_sort-ndgpu] %zu files, %llu records in, %llu out; ndg_ovl_sort %.3f s (upload, all kernels, download), process %.3f s\n", files.size(), (unsigned long long)n_in,
            (unsigned long long)n_out, t_call, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    ndg_destroy(g);
    return 0;
}
