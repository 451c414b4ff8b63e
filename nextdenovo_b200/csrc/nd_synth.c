/* nd_synth.c -- seeded synthetic long-read generator for benchmarks and tests (host side, not on the
 * hot path).  Writes reads straight into the .2bit word layout (ref: lib/bseq.c:114-139).
 * Model (BASELINE.json config 1): random genome, reads of fixed length sampled
 * uniformly with random strand, per-base errors sub/ins/del, PRNG splitmix64. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t sm64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* genome as one 2-bit code per byte */
void nds_genome(uint8_t *g, uint64_t G, uint64_t seed)
{
    uint64_t i;
#pragma omp parallel for schedule(static)
    for (i = 0; i < (G + 31) / 32; ++i) {
        uint64_t s = seed ^ (i * 0xD1342543DE82EF95ULL), r = sm64(&s), j, e = (i + 1) * 32 < G ? (i + 1) * 32 : G;
        for (j = i * 32; j < e; ++j) { g[j] = (uint8_t)(r & 3); r >>= 2; }
    }
}

/* reads [r0, r0+n): words_out holds n * ceil(L/16) u32.  rates in 1/65536 units. */
void nds_reads(const uint8_t *g, uint64_t G, int64_t r0, int64_t n, int L, uint32_t sub, uint32_t ins, uint32_t del,
               uint64_t seed, uint32_t *words_out, uint64_t *start_out, uint8_t *strand_out)
{
    const int nw = (L + 15) / 16;
    int64_t r;
#pragma omp parallel
    {
        uint8_t *buf = (uint8_t*)malloc((size_t)L + 16);
#pragma omp for schedule(dynamic, 64)
        for (r = 0; r < n; ++r) {
            uint64_t s = seed ^ ((uint64_t)(r0 + r) * 0xA24BAED4963EE407ULL) ^ 0x5851F42D4C957F2DULL;
            uint64_t span = (uint64_t)L + L / 4 + 64;
            uint64_t gp = span < G ? sm64(&s) % (G - span) : 0, gp0 = gp;
            int strand = (int)(sm64(&s) & 1), m = 0, k;
            uint32_t *w = words_out + (size_t)r * nw;
            while (m < L) {
                uint64_t x = sm64(&s);
                uint32_t u = (uint32_t)(x & 0xffff);
                if (u < ins) buf[m++] = (uint8_t)(x >> 16 & 3);
                else if (u < ins + del) ++gp;
                else if (u < ins + del + sub) { buf[m++] = (uint8_t)((g[gp % G] + 1 + (x >> 16) % 3) & 3); ++gp; }
                else { buf[m++] = g[gp % G]; ++gp; }
            }
            if (strand) for (k = 0; k < L / 2; ++k) { uint8_t a = 3 - buf[k], b = 3 - buf[L - 1 - k]; buf[k] = b; buf[L - 1 - k] = a; }
            if (strand && (L & 1)) buf[L / 2] = 3 - buf[L / 2];
            memset(w, 0, (size_t)nw * 4);
            for (k = 0; k < L; ++k) w[k >> 4] |= (uint32_t)buf[k] << (30 - 2 * (k & 15));
            if (start_out) start_out[r] = gp0;
            if (strand_out) strand_out[r] = (uint8_t)strand;
        }
        free(buf);
    }
}
