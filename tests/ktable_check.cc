// ktable_check.cc -- TEST INFRASTRUCTURE.  nds2::KTable (nextdenovo_b200/csrc/nd_step2.hpp) against the reference's own khash
// (util/khash.h, included from the reference tree at compile time; the test is skipped where that tree is absent): after the
// same insertions, the occupied buckets hold the same keys in the same order -- the order out_bl() prints the .bl lines in.
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include "khash.h"
#include "../nextdenovo_b200/csrc/nd_step2.hpp"
KHASH_MAP_INIT_INT(t32, int)

static uint64_t rng_state = 88172645463325252ULL;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }

int main()
{
    int bad = 0, cases = 0;
    for (int pattern = 0; pattern < 6; ++pattern)
        for (int n : {1, 3, 4, 5, 12, 13, 50, 197, 1000, 5000, 40000}) {
            khash_t(t32) *h = kh_init(t32);
            nds2::KTable T;
            for (int i = 0; i < n; ++i) {
                uint32_t k;
                switch (pattern) {
                    case 0: k = (uint32_t)i; break;                       // read ids 0..n-1
                    case 1: k = (uint32_t)(i * 16); break;                // colliding low bits
                    case 2: k = rnd(); break;
                    case 3: k = rnd() % (uint32_t)(n / 2 + 1); break;      // repeats
                    case 4: k = (uint32_t)(n - i) * 1024u; break;
                    default: k = (rnd() & 0xff) << 20 | (uint32_t)i % 7; break;
                }
                int absent; khint_t it = kh_put(t32, h, k, &absent); if (absent) kh_val(h, it) = i;
                bool a2; int32_t *v = T.put(k, &a2); if (a2) *v = i;
                if ((absent != 0) != a2) ++bad;
            }
            ++cases;
            if (kh_n_buckets(h) != T.n_buckets || kh_size(h) != T.size) { ++bad; continue; }
            for (khint_t b = 0; b < kh_end(h); ++b) {
                const bool u1 = kh_exist(h, b), u2 = T.used[b] != 0;
                if (u1 != u2 || (u1 && (kh_key(h, b) != T.keys[b] || kh_val(h, b) != T.val[b]))) { ++bad; break; }
            }
            kh_destroy(t32, h);
        }
    printf("cases=%d bad=%d\n", cases, bad);
    return bad != 0;
}
