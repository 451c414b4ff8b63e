// probe: predicate semantics of __vibmax_s16x2 on sm_100a (ties, signs)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const unsigned *a, const unsigned *b, unsigned *o, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool ph, pl;
    unsigned m = __vibmax_s16x2(a[i], b[i], &ph, &pl);
    o[2 * i] = m; o[2 * i + 1] = (ph ? 2u : 0u) | (pl ? 1u : 0u);
}
int main() {
    const int n = 6;
    unsigned ha[n] = {0x00050005u, 0x00030007u, 0xfffe0002u, 0x00000000u, 0xfff0fff0u, 0x00010000u};
    unsigned hb[n] = {0x00050005u, 0x00070003u, 0x0002fffeu, 0x0000ffffu, 0xfff0fff1u, 0x00000001u};
    unsigned *a, *b, *o, ho[2 * n];
    cudaMalloc(&a, sizeof ha); cudaMalloc(&b, sizeof hb); cudaMalloc(&o, sizeof ho);
    cudaMemcpy(a, ha, sizeof ha, cudaMemcpyHostToDevice); cudaMemcpy(b, hb, sizeof hb, cudaMemcpyHostToDevice);
    k<<<1, 32>>>(a, b, o, n);
    cudaMemcpy(ho, o, sizeof ho, cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; ++i) {
        short ahi = ha[i] >> 16, alo = ha[i] & 0xffff, bhi = hb[i] >> 16, blo = hb[i] & 0xffff;
        printf("a=(%d,%d) b=(%d,%d) max=%08x pred_hi=%u pred_lo=%u  expect(a>=b): hi=%d lo=%d\n", ahi, alo, bhi, blo, ho[2 * i], ho[2 * i + 1] >> 1, ho[2 * i + 1] & 1,
               ahi >= bhi, alo >= blo);
    }
    return 0;
}
