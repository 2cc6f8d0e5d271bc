// Microbenchmark: SM throughput of the fp32 -> packed 16-bit conversions the trunk's epilogues use, against an integer-pipe
// alternative for bf16 (add half an ulp, take the upper halves with one byte permute).  16 warps per SM, 8 independent
// chains per thread; reports cycles per warp instruction per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int kMode>
__global__ void __launch_bounds__(512, 1) k(int iters, long long *out, uint32_t *sink, float seed) {
    float a[8], b[8];
    uint32_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 0.001f + i; b[i] = seed * 0.5f + i; acc[i] = 0u; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t d;
            if (kMode == 0) asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;\n" : "=r"(d) : "f"(b[i]), "f"(a[i]));
            else if (kMode == 1) asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;\n" : "=r"(d) : "f"(b[i]), "f"(a[i]));
            else if (kMode == 2) {   // relu, + half ulp, upper halves
                const uint32_t x = __float_as_uint(fmaxf(a[i], 0.f)) + 0x8000u, y = __float_as_uint(fmaxf(b[i], 0.f)) + 0x8000u;
                d = __byte_perm(x, y, 0x7632);
            } else {                 // plain FFMA for reference
                d = __float_as_uint(fmaf(a[i], b[i], 1.0f));
            }
            acc[i] ^= d;
            a[i] = __uint_as_float((acc[i] & 0x007FFFFFu) | 0x3F800000u);   // keep a dependency so nothing is hoisted
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= acc[i];
    sink[blockIdx.x * 512 + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
}
template <int kMode>
void run(const char *name, long long *d, uint32_t *sink) {
    long long h;
    const int iters = 2000;
    k<kMode><<<148, 512>>>(iters, d, sink, 1.25f);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    // 16 warps per SM = 4 per sub-partition, each iters * 8 conversions
    printf("%-44s %.2f cycles per conversion instruction per sub-partition (loop body incl. 2-3 ALU ops of overhead)\n", name,
           (double)h / (iters * 8.0 * 4.0));
}
int main() {
    long long *d; uint32_t *sink;
    cudaMalloc(&d, 16); cudaMalloc(&sink, 148 * 512 * 4);
    run<3>("FFMA (reference: loop overhead)", d, sink);
    run<0>("cvt.rn.relu.bf16x2.f32 (F2FP)", d, sink);
    run<1>("cvt.rn.satfinite.f16x2.f32 (F2FP)", d, sink);
    run<2>("2 FMNMX + 2 IADD + PRMT (bf16, ties away)", d, sink);
    return 0;
}
