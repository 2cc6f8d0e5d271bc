// Microbenchmark: tcgen05.ld throughput per SM as a function of the number of warps reading (run on the GPU box).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_bw tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../alphaquoridorgnn_b200/csrc/tc_common.cuh"
using namespace aqtc;

__global__ void __launch_bounds__(1024, 1) tmem_ld_kernel(int iters, int mode, long long *cycles, float *sink) {
    __shared__ uint32_t tmem_base;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const int warp = threadIdx.x >> 5;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32) % 480;
    float acc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        float v[32];
        if (mode == 0) {
            tmem_ld32(taddr, v);
        } else {  // x16 loads
            uint32_t r[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                           "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(r[k]);
            for (int k = 16; k < 32; ++k) v[k] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < 32; k += 8) acc += v[k];
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512u) : "memory");
}

int main() {
    long long *cyc;
    float *sink;
    cudaMalloc(&cyc, 148 * sizeof(long long));
    cudaMalloc(&sink, 4);
    const int iters = 2000;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {1, 4, 8, 16, 20, 32}) {
            tmem_ld_kernel<<<148, warps * 32, 0>>>(iters, mode, cyc, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            const double bytes = (double)iters * warps * 32 * (mode == 0 ? 32 : 16) * 4;
            printf("mode x%d warps %2d: %lld cycles, %.1f B/cycle/SM, %.1f cycles per warp-load\n", mode == 0 ? 32 : 16, warps, h[0],
                   bytes / h[0], (double)h[0] / iters);
        }
    return 0;
}
