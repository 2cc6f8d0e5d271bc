// Microbenchmark: cycles per tcgen05.mma for the shapes the trunk kernels use (one CTA per SM, one issuing thread).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../alphaquoridorgnn_b200/csrc/tc_common.cuh"
using namespace aqtc;

__device__ inline void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, int tf32) {
    if (tf32) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
    else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
__global__ void __launch_bounds__(128, 1) k(int mode, int N, int iters, long long *out, int nchains) {
    extern __shared__ unsigned char raw[];
    unsigned char *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ unsigned long long mbar;
    __shared__ uint32_t tb;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(sm)[i] = 0;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tb)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (threadIdx.x == 0) {
        const uint32_t a_s = smem_u32(sm), b_s = smem_u32(sm) + 32768;
        const uint32_t f16 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24) | ((uint32_t)(N >> 3) << 17);
        const uint32_t t32 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24) | ((uint32_t)(N >> 3) << 17);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t dd = tb + (uint32_t)(i % nchains) * (uint32_t)(N <= 48 ? 48 : 96);
            if (mode == 0) mma_bf16(dd, desc_sw128(a_s + (i & 3) * 32), desc_sw128(b_s + (i & 3) * 32), f16, 1u);      // SS f16
            else if (mode == 1) mma_ts(dd, tb + 448 + (i & 7) * 8, desc_sw128(b_s + (i & 3) * 32), f16, 0);            // TS f16
            else mma_ts(dd, tb + 448 + (i & 7) * 8, desc_sw128(b_s + (i & 3) * 32), t32, 1);                            // TS tf32
        }
        const long long t1 = clock64();
        mma_commit(smem_u32(&mbar));
        mbar_wait(smem_u32(&mbar), 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "r"(512u) : "memory");
}
int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
    const char *names[3] = {"SS f16 (A, B smem)", "TS f16 (A TMEM)", "TS tf32 (A TMEM)"};
    for (int mode = 0; mode < 3; ++mode)
        for (int nch : {1, 4}) for (int N : {16, 32, 48, 64, 96}) {
            const int iters = 2000;
            k<<<148, 128, 70 * 1024>>>(mode, N, iters, d, nch);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error mode %d N %d\n", mode, N); return 1; }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%-20s chains=%d N=%3d: issue %.1f cycles/MMA, complete %.1f cycles/MMA\n", names[mode], nch, N, (double)h[0] / iters, (double)h[1] / iters);
        }
    return 0;
}
