// Microbenchmark: tensor-pipe cycles per tcgen05.mma (kind::f16, M = 128, K = 16, A in tensor memory) by N, issued back to back from
// warp-uniform code (descriptors in uniform registers, 8 MMAs per loop iteration): is there a per-instruction floor below N = 96?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../alphaquoridorgnn_b200/csrc/tc_common.cuh"
using namespace aqtc;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}
template <int N, bool kSS, int kChains>
__global__ void __launch_bounds__(128, 1) k(int iters, long long *out) {
    extern __shared__ unsigned char raw[];
    unsigned char *sm = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    __shared__ unsigned long long mbar;
    __shared__ uint32_t tb_s;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(sm)[i] = 0;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&mbar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tb_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tb = __shfl_sync(0xffffffffu, tb_s, 0);
    const uint32_t b_s = __shfl_sync(0xffffffffu, smem_u32(sm), 0) + 32768u, a_s = b_s - 32768u;
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24) | ((uint32_t)(N >> 3) << 17);
    if ((threadIdx.x >> 5) == 0) {
        long long t0 = 0, t2 = 0;
        uint32_t pred;
        asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
        if (pred) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (kSS) mma_bf16(tb + (u % kChains) * 96, desc_sw128(a_s + (u & 3) * 32), desc_sw128(b_s + (u & 3) * 32), idesc, 1u);
                    else mma_ts(tb + (u % kChains) * 96, tb + 448 + u * 8, desc_sw128(b_s + (u & 3) * 32), idesc);
                }
            }
            mma_commit(smem_u32(&mbar));
            mbar_wait(smem_u32(&mbar), 0);
            t2 = clock64();
            if (blockIdx.x == 0) out[0] = t2 - t0;
        }
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "r"(512u) : "memory");
}
template <int N, bool kSS, int kChains = 4>
void run(long long *d) {
    long long h;
    const int iters = 500;
    cudaFuncSetAttribute(k<N, kSS, kChains>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
    k<N, kSS, kChains><<<148, 128, 70 * 1024>>>(iters, d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error N %d\n", N); return; }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%s N=%3d, %d accumulator(s) in rotation: %.1f cycles per MMA (ideal %d)\n", kSS ? "SS" : "TS", N, kChains, (double)h / (iters * 8), N / 2);
}
int main() {
    long long *d;
    cudaMalloc(&d, 16);
    run<16, false>(d); run<32, false>(d); run<48, false>(d); run<64, false>(d); run<96, false>(d); run<128, false>(d);
    run<32, false, 1>(d); run<48, false, 1>(d); run<96, false, 1>(d); run<32, false, 2>(d); run<96, false, 2>(d); run<48, true, 1>(d); run<48, true, 2>(d);
    run<16, true>(d); run<32, true>(d); run<48, true>(d); run<64, true>(d); run<96, true>(d); run<128, true>(d);
    return 0;
}
