// tf32 variant: the aggregation MMA (kind::tf32) reads its A operand straight from the fp32 accumulator columns of MMA1.
// Probe for the operand layouts the tensor-core aggregation design relies on (run on the GPU box):
//   MMA1 (TS mode): A = W [128 out][128 in] bf16 in TMEM (two bf16 per 32-bit column), B = X^T [K = 128 feat][N = 96 nodes]
//                   MN-major SWIZZLE_64B in shared memory, D = Z^T [128 lanes][96 columns].
//   MMA2 (SS mode): A = Z^T [M = 128 feat][K = 96 nodes] K-major SWIZZLE_64B (the SAME bytes as the MN-major B tile above),
//                   B = banded adjacency blocks [48 out nodes][64 in nodes] K-major SWIZZLE_128B, hi and lo parts,
//                   plus one extra K step that adds the bias through a ones column.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../alphaquoridorgnn_b200/csrc/tc_common.cuh"
using namespace aqtc;

constexpr int kNodes = 96, kFeat = 128;
constexpr uint32_t kFmBlock = 16 * 512;  // feature-major tile: [3 node blocks of 32][16 atoms of 8 features][8 rows][64 B]

// byte offset of (feature f, node v) in the feature-major SWIZZLE_64B tile
__host__ __device__ inline uint32_t fm_off(int f, int v) {
    return (uint32_t)(v >> 5) * kFmBlock + (uint32_t)(f >> 3) * 512u + (uint32_t)(f & 7) * 64u +
           (uint32_t)((((v & 31) >> 3) ^ ((f & 7) >> 1)) << 4) + (uint32_t)(v & 7) * 2u;
}
__device__ inline uint64_t desc_fm_mn(uint32_t saddr) {  // MN-major SW64: LBO = node-block stride, SBO = 8-feature atom stride
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(kFmBlock >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ inline uint64_t desc_fm_k(uint32_t saddr) {   // K-major SW64: SBO = 512 (8 rows x 64 B)
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ inline void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ inline void tmem_st32(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
                   "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
                   "r"(r[30]), "r"(r[31]) : "memory");
}

struct Smem {
    unsigned char fm[3 * kFmBlock];        // 24 KB feature-major tile
    unsigned char ah[2][2 * 48 * 128];     // adjacency (tf32 = fp32 words), two blocks of [48][64]: two K-blocks of 32 elements each
    unsigned char bias[128 * 32];          // [128][16] K-major SW32: (b_hi, b_lo, 0...)
    unsigned char ones[48 * 32];           // [48][16] K-major SW32: (1, 1, 0...)
    unsigned long long mbar;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __nv_bfloat16 *W, const __nv_bfloat16 *X /*[96 nodes][128 feat]*/, const float *A /*[96][96] out,in*/,
             const float *bias, float *out1 /*[128][96]*/, float *out2 /*[128][96]*/, int bmode) {
    extern __shared__ unsigned char raw[];
    Smem &sm = *reinterpret_cast<Smem *>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&sm.mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&sm.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tbase = sm.tmem_base;
    const uint32_t lane_off = (uint32_t)((tid >> 5) * 32) << 16;
    // ---- W row `tid` -> TMEM columns [384, 448): column c holds (k = 2c, k = 2c + 1)
    {
        uint32_t r[32];
        for (int half = 0; half < 2; ++half) {
            for (int c = 0; c < 32; ++c) r[c] = reinterpret_cast<const uint32_t *>(W + tid * 128)[half * 32 + c];
            tmem_st32(tbase + lane_off + 384 + half * 32, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    // ---- X^T into the feature-major tile: thread = feature, 96 nodes contiguous
    for (int v = 0; v < kNodes; ++v) *reinterpret_cast<__nv_bfloat16 *>(sm.fm + fm_off(tid, v)) = X[v * kFeat + tid];
    // ---- adjacency band blocks (hi/lo), bias and ones tiles
    for (int i = tid; i < 2 * 48 * 64; i += 128) {
        const int blk = i / (48 * 64), r = (i / 64) % 48, kl = i % 64;
        const int n = blk * 48 + r, k = (blk ? 32 : 0) + kl;
        const uint32_t off = (uint32_t)(kl >> 5) * (48 * 128) + (uint32_t)r * 128u + (uint32_t)(((((kl & 31) >> 2) ^ (r & 7)) << 4)) + (uint32_t)(kl & 3) * 4u;
        *reinterpret_cast<float *>(sm.ah[blk] + off) = A[n * 96 + k];
    }
    {
        const float b = bias[tid];
        const __nv_bfloat16 hi = __float2bfloat16_rn(b), lo = __float2bfloat16_rn(b - __bfloat162float(hi));
        uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
        c0.x = (uint32_t)bf16_bits(__bfloat162float(hi)) | ((uint32_t)bf16_bits(__bfloat162float(lo)) << 16);
        *reinterpret_cast<uint4 *>(sm.bias + sw32_chunk(tid, 0)) = c0;
        *reinterpret_cast<uint4 *>(sm.bias + sw32_chunk(tid, 1)) = c1;
        if (tid < 48) {
            uint4 o0 = make_uint4(0x3F803F80u, 0, 0, 0);
            *reinterpret_cast<uint4 *>(sm.ones + sw32_chunk(tid, 0)) = o0;
            *reinterpret_cast<uint4 *>(sm.ones + sw32_chunk(tid, 1)) = c1;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    const uint32_t bar = smem_u32(&sm.mbar);
    // ---- MMA1: D[0,96) = W (TMEM) x X^T (MN-major B)
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(96 >> 3) << 17) | ((128u >> 4) << 24);
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        for (int k = 0; k < 8; ++k)
            mma_ts(tbase, tbase + 384 + k * 8, desc_fm_mn(smem_u32(sm.fm) + k * 1024), idesc1, k > 0);
        mma_commit(bar);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    float z[96];
    for (int cb = 0; cb < 3; ++cb) tmem_ld32(tbase + lane_off + cb * 32, z + cb * 32);
    for (int v = 0; v < 96; ++v) out1[tid * 96 + v] = z[v];
    __syncthreads();
    // ---- MMA2 (kind::tf32, TS): A = accumulator columns of MMA1 (K window of 64 nodes), B = tf32 adjacency block, D = columns 96..191
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(48 >> 3) << 17) | ((128u >> 4) << 24);
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        for (int blk = 0; blk < 2; ++blk) {
            const uint32_t d = tbase + 96 + blk * 48;
            for (int s = 0; s < 8; ++s) {   // 64 in-nodes = 8 K steps of 8
                const uint32_t a_t = tbase + (blk ? 32 : 0) + s * 8;
                const uint64_t b = desc_sw128(smem_u32(sm.ah[blk]) + (s >> 2) * (48 * 128) + (s & 3) * 32);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                             ::"r"(d), "r"(a_t), "l"(b), "r"(idesc2), "r"(s ? 1u : 0u) : "memory");
            }
        }
        mma_commit(bar);
    }
    mbar_wait(bar, 1);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    for (int cb = 0; cb < 3; ++cb) tmem_ld32(tbase + lane_off + 96 + cb * 32, z + cb * 32);
    for (int v = 0; v < 96; ++v) out2[tid * 96 + v] = z[v];
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tbase), "r"(512u) : "memory");
}

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
static float tf(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main(int argc, char **argv) {
    std::vector<__nv_bfloat16> W(128 * 128), X(96 * 128);
    std::vector<float> A(96 * 96, 0.f), bias(128);
    srand(1);
    auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (auto &w : W) w = __float2bfloat16_rn(rnd() * 0.2f);
    for (auto &x : X) x = __float2bfloat16_rn(rnd());
    for (int n = 0; n < 81; ++n)  // 5-point stencil pattern with arbitrary coefficients
        for (int k : {n - 9, n - 1, n, n + 1, n + 9})
            if (k >= 0 && k < 81) A[n * 96 + k] = 0.2f + 0.3f * fabsf(rnd());
    for (auto &b : bias) b = rnd();
    __nv_bfloat16 *dW, *dX; float *dA, *db, *o1, *o2;
    cudaMalloc(&dW, W.size() * 2); cudaMalloc(&dX, X.size() * 2); cudaMalloc(&dA, A.size() * 4); cudaMalloc(&db, 512);
    cudaMalloc(&o1, 128 * 96 * 4); cudaMalloc(&o2, 128 * 96 * 4);
    cudaMemcpy(dW, W.data(), W.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, bias.data(), 512, cudaMemcpyHostToDevice);
    const size_t smem = sizeof(Smem) + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int bmode = argc > 1 ? atoi(argv[1]) : 0;
    probe_kernel<<<1, 128, smem>>>(dW, dX, dA, db, o1, o2, bmode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    printf("bmode %d (%s)\n", bmode, bmode ? "A bf16 x B fp16 single part" : "bf16 hi/lo");
    std::vector<float> h1(128 * 96), h2(128 * 96);
    cudaMemcpy(h1.data(), o1, h1.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(h2.data(), o2, h2.size() * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0, m1 = 0, m2 = 0;
    std::vector<float> Z(128 * 96);
    for (int f = 0; f < 128; ++f)
        for (int v = 0; v < 96; ++v) {
            double s = 0;
            for (int k = 0; k < 128; ++k) s += (double)__bfloat162float(W[f * 128 + k]) * __bfloat162float(X[v * 128 + k]);
            Z[f * 96 + v] = (float)s;
            e1 = fmax(e1, fabs(s - h1[f * 96 + v])); m1 = fmax(m1, fabs(s));
        }
    for (int f = 0; f < 128; ++f)
        for (int n = 0; n < 81; ++n) {
            double s = 0;
            for (int k = 0; k < 96; ++k) s += (double)tf(A[n * 96 + k]) * tf(h1[f * 96 + k]);
            e2 = fmax(e2, fabs(s - h2[f * 96 + n])); m2 = fmax(m2, fabs(s));
        }
    printf("MMA1 (TS, MN-major SW64 B): max err %.3e (max |ref| %.3f)\n", e1, m1);
    printf("MMA2 (SS, K-major SW64 A, banded hi/lo B, bias step): max err %.3e (max |ref| %.3f)\n", e2, m2);
    printf("sample out1[0][0..3] = %f %f %f %f ; ref %f %f %f %f\n", h1[0], h1[1], h1[2], h1[3], Z[0], Z[1], Z[2], Z[3]);
    return (e1 < 1e-3 && e2 < 2e-3) ? 0 : 2;
}
