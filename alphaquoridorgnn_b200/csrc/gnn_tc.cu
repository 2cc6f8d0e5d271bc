// bf16 tensor-core path of the GCN trunk (inference): graph build + 3 GCN layers + mean pool with
// the two 128x128 node transforms on the 5th-gen tensor cores.
//
//   * one CTA (256 threads) per SM, persistent over boards; W2, W3 live in shared memory as bf16 in
//     the canonical UMMA K-major SWIZZLE_128B layout for the whole kernel;
//   * per board and layer: the activations X (81 rows, padded to the M=128 tile) are written as bf16
//     into the swizzled A tile, ONE thread issues 8 x tcgen05.mma (M128 N128 K16, kind::f16, fp32
//     accumulate in TMEM), tcgen05.commit arrives on an mbarrier, the 8 warps read the accumulator
//     back with tcgen05.ld (32 lanes x 32 columns per instruction) into an fp32 Z buffer, and the
//     A_hat aggregation (+bias, ReLU) runs warp-per-node on the CUDA cores, writing the next
//     layer's A tile directly in the swizzled layout;
//   * layer 1 (K = 6) and the pooling stay on the CUDA cores in fp32.
// Rows 81..127 of the A tile are never written: every accumulator row depends only on its own A
// row, and rows >= 81 of the accumulator are never read.
#include <cuda_bf16.h>
#include "gnn_fp32.cuh"

using namespace aq;

namespace {

constexpr int kTcThreads = 256;
constexpr int kZStride = 132;                // fp32 Z rows padded: conflict-free per-row float4 stores
constexpr uint32_t kTileBytes = 128 * 256;   // 128 rows x 128 bf16 = two K-blocks of 128 rows x 128 B
constexpr uint32_t kKBlockBytes = 128 * 128;
constexpr uint32_t kTmemCols = 128;

// tcgen05 instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1),
// both K-major (bits 15,16 = 0), N>>3 at bits 17-22, M>>4 at bits 24-28.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

struct TcSmem {
    // 1024-byte aligned (SWIZZLE_128B atoms are 8 rows x 128 B)
    unsigned char w2[kTileBytes];
    unsigned char w3[kTileBytes];
    unsigned char a[kTileBytes];
    float z[kV * kZStride];
    float w1t[kF * kH];
    float b1[kH], b2[kH], b3[kH];
    float coef[kV * 5 + 3];
    float x0[kV * kF + 2];
    float ax0[kV * kF + 2];
    float red[8 * kH];
    unsigned long long mbar;
    uint32_t tmem_base;
    uint8_t open_s[96];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of the 16-byte chunk (row, chunk j of 16) inside a K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_chunk(int row, int j) {
    return (uint32_t)(j >> 3) * kKBlockBytes + (uint32_t)row * 128u + (uint32_t)(((j & 7) ^ (row & 7)) << 4);
}

// shared-memory matrix descriptor: start address>>4, LBO=1 (unused for swizzled K-major), SBO=1024 B,
// version 1 (Blackwell), layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&t);
}

// fp32 [128][128] row-major weight in global memory -> bf16 swizzled K-major tile (row = n, col = k)
__device__ __forceinline__ void load_weight_bf16(unsigned char *tile, const float *__restrict__ W, int tid) {
    for (int c = tid; c < 128 * 16; c += kTcThreads) {
        const int n = c >> 4, j = c & 15;
        const float4 lo = __ldg(reinterpret_cast<const float4 *>(W + n * kH + j * 8));
        const float4 hi = __ldg(reinterpret_cast<const float4 *>(W + n * kH + j * 8) + 1);
        uint4 v;
        v.x = pack_bf16(lo.x, lo.y); v.y = pack_bf16(lo.z, lo.w);
        v.z = pack_bf16(hi.x, hi.y); v.w = pack_bf16(hi.z, hi.w);
        *reinterpret_cast<uint4 *>(tile + sw128_chunk(n, j)) = v;
    }
}

__global__ void __launch_bounds__(kTcThreads, 1)
gcn_forward_tc_kernel(const float *__restrict__ params, const AqState *__restrict__ states, int64_t B,
                      float *__restrict__ pooled_out) {
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    TcSmem &sm = *reinterpret_cast<TcSmem *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    load_weight_bf16(sm.w2, params + kOffW2, tid);
    load_weight_bf16(sm.w3, params + kOffW3, tid);
    for (int i = tid; i < kF * kH; i += kTcThreads) {
        const int n = i / kF, f = i % kF;
        sm.w1t[f * kH + n] = __ldg(params + kOffW1 + i);
    }
    if (tid < kH) {
        sm.b1[tid] = __ldg(params + kOffB1 + tid);
        sm.b2[tid] = __ldg(params + kOffB2 + tid);
        sm.b3[tid] = __ldg(params + kOffB3 + tid);
    }
    const uint32_t bar = smem_u32(&sm.mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {  // TMEM accumulator: 128 lanes x 128 fp32 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = sm.tmem_base;
    const uint32_t a_addr = smem_u32(sm.a), w2_addr = smem_u32(sm.w2), w3_addr = smem_u32(sm.w3);
    uint32_t phase = 0;

    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        // ---- inputs ---------------------------------------------------------------------------
        {
            const AqState s = load_state(states + b);
            board_inputs_from_state(s, sm.x0, sm.open_s, tid);
        }
        __syncthreads();
        board_coefficients(sm.open_s, sm.coef, tid);
        __syncthreads();
        for (int i = tid; i < kV * kF; i += kTcThreads) {
            const int v = i / kF, f = i % kF;
            const float *c = sm.coef + v * 5;
            float s = c[0] * sm.x0[i];
            if (c[1] != 0.f) s = fmaf(c[1], sm.x0[(v - 9) * kF + f], s);
            if (c[2] != 0.f) s = fmaf(c[2], sm.x0[(v + 9) * kF + f], s);
            if (c[3] != 0.f) s = fmaf(c[3], sm.x0[(v - 1) * kF + f], s);
            if (c[4] != 0.f) s = fmaf(c[4], sm.x0[(v + 1) * kF + f], s);
            sm.ax0[i] = s;
        }
        __syncthreads();
        // ---- layer 1 on the CUDA cores, written straight into the swizzled bf16 A tile ----------
        for (int c = tid; c < kV * 16; c += kTcThreads) {
            const int v = c >> 4, j = c & 15;
            float a6[kF];
#pragma unroll
            for (int f = 0; f < kF; ++f) a6[f] = sm.ax0[v * kF + f];
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int n = j * 8 + e;
                float s = sm.b1[n];
#pragma unroll
                for (int f = 0; f < kF; ++f) s = fmaf(a6[f], sm.w1t[f * kH + n], s);
                o[e] = fmaxf(s, 0.f);
            }
            uint4 pk;
            pk.x = pack_bf16(o[0], o[1]); pk.y = pack_bf16(o[2], o[3]);
            pk.z = pack_bf16(o[4], o[5]); pk.w = pack_bf16(o[6], o[7]);
            *reinterpret_cast<uint4 *>(sm.a + sw128_chunk(v, j)) = pk;
        }
        float4 pool = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int layer = 1; layer < kLayers; ++layer) {
            // make the generic-proxy writes of the A tile visible to the tensor core (async proxy)
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t w_addr = layer == 1 ? w2_addr : w3_addr;
#pragma unroll
                for (int k = 0; k < 8; ++k) {  // K = 128 = 8 x UMMA_K(16); 4 steps of 32 B per 128 B swizzle span
                    const uint32_t off = (uint32_t)(k >> 2) * kKBlockBytes + (uint32_t)(k & 3) * 32u;
                    mma_bf16(tmem, umma_desc(a_addr + off), umma_desc(w_addr + off), kIdesc, k > 0 ? 1u : 0u);
                }
                // arrives on the mbarrier once all MMAs above have completed
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            // ---- epilogue: TMEM -> registers -> fp32 Z (row = TMEM lane) ---------------------------
            {
                const int q = warp & 3, half = warp >> 2;
                const int r = q * 32 + lane;
#pragma unroll
                for (int cb = 0; cb < 2; ++cb) {
                    const int col0 = half * 64 + cb * 32;
                    float v[32];
                    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
                    if (r < kV) {
                        float4 *dst = reinterpret_cast<float4 *>(sm.z + r * kZStride + col0);
#pragma unroll
                        for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncthreads();
            // ---- aggregation + bias + ReLU, warp per node --------------------------------------------
            {
                const float4 bb = reinterpret_cast<const float4 *>(layer == 1 ? sm.b2 : sm.b3)[lane];
                for (int v = warp; v < kV; v += kTcThreads / 32) {
                    const float *c = sm.coef + v * 5;
                    const float c0 = c[0], cu = c[1], cd = c[2], cl = c[3], cr = c[4];
                    float4 a = *reinterpret_cast<const float4 *>(sm.z + v * kZStride + lane * 4);
                    float4 s = make_float4(c0 * a.x, c0 * a.y, c0 * a.z, c0 * a.w);
                    if (cu != 0.f) { a = *reinterpret_cast<const float4 *>(sm.z + (v - 9) * kZStride + lane * 4); s.x = fmaf(cu, a.x, s.x); s.y = fmaf(cu, a.y, s.y); s.z = fmaf(cu, a.z, s.z); s.w = fmaf(cu, a.w, s.w); }
                    if (cd != 0.f) { a = *reinterpret_cast<const float4 *>(sm.z + (v + 9) * kZStride + lane * 4); s.x = fmaf(cd, a.x, s.x); s.y = fmaf(cd, a.y, s.y); s.z = fmaf(cd, a.z, s.z); s.w = fmaf(cd, a.w, s.w); }
                    if (cl != 0.f) { a = *reinterpret_cast<const float4 *>(sm.z + (v - 1) * kZStride + lane * 4); s.x = fmaf(cl, a.x, s.x); s.y = fmaf(cl, a.y, s.y); s.z = fmaf(cl, a.z, s.z); s.w = fmaf(cl, a.w, s.w); }
                    if (cr != 0.f) { a = *reinterpret_cast<const float4 *>(sm.z + (v + 1) * kZStride + lane * 4); s.x = fmaf(cr, a.x, s.x); s.y = fmaf(cr, a.y, s.y); s.z = fmaf(cr, a.z, s.z); s.w = fmaf(cr, a.w, s.w); }
                    s.x = fmaxf(s.x + bb.x, 0.f); s.y = fmaxf(s.y + bb.y, 0.f);
                    s.z = fmaxf(s.z + bb.z, 0.f); s.w = fmaxf(s.w + bb.w, 0.f);
                    if (layer + 1 < kLayers) {
                        // next layer's A tile: columns 4*lane..4*lane+3 = half of 16-byte chunk lane/2
                        uint2 pk;
                        pk.x = pack_bf16(s.x, s.y);
                        pk.y = pack_bf16(s.z, s.w);
                        *reinterpret_cast<uint2 *>(sm.a + sw128_chunk(v, lane >> 1) + (lane & 1) * 8) = pk;
                    } else {
                        pool.x += s.x; pool.y += s.y; pool.z += s.z; pool.w += s.w;
                    }
                }
            }
        }
        // ---- global_mean_pool -----------------------------------------------------------------------
        reinterpret_cast<float4 *>(sm.red + warp * kH)[lane] = pool;
        __syncthreads();
        if (tid < kH) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += sm.red[w * kH + tid];
            pooled_out[b * kH + tid] = s / (float)kV;
        }
        __syncthreads();
    }
    // ---- teardown ------------------------------------------------------------------------------------
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

}  // namespace

int aq_gcn_forward_tc(const float *params, const AqState *states, int64_t B, float *pooled, cudaStream_t st) {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const size_t smem = sizeof(TcSmem) + 1024;
    cudaError_t e = cudaFuncSetAttribute(gcn_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return aq_set_error((int)e, "gcn_forward_tc smem");
    const unsigned grid = (unsigned)(B < sms ? B : sms);
    gcn_forward_tc_kernel<<<grid, kTcThreads, smem, st>>>(params, states, B, pooled);
    return aq_check_launch("gcn_forward_tc_kernel");
}
