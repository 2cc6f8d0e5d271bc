// Tensor-core backward of the GCN trunk (bf16 operands, fp32 accumulate), the counterpart of
// gcn_forward_tc_kernel<.., kSave = true>.  Same trick as the forward: every product is arranged so that
// the accumulator has one FEATURE per TMEM lane and the board's nodes along the columns, which makes the
// ReLU mask, the A_hat aggregation of the gradient (A_hat is symmetric: the transposed-CSR scatter is the
// same 5-point gather), the bias gradients and the operand conversions thread-local.
//
// One CTA (128 threads = 128 TMEM lanes), persistent over boards.  Per board and layer l = 3, 2:
//     dY_l = (X_l > 0) * dX_l                       thread n holds dY_l[:, n]   (dX_3 = dg / 81)
//     dZ_l = A_hat dY_l                             unrolled stencil in registers
//     dX_{l-1}^T = W_l^T dZ_l^T   MMA  A = W_l^T [k][n],  B = dZ_l node-major [v][n]      -> TMEM [k lanes][v]
//     dW_l      += dZ_l^T X_{l-1} MMA  A = dZ_l feature-major [n][v], B = X_{l-1} feature-major [k][v]
// and for layer 1 (forward: Y_1 = (A_hat X0 | 1) W1ext^T):  dW1ext += dY_1^T A1.
// The weight-gradient accumulators live in TMEM for the whole kernel (accumulate across boards) and are
// written once per CTA into its slot of the partial-gradient buffer; reduce_partials_kernel sums the slots
// in a fixed order (deterministic, atomic-free).
#include <cuda_bf16.h>
#include "gnn_fp32.cuh"
#include "tc_common.cuh"

using namespace aq;
using namespace aqtc;

namespace {

constexpr int kNodesPad = 96;
constexpr uint32_t kRowBlock = 128 * 128;   // K-block of a 128-row tile (64 bf16 = 128 B per row)
constexpr uint32_t kNodeBlock = 96 * 128;   // K-block of the 96-row node-major tile
constexpr uint32_t kA1Block = 16 * 128;     // K-block of the 16-row layer-1 operand tile
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColX = 0, kColW3 = 96, kColW2 = 224, kColW1 = 352;  // TMEM column map

constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct BwdSmem {
    unsigned char w2t[2 * kRowBlock];   // W2^T: row k, K = n
    unsigned char w3t[2 * kRowBlock];
    unsigned char tnm[2 * kNodeBlock];  // dZ node-major   [96 v][128 n]
    unsigned char tfm[2 * kRowBlock];   // dZ / dY1 feature-major [128 n][96 v] (K-block 1 half used)
    unsigned char xfm[2 * kRowBlock];   // X_{l-1} feature-major  [128 k][96 v]
    unsigned char a1t[2 * kA1Block];    // layer-1 node operand transposed [16][96 v]
    float4 rec[kV + 3];
    unsigned long long mbar;
    uint32_t tmem_base;
};
static_assert(sizeof(BwdSmem) + 1024 <= 227 * 1024, "BwdSmem exceeds shared memory");

// 16-byte chunk j (8 bf16) of `row` in a K-major SWIZZLE_128B tile whose K-blocks are `kblock` bytes
__device__ __forceinline__ uint32_t chunk_off(int row, int j, uint32_t kblock) {
    return (uint32_t)(j >> 3) * kblock + (uint32_t)row * 128u + (uint32_t)(((j & 7) ^ (row & 7)) << 4);
}
// byte offset of K-step s (16 bf16 = 32 B) inside such a tile
__device__ __forceinline__ uint32_t kstep_off(int s, uint32_t kblock) { return (uint32_t)(s >> 2) * kblock + (uint32_t)(s & 3) * 32u; }

// bit i of the result = element i of the 8 packed bf16 is > 0 (post-ReLU values are >= 0)
__device__ __forceinline__ uint32_t positive_bits(uint4 c) {
    const uint32_t w[4] = {c.x, c.y, c.z, c.w};
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m |= ((w[i] & 0x7FFFu) != 0u ? 1u : 0u) << (2 * i);
        m |= ((w[i] & 0x7FFF0000u) != 0u ? 1u : 0u) << (2 * i + 1);
    }
    return m;
}

// dZ = A_hat dY for this thread's feature, then both operand layouts:
//   node-major (row v, this feature's column): 16-bit stores through the per-phase addresses xs[]
//   feature-major (this feature's row, 96 columns): 16-byte chunk stores
#define AQ_Y(c) ((c) < 32 ? ya[(c)] : ((c) < 64 ? yb[(c) - 32] : yc[(c) - 64]))

__device__ __forceinline__ void stencil_and_store(const float *ya, const float *yb, const float *yc, const float4 *rec,
                                                  const uint32_t *xs, unsigned char *tfm_row_base, int row) {
    float hold[8];
    float4 rn = rec[0];
#pragma unroll
    for (int v = 0; v < kV; ++v) {
        const float4 r0 = rn;
        if (v + 1 < kV) rn = rec[v + 1];
        const float cr = rn.w;
        float s = r0.x * AQ_Y(v);
        if (v >= 9) s = fmaf(r0.y, AQ_Y(v - 9), s);
        if (v < kV - 9) s = fmaf(r0.z, AQ_Y(v + 9), s);
        if (v % 9 != 0) s = fmaf(r0.w, AQ_Y(v - 1), s);
        if (v % 9 != 8) s = fmaf(cr, AQ_Y(v + 1), s);
        asm volatile("{\n\t.reg .b16 t;\n\tcvt.rn.bf16.f32 t, %1;\n\tst.shared.b16 [%0], t;\n\t}\n" ::"r"(xs[v & 7] + v * 128), "f"(s));
        hold[v & 7] = s;
        if ((v & 7) == 7) *reinterpret_cast<uint4 *>(tfm_row_base + chunk_off(row, v >> 3, kRowBlock)) = pack8_bf16(hold);
    }
    const float t[8] = {hold[0], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    *reinterpret_cast<uint4 *>(tfm_row_base + chunk_off(row, 10, kRowBlock)) = pack8_bf16(t);
    *reinterpret_cast<uint4 *>(tfm_row_base + chunk_off(row, 11, kRowBlock)) = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void __launch_bounds__(kGroupThreads, 1)
gcn_backward_tc_kernel(const float *__restrict__ params, float *__restrict__ saved, const float *__restrict__ dg,
                       int64_t B, float *__restrict__ partial) {
    extern __shared__ unsigned char smem_raw[];
    BwdSmem &sm = *reinterpret_cast<BwdSmem *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x;  // = feature = TMEM lane
    float *slot = partial + (int64_t)blockIdx.x * kNumParams;

    if ((int64_t)blockIdx.x >= B) {  // no board for this CTA: its slot contributes zeros to the GCN ranges
        for (int i = tid; i < kOffWP0; i += kGroupThreads) slot[i] = 0.f;
        return;
    }
    // W_l^T tiles: element (row k, col n) = W_l[n][k]
    for (int i = tid; i < 2 * kH * kH; i += kGroupThreads) {
        const int which = i >> 14, e = i & 16383;
        const int n = e >> 7, k = e & 127;  // coalesced along k
        const float w = __ldg(params + (which ? kOffW3 : kOffW2) + e);
        unsigned char *tile = which ? sm.w3t : sm.w2t;
        *reinterpret_cast<unsigned short *>(tile + chunk_off(k, n >> 3, kRowBlock) + (n & 7) * 2) = bf16_bits(w);
    }
    const uint32_t bar = smem_u32(&sm.mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    const uint32_t tmem = sm.tmem_base;
    const uint32_t lane_base = tmem + ((uint32_t)((tid >> 5) * 32) << 16);
    const uint32_t w2t_addr = smem_u32(sm.w2t), w3t_addr = smem_u32(sm.w3t), tnm_addr = smem_u32(sm.tnm);
    const uint32_t tfm_addr = smem_u32(sm.tfm), xfm_addr = smem_u32(sm.xfm), a1t_addr = smem_u32(sm.a1t);
    uint32_t xs[8];  // node-major store addresses of this feature's column, one per (row & 7) swizzle phase
    {
        const int j = tid >> 3;
#pragma unroll
        for (int t = 0; t < 8; ++t) xs[t] = tnm_addr + (uint32_t)(j >> 3) * kNodeBlock + (uint32_t)(((j & 7) ^ t) << 4) + (tid & 7) * 2;
    }
    const SavedLayout L{B};
    float db1 = 0.f, db2 = 0.f, db3 = 0.f;
    uint32_t phase = 0;
    bool first = true;

    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        // ---- coefficients, the ReLU mask of layer 3, and X2 (feature-major) into its operand tile -------------
        if (tid < kV) {
            const float *rs = saved + L.coef() + (b * kV + tid) * 4;
            sm.rec[tid] = make_float4(rs[0], rs[1], rs[2], rs[3]);
        }
        uint32_t m[3] = {0u, 0u, 0u};
        {
            const uint4 *x3 = reinterpret_cast<const uint4 *>(tc_xfm(saved, B, 2) + (b * kH + tid) * kNodesPad);
            const uint4 *x2 = reinterpret_cast<const uint4 *>(tc_xfm(saved, B, 1) + (b * kH + tid) * kNodesPad);
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                m[j >> 2] |= positive_bits(x3[j]) << (8 * (j & 3));
                *reinterpret_cast<uint4 *>(sm.xfm + chunk_off(tid, j, kRowBlock)) = x2[j];
            }
        }
        const float dgn = dg[b * kH + tid] / (float)kV;  // d mean / d x_v
        __syncthreads();
#pragma unroll 1
        for (int layer = 2; layer >= 1; --layer) {
            // ---- dY of layer (layer+1): masked dX, this thread's feature for all nodes --------------------------
            float ya[32], yb[32], yc[32];
            if (layer == 2) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    ya[i] = (m[0] >> i) & 1 ? dgn : 0.f;
                    yb[i] = (m[1] >> i) & 1 ? dgn : 0.f;
                    yc[i] = (m[2] >> i) & 1 ? dgn : 0.f;
                }
            } else {
                tmem_ld32(lane_base + kColX, ya);
                tmem_ld32(lane_base + kColX + 32, yb);
                tmem_ld32(lane_base + kColX + 64, yc);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    ya[i] = (m[0] >> i) & 1 ? ya[i] : 0.f;
                    yb[i] = (m[1] >> i) & 1 ? yb[i] : 0.f;
                    yc[i] = (m[2] >> i) & 1 ? yc[i] : 0.f;
                }
            }
            float bsum = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) bsum += ya[i] + yb[i] + yc[i];  // columns >= 81 are masked to 0 (saved zeros)
            if (layer == 2) db3 += bsum; else db2 += bsum;
            // ---- dZ = A_hat dY in both operand layouts ------------------------------------------------------------
            stencil_and_store(ya, yb, yc, sm.rec, xs, sm.tfm, tid);
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t wt = layer == 2 ? w3t_addr : w2t_addr;
#pragma unroll
                for (int k = 0; k < 8; ++k)  // dX^T = W^T dZ^T : M = 128 (k_in), N = 96 (nodes), K = 128 (n_out)
                    mma_bf16(tmem + kColX, desc_sw128(wt + kstep_off(k, kRowBlock)), desc_sw128(tnm_addr + kstep_off(k, kNodeBlock)),
                             idesc_bf16(128, kNodesPad), k > 0 ? 1u : 0u);
                const uint32_t accw = tmem + (layer == 2 ? kColW3 : kColW2);
#pragma unroll
                for (int k = 0; k < 6; ++k)  // dW += dZ^T X : M = 128 (n_out), N = 128 (k_in), K = 96 (nodes)
                    mma_bf16(accw, desc_sw128(tfm_addr + kstep_off(k, kRowBlock)), desc_sw128(xfm_addr + kstep_off(k, kRowBlock)),
                             idesc_bf16(128, 128), (first && k == 0) ? 0u : 1u);
                mma_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            // ---- mask of the layer below from its activations (this thread's row of the operand tile), then
            //      refill the tile with the next layer's operand (X1 for layer 1; nothing after that) --------------
            m[0] = m[1] = m[2] = 0u;
#pragma unroll
            for (int j = 0; j < 12; ++j) m[j >> 2] |= positive_bits(*reinterpret_cast<const uint4 *>(sm.xfm + chunk_off(tid, j, kRowBlock))) << (8 * (j & 3));
            if (layer == 2) {
                const uint4 *x1 = reinterpret_cast<const uint4 *>(tc_xfm(saved, B, 0) + (b * kH + tid) * kNodesPad);
#pragma unroll
                for (int j = 0; j < 12; ++j) *reinterpret_cast<uint4 *>(sm.xfm + chunk_off(tid, j, kRowBlock)) = x1[j];
            }
        }
        // ---- layer 1: dY1 = (X1 > 0) * dX1 ; dW1ext += dY1^T A1 --------------------------------------------------
        {
            float ya[32], yb[32], yc[32];
            tmem_ld32(lane_base + kColX, ya);
            tmem_ld32(lane_base + kColX + 32, yb);
            tmem_ld32(lane_base + kColX + 64, yc);
            float bsum = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                ya[i] = (m[0] >> i) & 1 ? ya[i] : 0.f;
                yb[i] = (m[1] >> i) & 1 ? yb[i] : 0.f;
                yc[i] = (m[2] >> i) & 1 ? yc[i] : 0.f;
                bsum += ya[i] + yb[i] + yc[i];
            }
            db1 += bsum;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                *reinterpret_cast<uint4 *>(sm.tfm + chunk_off(tid, j, kRowBlock)) = pack8_bf16(ya + 8 * j);
                *reinterpret_cast<uint4 *>(sm.tfm + chunk_off(tid, 4 + j, kRowBlock)) = pack8_bf16(yb + 8 * j);
                *reinterpret_cast<uint4 *>(sm.tfm + chunk_off(tid, 8 + j, kRowBlock)) = pack8_bf16(yc + 8 * j);
            }
            // A1^T [16][96] of this board: 16 rows x 12 chunks
            const uint4 *a1 = reinterpret_cast<const uint4 *>(tc_a1t(saved, B) + b * 16 * kNodesPad);
            for (int c = tid; c < 16 * 12; c += kGroupThreads) {
                const int r = c / 12, j = c % 12;
                *reinterpret_cast<uint4 *>(sm.a1t + chunk_off(r, j, kA1Block)) = a1[c];
            }
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
                for (int k = 0; k < 6; ++k)  // dW1ext += dY1^T A1 : M = 128, N = 16, K = 96
                    mma_bf16(tmem + kColW1, desc_sw128(tfm_addr + kstep_off(k, kRowBlock)), desc_sw128(a1t_addr + kstep_off(k, kA1Block)),
                             idesc_bf16(128, 16), (first && k == 0) ? 0u : 1u);
                mma_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        }
        first = false;
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();  // next board overwrites rec and the operand tiles
    }
    // ---- this CTA's partial gradients: accumulator rows -> its slot ---------------------------------------------
    {
        float v[32];
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
            tmem_ld32(lane_base + kColW2 + cb * 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) slot[kOffW2 + tid * kH + cb * 32 + i] = v[i];
            tmem_ld32(lane_base + kColW3 + cb * 32, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) slot[kOffW3 + tid * kH + cb * 32 + i] = v[i];
        }
        tmem_ld32(lane_base + kColW1, v);  // 16 columns used: [hi part (6) | lo part (6) | bias_hi | bias_lo | 0 | 0]
#pragma unroll
        for (int f = 0; f < kF; ++f) slot[kOffW1 + tid * kF + f] = v[f] + v[kF + f];  // both halves multiply W1
        slot[kOffB1 + tid] = db1;
        slot[kOffB2 + tid] = db2;
        slot[kOffB3 + tid] = db3;
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace

// partial: [kSlots = 148][64082] floats; this kernel fills the GCN ranges (W1,B1,W2,B2,W3,B3) of every slot
int aq_gcn_backward_tc(const float *params, float *saved, const float *dg, int64_t B, float *partial, cudaStream_t st) {
    const size_t smem = sizeof(BwdSmem) + 1024;
    cudaError_t e = cudaFuncSetAttribute(gcn_backward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return aq_set_error((int)e, "gcn_backward_tc smem");
    gcn_backward_tc_kernel<<<148, kGroupThreads, smem, st>>>(params, saved, dg, B, partial);
    return aq_check_launch("gcn_backward_tc_kernel");
}
