// bf16 tensor-core path of the GCN trunk (inference): graph build + 3 GCN layers + mean pool with
// all three node transforms on the 5th-gen tensor cores.
//
// Layout idea: compute the TRANSPOSED product  Z^T = W X^T  so that the accumulator puts one
// FEATURE on each TMEM lane and the board's 81 NODES along the TMEM columns:
//     A operand = W   [M = 128 out-features][K]   (bf16, K-major, resident in shared memory)
//     B operand = X   [N = 96 >= 81 nodes  ][K]   (bf16, K-major SWIZZLE_128B, rebuilt per layer)
//     D         = Z^T [128 lanes = features][96 columns = nodes]  fp32 in TMEM
// A thread then owns one feature of every node of its board: the A_hat aggregation (a 5-point
// stencil over the 9x9 board with coefficients dinv_i*dinv_j), bias, ReLU, the bf16 conversion for
// the next layer and the final mean pool are all thread-local, straight out of tcgen05.ld
// registers -- no shared-memory staging of Z, no cross-thread exchange, no reduction.
//
//   * one persistent CTA per SM; kGroups independent 4-warp groups (128 threads = 128 TMEM lanes),
//     each owning one board at a time (own X tile, 96 TMEM columns, mbarrier, named barrier);
//     all groups share the bf16 weight tiles.  While one group waits on its MMAs the others run.
//   * layer 1 (K = 6) is one K = 16 MMA: the node operand carries
//     [hi(A_hat x0) | lo(A_hat x0) | 1 | 1 | 0 | 0] and the weight operand [W1 | W1 | b1_hi | b1_lo | 0 | 0],
//     so the input keeps ~16 mantissa bits and the bias is folded in.
//   * rows 81..95 of the X tile are never written: accumulator column n depends only on X row n, and
//     columns >= 81 are never read (the stencil is unrolled, out-of-board neighbours do not exist in
//     the code).
#include <cstddef>
#include <cstdlib>
#include <cuda_bf16.h>
#include "gnn_fp32.cuh"
#include "tc_common.cuh"

using namespace aq;
using namespace aqtc;

namespace {

constexpr int kNodesPad = 96;                              // MMA N: 81 nodes padded to a multiple of 16
constexpr uint32_t kWKBlock = 128 * 128;                   // weight tile: 128 rows x 128 B per K-block
constexpr uint32_t kXKBlock = kNodesPad * 128;             // node tile: 96 rows x 128 B per K-block
constexpr uint32_t kTmemCols = 512;

// tcgen05 instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1),
// both K-major (bits 15,16 = 0), N>>3 at bits 17-22, M>>4 at bits 24-28.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kNodesPad >> 3) << 17) | ((128u >> 4) << 24);

struct TcGroupSmem {
    unsigned char x[2 * kXKBlock];   // 24 KB, 1024-byte aligned (SWIZZLE_128B atoms are 8 rows x 128 B)
    float4 rec[kV + 3];              // per node {c0,cu,cd,cl}: A_hat coefficients (0 = closed); c_right(v) = c_left(v+1)
    float x0[kV * kF + 2];
    uint8_t open_s[96];
    unsigned char pad[1024 - ((kV + 3) * 16 + (kV * kF + 2) * 4 + 96) % 1024];
};
static_assert(sizeof(TcGroupSmem) % 1024 == 0, "group smem must keep 1024-byte alignment");

template <int kGroups>
struct TcSmem {
    unsigned char w2[2 * kWKBlock];
    unsigned char w3[2 * kWKBlock];
    unsigned char w1[128 * 32];      // layer-1 weight operand: bf16 [128 m][16 k], K-major SWIZZLE_32B
    TcGroupSmem g[kGroups];
    unsigned long long mbar[kGroups];
    uint32_t tmem_base;
};
static_assert(sizeof(TcSmem<5>) + 1024 <= 227 * 1024, "TcSmem exceeds shared memory");
static_assert(offsetof(TcSmem<5>, w3) == kPrepW3 && offsetof(TcSmem<5>, w1) == kPrepW1, "prepared layout must match the shared-memory layout");
static_assert(5 * kNodesPad <= (int)kTmemCols, "TMEM columns");

// Accumulator column c of the current layer: columns 0..31 / 64..95 live in za (reloaded once), 32..63 in zb.
#define AQ_Z(c) ((c) < 32 ? za[(c)] : ((c) < 64 ? zb[(c) - 32] : za[(c) - 64]))

// kSave (training forward): additionally writes what the tensor-core backward needs into the SavedLayout
// regions (gnn_layout.cuh) -- the post-ReLU activations of the three layers FEATURE-major as bf16
// [B][128][96] (zero beyond node 80) inside the x(l) regions, the layer-1 node operand transposed
// [B][16][96] behind them, and the A_hat coefficients [B][81][4] in the coef region.
template <int kGroups, bool kSave>
__global__ void __launch_bounds__(kGroups * kGroupThreads, 1)
gcn_forward_tc_kernel(const float *__restrict__ params, const unsigned char *__restrict__ prepared,
                      const AqState *__restrict__ states, int64_t B, float *__restrict__ pooled_out,
                      float *__restrict__ saved) {
    constexpr int kTcThreads = kGroups * kGroupThreads;
    extern __shared__ unsigned char smem_raw[];
    TcSmem<kGroups> &sm = *reinterpret_cast<TcSmem<kGroups> *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int gtid = threadIdx.x;
    const int grp = gtid / kGroupThreads, tid = gtid % kGroupThreads;  // tid = feature = TMEM lane

    if (prepared) {  // operand tiles were built once by aq_prepare_inference: w2 | w3 | w1 are contiguous in both places
        const uint4 *src = reinterpret_cast<const uint4 *>(prepared);
        uint4 *dst = reinterpret_cast<uint4 *>(sm.w2);
        for (int c = gtid; c < (int)(kPrepHeadB1 / 16); c += kTcThreads) dst[c] = __ldg(src + c);
    } else {
        for (int c = gtid; c < 2 * 128 * 16; c += kTcThreads) {  // both 128x128 weight tiles, 16-byte chunks
            const int which = c >> 11, cc = c & 2047;
            const int n = cc >> 4, j = cc & 15;
            const float *W = params + (which ? kOffW3 : kOffW2);
            const float4 lo = __ldg(reinterpret_cast<const float4 *>(W + n * kH + j * 8));
            const float4 hi = __ldg(reinterpret_cast<const float4 *>(W + n * kH + j * 8) + 1);
            uint4 v;
            v.x = pack_bf16(lo.x, lo.y); v.y = pack_bf16(lo.z, lo.w);
            v.z = pack_bf16(hi.x, hi.y); v.w = pack_bf16(hi.z, hi.w);
            *reinterpret_cast<uint4 *>((which ? sm.w3 : sm.w2) + sw128_chunk(n, j, kWKBlock)) = v;
        }
        if (gtid < kH) {  // layer-1 weight operand: [W1 (6) | W1 (6) | b1_hi | b1_lo | 0 | 0]
            const int n = gtid;
            float w[kF];
    #pragma unroll
            for (int f = 0; f < kF; ++f) w[f] = __ldg(params + kOffW1 + n * kF + f);
            const float bias = __ldg(params + kOffB1 + n);
            const float bias_hi = __bfloat162float(__float2bfloat16_rn(bias));
            uint4 c0, c1;
            c0.x = pack_bf16(w[0], w[1]); c0.y = pack_bf16(w[2], w[3]); c0.z = pack_bf16(w[4], w[5]); c0.w = pack_bf16(w[0], w[1]);
            c1.x = pack_bf16(w[2], w[3]); c1.y = pack_bf16(w[4], w[5]); c1.z = pack_bf16(bias_hi, bias - bias_hi); c1.w = 0u;
            *reinterpret_cast<uint4 *>(sm.w1 + sw32_chunk(n, 0)) = c0;
            *reinterpret_cast<uint4 *>(sm.w1 + sw32_chunk(n, 1)) = c1;
        }
    }
    if (gtid < kGroups) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&sm.mbar[gtid])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (gtid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    TcGroupSmem &gs = sm.g[grp];
    const uint32_t tmem_grp = sm.tmem_base + (uint32_t)(grp * kNodesPad);                    // group's columns
    const uint32_t tmem_me = tmem_grp + ((uint32_t)((tid >> 5) * 32) << 16);                  // + this warp's lane quadrant
    const uint32_t bar = smem_u32(&sm.mbar[grp]);
    const uint32_t x_addr = smem_u32(gs.x), w1_addr = smem_u32(sm.w1), w2_addr = smem_u32(sm.w2), w3_addr = smem_u32(sm.w3);
    const float bias2 = __ldg(params + kOffB2 + tid), bias3 = __ldg(params + kOffB3 + tid);
    // store addresses of this thread's feature column inside the node tile, one per (row & 7) swizzle phase
    uint32_t xs[8];
    {
        const int j = tid >> 3;
#pragma unroll
        for (int t = 0; t < 8; ++t) xs[t] = x_addr + (uint32_t)(j >> 3) * kXKBlock + (uint32_t)(((j & 7) ^ t) << 4) + (tid & 7) * 2;
    }
    uint32_t phase = 0;
    // MMAs are issued by the first warp of a group from warp-uniform values (descriptors in uniform registers; issuing from
    // `if (tid == 0)` costs a register -> uniform-register broadcast loop of ~80 cycles per MMA)
    const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int grp_u = warp_u >> 2;
    const bool issuer_warp = (warp_u & 3) == 0;
    const uint32_t tmem_grp_u = __shfl_sync(0xffffffffu, sm.tmem_base, 0) + (uint32_t)(grp_u * kNodesPad);
    const uint32_t smem_u = smem_u32(&sm);
    const uint32_t x_u = smem_u + (uint32_t)offsetof(TcSmem<kGroups>, g) + (uint32_t)grp_u * (uint32_t)sizeof(TcGroupSmem);
    const uint32_t w1_u = smem_u + (uint32_t)offsetof(TcSmem<kGroups>, w1), w2_u = smem_u + (uint32_t)offsetof(TcSmem<kGroups>, w2);
    const uint32_t w3_u = smem_u + (uint32_t)offsetof(TcSmem<kGroups>, w3);
    const uint32_t bar_u = smem_u + (uint32_t)offsetof(TcSmem<kGroups>, mbar) + (uint32_t)grp_u * 8u;

    for (int64_t b = (int64_t)blockIdx.x * kGroups + grp; b < B; b += (int64_t)gridDim.x * kGroups) {
        // ---- inputs: node features + open-direction masks ------------------------------------------
        {
            const AqState s = load_state(states + b);
            board_inputs_from_state(s, gs.x0, gs.open_s, tid);
        }
        group_sync(grp);
        // ---- per node: A_hat coefficients, and the layer-1 node operand row
        //      [hi(A_hat x0) (6) | lo(A_hat x0) (6) | 1 | 1 | 0 | 0] as bf16 --------------------------------
        if (tid < kV) {
            const int v = tid, m = gs.open_s[v];
            const float dv = dinv_of(m);
            const int iu = (m & 1) ? v - 9 : v, id = (m & 2) ? v + 9 : v, il = (m & 4) ? v - 1 : v, ir = (m & 8) ? v + 1 : v;
            const float c0 = dv * dv;
            const float cu = (m & 1) ? dv * dinv_of(gs.open_s[iu]) : 0.f, cd = (m & 2) ? dv * dinv_of(gs.open_s[id]) : 0.f;
            const float cl = (m & 4) ? dv * dinv_of(gs.open_s[il]) : 0.f, cr = (m & 8) ? dv * dinv_of(gs.open_s[ir]) : 0.f;
            gs.rec[v] = make_float4(c0, cu, cd, cl);
            unsigned short hi[kF], lo[kF];
#pragma unroll
            for (int f = 0; f < kF; ++f) {
                float s = c0 * gs.x0[v * kF + f];
                s = fmaf(cu, gs.x0[iu * kF + f], s);
                s = fmaf(cd, gs.x0[id * kF + f], s);
                s = fmaf(cl, gs.x0[il * kF + f], s);
                s = fmaf(cr, gs.x0[ir * kF + f], s);
                const float h = __bfloat162float(__float2bfloat16_rn(s));
                hi[f] = bf16_bits(s);
                lo[f] = bf16_bits(s - h);
            }
            uint4 c0v, c1v;
            c0v.x = hi[0] | ((uint32_t)hi[1] << 16); c0v.y = hi[2] | ((uint32_t)hi[3] << 16);
            c0v.z = hi[4] | ((uint32_t)hi[5] << 16); c0v.w = lo[0] | ((uint32_t)lo[1] << 16);
            c1v.x = lo[2] | ((uint32_t)lo[3] << 16); c1v.y = lo[4] | ((uint32_t)lo[5] << 16);
            c1v.z = 0x3F803F80u; c1v.w = 0u;  // 1, 1, 0, 0
            *reinterpret_cast<uint4 *>(gs.x + sw128_chunk(v, 0, kXKBlock)) = c0v;
            *reinterpret_cast<uint4 *>(gs.x + sw128_chunk(v, 1, kXKBlock)) = c1v;
            if (kSave) {
                const SavedLayout L{B};
                float *rs = saved + L.coef() + (b * kV + v) * 4;
                rs[0] = c0; rs[1] = cu; rs[2] = cd; rs[3] = cl;
                unsigned short *at = tc_a1t(saved, B) + b * 16 * kNodesPad + v;
                const uint32_t wds[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
#pragma unroll
                for (int k = 0; k < 16; ++k) at[k * kNodesPad] = (unsigned short)(wds[k >> 1] >> (16 * (k & 1)));
            }
        } else if (kSave && tid < kNodesPad) {  // K padding of the transposed layer-1 operand must be zero
            unsigned short *at = tc_a1t(saved, B) + b * 16 * kNodesPad + tid;
#pragma unroll
            for (int k = 0; k < 16; ++k) at[k * kNodesPad] = 0;
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        group_sync(grp);
        if (issuer_warp) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (elect_one_lane()) {
                mma_bf16(tmem_grp_u, desc_sw32(w1_u), desc_sw128(x_u), kIdesc, 0u);  // one K = 16 step
                mma_commit(bar_u);
            }
            __syncwarp();
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        // ---- layer 1 epilogue: ReLU -> bf16 -> this thread's feature column of the node tile ---------------
        {
            float za[32];
#pragma unroll
            for (int cb = 0; cb < 3; ++cb) {
                tmem_ld32(tmem_me + cb * 32, za);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int v = cb * 32 + i;
                    if (v < kV) st_relu_bf16(xs[v & 7] + v * 128, za[i]);
                }
                if (kSave) {
                    uint4 *row = reinterpret_cast<uint4 *>(tc_xfm(saved, B, 0) + (b * kH + tid) * kNodesPad) + cb * 4;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float h[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) h[e] = (cb * 32 + i * 8 + e < kV) ? fmaxf(za[i * 8 + e], 0.f) : 0.f;
                        row[i] = pack8_bf16(h);
                    }
                }
            }
        }
        float pool = 0.f;
#pragma unroll 1
        for (int layer = 1; layer < kLayers; ++layer) {
            // make the generic-proxy writes of the node tile visible to the tensor core (async proxy)
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            group_sync(grp);
            if (issuer_warp) {
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                if (elect_one_lane()) {
                    const uint32_t w_addr = layer == 1 ? w2_u : w3_u;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {  // K = 128 = 8 x UMMA_K(16); 4 steps of 32 B per 128 B swizzle span
                        const uint32_t woff = (uint32_t)(k >> 2) * kWKBlock + (uint32_t)(k & 3) * 32u;
                        const uint32_t xoff = (uint32_t)(k >> 2) * kXKBlock + (uint32_t)(k & 3) * 32u;
                        mma_bf16(tmem_grp_u, desc_sw128(w_addr + woff), desc_sw128(x_u + xoff), kIdesc, k > 0 ? 1u : 0u);
                    }
                    mma_commit(bar_u);
                }
                __syncwarp();
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            // ---- aggregation + bias + ReLU, thread-local: this thread holds feature `tid` of all 81 nodes.
            //      The stencil is fully unrolled; neighbours that fall off the board are absent from the code.
            if (layer + 1 < kLayers) {
                float za[32], zb[32];
                tmem_ld32(tmem_me, za);
                tmem_ld32(tmem_me + 32, zb);
                float4 rn = gs.rec[0];
                float hold[8];
                uint4 *srow = kSave ? reinterpret_cast<uint4 *>(tc_xfm(saved, B, 1) + (b * kH + tid) * kNodesPad) : nullptr;
#pragma unroll
                for (int v = 0; v < kV; ++v) {
                    if (v == 41) tmem_ld32(tmem_me + 64, za);  // columns 0..31 are dead after node 40
                    const float4 r0 = rn;
                    if (v + 1 < kV) rn = gs.rec[v + 1];
                    const float cr = rn.w;  // c_right(v) = c_left(v + 1): A_hat is symmetric
                    float s = fmaf(r0.x, AQ_Z(v), bias2);
                    if (v >= 9) s = fmaf(r0.y, AQ_Z(v - 9), s);
                    if (v < kV - 9) s = fmaf(r0.z, AQ_Z(v + 9), s);
                    if (v % 9 != 0) s = fmaf(r0.w, AQ_Z(v - 1), s);
                    if (v % 9 != 8) s = fmaf(cr, AQ_Z(v + 1), s);
                    st_relu_bf16(xs[v & 7] + v * 128, s);  // next layer's node operand
                    if (kSave) {
                        hold[v & 7] = fmaxf(s, 0.f);
                        if ((v & 7) == 7) srow[v >> 3] = pack8_bf16(hold);
                    }
                }
                if (kSave) {  // node 80 and the zero K padding (columns 81..95)
                    const float t[8] = {hold[0], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    srow[10] = pack8_bf16(t);
                    srow[11] = make_uint4(0u, 0u, 0u, 0u);
                }
            } else {
                float za[32], zb[32];
                tmem_ld32(tmem_me, za);
                tmem_ld32(tmem_me + 32, zb);
                float4 rn = gs.rec[0];
                float hold[8];
                uint4 *srow = kSave ? reinterpret_cast<uint4 *>(tc_xfm(saved, B, 2) + (b * kH + tid) * kNodesPad) : nullptr;
#pragma unroll
                for (int v = 0; v < kV; ++v) {
                    if (v == 41) tmem_ld32(tmem_me + 64, za);
                    const float4 r0 = rn;
                    if (v + 1 < kV) rn = gs.rec[v + 1];
                    const float cr = rn.w;  // c_right(v) = c_left(v + 1): A_hat is symmetric
                    float s = fmaf(r0.x, AQ_Z(v), bias3);
                    if (v >= 9) s = fmaf(r0.y, AQ_Z(v - 9), s);
                    if (v < kV - 9) s = fmaf(r0.z, AQ_Z(v + 9), s);
                    if (v % 9 != 0) s = fmaf(r0.w, AQ_Z(v - 1), s);
                    if (v % 9 != 8) s = fmaf(cr, AQ_Z(v + 1), s);
                    pool += fmaxf(s, 0.f);  // last layer feeds only the mean pool
                    if (kSave) {
                        hold[v & 7] = fmaxf(s, 0.f);
                        if ((v & 7) == 7) srow[v >> 3] = pack8_bf16(hold);
                    }
                }
                if (kSave) {
                    const float t[8] = {hold[0], 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    srow[10] = pack8_bf16(t);
                    srow[11] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
        }
        // ---- global_mean_pool: thread-local ------------------------------------------------------------------
        pooled_out[b * kH + tid] = pool / (float)kV;
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        group_sync(grp);  // the next board overwrites x0 / open_s / rec / the node tile
    }
    // ---- teardown ------------------------------------------------------------------------------------
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (gtid < 32) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(sm.tmem_base), "r"(kTmemCols) : "memory");
    }
}

}  // namespace

template <int kGroups>
static int launch_tc(const float *params, const unsigned char *prepared, const AqState *states, int64_t B, float *pooled,
                     float *saved, int sms, cudaStream_t st) {
    const size_t smem = sizeof(TcSmem<kGroups>) + 1024;
    const int64_t want = (B + kGroups - 1) / kGroups;
    const unsigned grid = (unsigned)(want < sms ? want : sms);
    cudaError_t e;
    if (saved) {
        e = cudaFuncSetAttribute(gcn_forward_tc_kernel<kGroups, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return aq_set_error((int)e, "gcn_forward_tc smem");
        gcn_forward_tc_kernel<kGroups, true><<<grid, kGroups * kGroupThreads, smem, st>>>(params, prepared, states, B, pooled, saved);
    } else {
        e = cudaFuncSetAttribute(gcn_forward_tc_kernel<kGroups, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return aq_set_error((int)e, "gcn_forward_tc smem");
        gcn_forward_tc_kernel<kGroups, false><<<grid, kGroups * kGroupThreads, smem, st>>>(params, prepared, states, B, pooled, nullptr);
    }
    return aq_check_launch("gcn_forward_tc_kernel");
}

int aq_gcn_forward_tc2(const float *params, const void *prepared, const AqState *states, int64_t B, float *pooled, float *saved,
                       cudaStream_t st, bool after_legal);  // gnn_tc2.cu
int aq_gcn_forward_tc3(const float *params, const void *prepared, const AqState *states, int64_t B, float *pooled, cudaStream_t st);  // gnn_tc3.cu

// Which tensor-core training pair (forward with saved activations + backward) is used at precision 1; both sides read it here.
int aq_train_tc_version() {
    static int v = 0;
    if (v == 0) {
        const char *env = getenv("AQ_TRAIN_TC_VERSION");  // 1 = gnn_tc.cu + gnn_tc_bwd.cu, 2 = gnn_tc2.cu + gnn_tc2_bwd.cu
        v = env ? atoi(env) : 2;
        if (v != 1) v = 2;
    }
    return v;
}

// saved == nullptr: inference.  saved != nullptr: training forward (activations kept for aq_gnn_backward, precision 1).
int aq_gcn_forward_tc(const float *params, const void *prepared_v, const AqState *states, int64_t B, float *pooled, float *saved,
                      cudaStream_t st, bool after_legal) {
    const unsigned char *prepared = reinterpret_cast<const unsigned char *>(prepared_v);
    static int sms = 0, groups = 0, version = 2;
    if (sms == 0) {
        const char *ver = getenv("AQ_TC_VERSION");  // 1 = CUDA-core stencil aggregation (this file), 2 = fp16 tensor-core aggregation (gnn_tc2.cu), 3 = tf32 aggregation from the accumulator (gnn_tc3.cu)
        if (ver) version = atoi(ver);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
        const char *env = getenv("AQ_TC_GROUPS");  // tuning knob: boards in flight per SM (3, 4 or 5)
        groups = env ? atoi(env) : 5;
    }
    if (!saved && version == 3) return aq_gcn_forward_tc3(params, prepared_v, states, B, pooled, st);
    if (!saved && version == 2) return aq_gcn_forward_tc2(params, prepared_v, states, B, pooled, nullptr, st, after_legal);
    if (saved && aq_train_tc_version() == 2) return aq_gcn_forward_tc2(params, prepared_v, states, B, pooled, saved, st, false);
    if (saved || groups == 3) return launch_tc<3>(params, prepared, states, B, pooled, saved, sms, st);  // the save variant needs the registers
    if (groups == 4) return launch_tc<4>(params, prepared, states, B, pooled, saved, sms, st);
    return launch_tc<5>(params, prepared, states, B, pooled, saved, sms, st);
}

// ---- aq_prepare_inference: fp32 parameters -> the bf16 operand tiles of the inference kernels ---------------------
namespace {
__global__ void prepare_inference_kernel(const float *__restrict__ params, unsigned char *__restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte output chunk per thread
    auto load8 = [&](const float *p, bool aligned, float *f) {
        if (aligned) {
            const float4 lo = __ldg(reinterpret_cast<const float4 *>(p)), hi = __ldg(reinterpret_cast<const float4 *>(p) + 1);
            f[0] = lo.x; f[1] = lo.y; f[2] = lo.z; f[3] = lo.w; f[4] = hi.x; f[5] = hi.y; f[6] = hi.z; f[7] = hi.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __ldg(p + e);
        }
    };
    float f[8];
    if (c < 4096) {  // trunk W2 / W3
        const int which = c >> 11, cc = c & 2047, n = cc >> 4, j = cc & 15;
        load8(params + (which ? kOffW3 : kOffW2) + n * kH + j * 8, true, f);
        *reinterpret_cast<uint4 *>(out + (which ? kPrepW3 : kPrepW2) + sw128_chunk(n, j, kWKBlock)) = pack8_bf16(f);
    } else if (c < 4096 + 128) {  // trunk layer-1 operand [W1 | W1 | b1_hi | b1_lo | 0 | 0]
        const int n = c - 4096;
        float w[kF];
#pragma unroll
        for (int k = 0; k < kF; ++k) w[k] = __ldg(params + kOffW1 + n * kF + k);
        const float bias = __ldg(params + kOffB1 + n);
        const float bias_hi = __bfloat162float(__float2bfloat16_rn(bias));
        uint4 c0, c1;
        c0.x = pack_bf16(w[0], w[1]); c0.y = pack_bf16(w[2], w[3]); c0.z = pack_bf16(w[4], w[5]); c0.w = pack_bf16(w[0], w[1]);
        c1.x = pack_bf16(w[2], w[3]); c1.y = pack_bf16(w[4], w[5]); c1.z = pack_bf16(bias_hi, bias - bias_hi); c1.w = 0u;
        *reinterpret_cast<uint4 *>(out + kPrepW1 + sw32_chunk(n, 0)) = c0;
        *reinterpret_cast<uint4 *>(out + kPrepW1 + sw32_chunk(n, 1)) = c1;
    } else if (c < 4096 + 128 + 2048) {  // heads B1: rows 0..63 = Wp0, 64..127 = Wv0 (Wv0 is not 16-byte aligned)
        const int cc = c - (4096 + 128), n = cc >> 4, j = cc & 15;
        if (n < kHH) load8(params + kOffWP0 + n * kH + j * 8, true, f);
        else load8(params + kOffWV0 + (n - kHH) * kH + j * 8, false, f);
        *reinterpret_cast<uint4 *>(out + kPrepHeadB1 + sw128_chunk(n, j, kWKBlock)) = pack8_bf16(f);
    } else if (c < 4096 + 128 + 2048 + 224 * 8) {  // heads B2 = Wp2 [209][64] padded to 224 rows
        const int cc = c - (4096 + 128 + 2048), n = cc >> 3, j = cc & 7;
        if (n < kP) load8(params + kOffWP2 + n * kHH + j * 8, true, f);
        else {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = 0.f;
        }
        *reinterpret_cast<uint4 *>(out + kPrepHeadB2 + sw128_chunk(n, j, kWKBlock)) = pack8_bf16(f);
    }
}
}  // namespace

extern "C" int64_t aq_prepared_bytes(void) { return kPrepBytes; }

extern "C" int aq_prepare_inference(const float *params, void *prepared, void *stream) {
    if (!params || !prepared) return aq_set_error(AQ_ERR_ARG, "aq_prepare_inference");
    const int chunks = 4096 + 128 + 2048 + 224 * 8;
    prepare_inference_kernel<<<(chunks + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        params, reinterpret_cast<unsigned char *>(prepared));
    return aq_check_launch("prepare_inference_kernel");
}
