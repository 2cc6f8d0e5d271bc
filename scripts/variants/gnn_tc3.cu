// bf16 / tf32 tensor-core GCN trunk, version 3 (inference).  Like version 2 (gnn_tc2.cu) both the node transform and
// the A_hat aggregation run on the tensor cores, but the aggregation no longer needs its input converted or staged:
//
//   transform   Z^T = W X^T        kind::f16   A = W   [128 out][128 in]  bf16 resident in TENSOR MEMORY
//                                              B = X^T [K = 128 feat][N = 96 nodes]  bf16, MN-major SWIZZLE_64B, shared memory
//                                              D = Z^T [128 lanes = features][96 columns = nodes]  fp32 in TMEM
//   aggregate   Y^T = Z^T A_hat^T  kind::tf32  A = Z^T -- the transform's fp32 ACCUMULATOR COLUMNS read in place as a TF32 operand
//                                              B = A_hat as tf32, banded: two blocks of [48 out nodes][64 in nodes], K-major SWIZZLE_128B
//                                              D = Y^T in a second set of 96 TMEM columns
//
// The transform and the aggregation of a layer are issued back to back by one thread and complete with one commit; between two layers
// there is a single epilogue (tcgen05.ld -> + bias -> ReLU -> bf16 -> the thread's own feature row of the X^T tile).  Per board that is
// two conversion passes instead of four (fp32 -> bf16/fp16 packing runs on the XU pipe at 1/8 rate on this part and was the limiter
// of version 2), half the shared-memory traffic, and three MMA round trips instead of five.
//
// One persistent CTA per SM; 2 groups of 8 warps, one board per group in flight.  A feature (TMEM lane) is shared by two threads, one
// per column half.  TMEM: 2 x (96 + 96) accumulator columns + 2 x 64 columns holding W2 and W3 = 512.  The node phase of a group's
// next board (bitboard windows -> degrees -> tf32 adjacency tile + layer-1 operand) overlaps the last MMAs of the current one.
// Layer 1 (K = 6): fp32 aggregation of the 6-wide input by the node threads, one K = 16 MMA (hi/lo split input, bias folded).
// TF32 operands are read with their low 13 mantissa bits ignored (measured: scripts/probe/tf32_probe.cu); the A_hat entries are
// rounded to tf32 when the table is built.
#include <cstddef>
#include <cstdlib>
#include <cuda_bf16.h>
#include "gnn_fp32.cuh"
#include "tc_common.cuh"

using namespace aq;
using namespace aqtc;

#ifndef TC3_TIMING
#define TC3_TIMING 0     // 1: per-phase clock64 accounting by thread 0 of group 0 of CTA 0 (debug variant)
#endif
#if TC3_TIMING
__device__ long long g_tc3_timing[16];
#define TC3_T(slot) do { if (blockIdx.x == 0 && gtid == 0) { const long long t_ = clock64(); g_tc3_timing[slot] += t_ - t_last; t_last = t_; } } while (0)
extern "C" int aq_debug_tc2_timing(long long *out) {
    cudaMemcpyFromSymbol(out, g_tc3_timing, sizeof(long long) * 16);
    long long z[16] = {0};
    cudaMemcpyToSymbol(g_tc3_timing, z, sizeof(z));
    return 0;
}
#else
#define TC3_T(slot) do { } while (0)
#endif

namespace {

constexpr int kG3 = 2;                               // groups (boards in flight) per CTA
constexpr int kGT = 256;                             // threads per group: 128 features x 2 column halves
constexpr int kNodesPad = 96;
constexpr uint32_t kFmBlock = 16 * 512;              // feature-major tile: [3 node blocks of 32][16 atoms of 8 features][8][64 B]
constexpr uint32_t kAdjKBlock = 48 * 128;            // adjacency block, one K-block: 48 out-node rows x 32 in-nodes (tf32, 128 B)
constexpr uint32_t kAdjBlock = 2 * kAdjKBlock;       // 64 in-nodes
constexpr uint32_t kAdjBoard = 2 * kAdjBlock;        // two blocks of out nodes: 24 KB
constexpr uint32_t kWKBlock = 128 * 128;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemGroup = 2 * kNodesPad;       // D_T | D_A
constexpr uint32_t kTmemW2 = kG3 * kTmemGroup, kTmemW3 = kTmemW2 + 64;
static_assert(kTmemW3 + 64 <= kTmemCols, "TMEM columns");

constexpr uint32_t kIdescBase = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);      // D = f32, A = B = bf16, M = 128
constexpr uint32_t kIdescL1 = kIdescBase | ((uint32_t)(kNodesPad >> 3) << 17);                 // K-major A and B, N = 96
constexpr uint32_t kIdescT = kIdescBase | (1u << 16) | ((uint32_t)(kNodesPad >> 3) << 17);     // B MN-major, N = 96
constexpr uint32_t kIdescA = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24) | ((uint32_t)(48 >> 3) << 17);  // tf32, N = 48

// Loop-invariant facts about node v = (r, c) (same as version 2).  Wall slots are read through an 18-bit window of the H / V
// bitboards that starts at slot 8 r + c - 9: bit 0 = slot (r-1, c-1), 1 = (r-1, c), 8 = (r, c-1), 9 = (r, c), 10 = (r, c+1),
// 17 = (r+1, c).  Bit 31 of a blocking mask stands for "this direction does not exist".
struct NodeConst3 {
    uint32_t upm, dnm, lfm, rtm;             // slots whose wall closes the move up / down (H board) and left / right (V board)
    uint32_t pv, sh, adj01, adj23;           // valid wall-plane bits {self 9, up 1, down 17, left 8, right 10}; window shift; tile offsets
    uint32_t adj4, row_off, pad0, pad1;      // ... of the stencil positions self|up, down|left, right (0xFFFF = absent); layer-1 operand row offset
};

struct Tc3Group {
    unsigned char fm[3 * kFmBlock];          // 24 KB: X^T, the transform's B operand
    unsigned char adj[2][kAdjBoard];         // 2 x 24 KB: A_hat (tf32) [board parity][out block][K-block][48][128 B]
    unsigned char l1op[kNodesPad * 32];      // 3 KB: layer-1 node operand [96 nodes][16] K-major SWIZZLE_32B (rows >= 81 stay zero)
    float xch[128];                          // pool partials of the second column half
    uint8_t deg[128];                        // degree (1 + open directions) of node v at [16 + v]
    unsigned char pad[1024 - 512 - 128];
};
static_assert(sizeof(Tc3Group) % 1024 == 0, "group state must keep 1024-byte alignment");

struct Tc3Smem {
    unsigned char w1[128 * 32];              // layer-1 weight operand [128][16] K-major SWIZZLE_32B: [W1 | W1 | b1_hi | b1_lo | 0 | 0]
    Tc3Group g[kG3];
    NodeConst3 nc[kV];
    float2 lut[64];                          // [deg_v * 8 + deg_u] -> {dinv_v dinv_u (fp32), the same rounded to tf32}; entry 0 = closed edge
    unsigned long long mbar[kG3];            // per group: all MMAs of a phase done
    unsigned long long mbar_t[kG3];          // per group: transform done (the aggregation reads its accumulator)
    uint32_t tmem_base;
};
static_assert(offsetof(Tc3Smem, g) % 1024 == 0, "group state must be 1024-byte aligned");
static_assert(sizeof(Tc3Smem) + 1024 <= 227 * 1024, "Tc3Smem exceeds shared memory");

__device__ __forceinline__ void group_sync3(int grp) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(grp + 1), "r"(kGT) : "memory");
}
__device__ __forceinline__ uint64_t desc_fm_mn3(uint32_t saddr) {  // MN-major SWIZZLE_64B: LBO = node-block stride, SBO = 8-feature atom stride
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(kFmBlock >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void mma_ts_f16(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_tf32(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st32_3(uint32_t taddr, const uint32_t *r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                 "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
                   "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
                   "r"(r[30]), "r"(r[31]) : "memory");
}
// 16 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// relu(a), relu(b) -> packed bf16x2 (a in the low half)
__device__ __forceinline__ uint32_t cvt2_relu(float a, float b) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;\n" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
__device__ __forceinline__ uint32_t cvt2_bf16(float a, float b) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
__device__ __forceinline__ void sts128_3(uint32_t saddr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32_3(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(saddr), "r"(v) : "memory");
}
// (1 + popcount(open directions))^-1/2 without branches
__device__ __forceinline__ float dinv_sel3(int open_mask) {
    const int deg = 1 + __popc(open_mask & 15);
    float d = 1.0f;
    d = deg == 2 ? 0.70710678118654752f : d;
    d = deg == 3 ? 0.57735026918962576f : d;
    d = deg == 4 ? 0.5f : d;
    d = deg == 5 ? 0.44721359549995794f : d;
    return d;
}
// one lane of a fully converged warp
__device__ __forceinline__ bool elect_one3() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

// Eight consecutive nodes (one 16-byte chunk c8 of the thread's feature row): relu(z + bias) -> bf16.  Nodes >= 81 are written as zero so
// that the accumulator columns they produce stay finite (they are K padding of the next aggregation).
__device__ __forceinline__ void store_chunk(uint32_t row_addr, int swz, int c8, const float *z, float bias) {
    uint4 v;
    if (c8 == 11) v = make_uint4(0u, 0u, 0u, 0u);
    else if (c8 == 10) v = make_uint4(cvt2_relu(z[0] + bias, 0.f), 0u, 0u, 0u);
    else {
        v.x = cvt2_relu(z[0] + bias, z[1] + bias); v.y = cvt2_relu(z[2] + bias, z[3] + bias);
        v.z = cvt2_relu(z[4] + bias, z[5] + bias); v.w = cvt2_relu(z[6] + bias, z[7] + bias);
    }
    sts128_3(row_addr + (uint32_t)(c8 >> 2) * kFmBlock + (uint32_t)(((c8 & 3) ^ swz) << 4), v);
}

// This thread's 48 accumulator columns [48 half, 48 half + 48) of the D region at `tmem_row` -> its part of the feature row.
__device__ __forceinline__ void epilogue_store3(uint32_t tmem_row, int half, uint32_t row_addr, int swz, float bias) {
    float za[32], zb[16];
    if (half == 0) {
        tmem_ld32(tmem_row, za);            // nodes 0..31
        tmem_ld16(tmem_row + 32, zb);       // nodes 32..47
#pragma unroll
        for (int q = 0; q < 4; ++q) store_chunk(row_addr, swz, q, za + 8 * q, bias);
        store_chunk(row_addr, swz, 4, zb, bias);
        store_chunk(row_addr, swz, 5, zb + 8, bias);
    } else {
        tmem_ld16(tmem_row + 48, zb);       // nodes 48..63
        tmem_ld32(tmem_row + 64, za);       // nodes 64..95
        store_chunk(row_addr, swz, 6, zb, bias);
        store_chunk(row_addr, swz, 7, zb + 8, bias);
#pragma unroll
        for (int q = 0; q < 4; ++q) store_chunk(row_addr, swz, 8 + q, za + 8 * q, bias);
    }
}

__global__ void __launch_bounds__(kG3 * kGT, 1)
gcn_forward_tc3_kernel(const float *__restrict__ params, const unsigned char *__restrict__ prepared,
                       const AqState *__restrict__ states, int64_t B, float *__restrict__ pooled_out) {
    extern __shared__ unsigned char smem_raw[];
    Tc3Smem &sm = *reinterpret_cast<Tc3Smem *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int gtid = threadIdx.x;
    const int grp = gtid / kGT, tid = gtid % kGT;
    const int lane = tid & 127, half = tid >> 7;     // lane = feature = TMEM lane; half = which 48 accumulator columns
    Tc3Group &gs = sm.g[grp];

    // ---- one-time setup ---------------------------------------------------------------------------------------
    {
        uint4 *z = reinterpret_cast<uint4 *>(&gs.adj[0][0]);  // adjacency tiles (and the layer-1 operand behind them) start as zero;
        for (int c = tid; c < (int)((2 * kAdjBoard + kNodesPad * 32) / 16); c += kGT) z[c] = make_uint4(0u, 0u, 0u, 0u);  // only stencil positions change
    }
    if (gtid < kV) {
        const int v = gtid, r = v / 9, c = v - 9 * r;
        NodeConst3 k;
        const uint32_t none = 0x80000000u;
        k.upm = r >= 1 ? ((c >= 1 ? 1u : 0u) | (c <= 7 ? 2u : 0u)) : none;                 // H slots (r-1, c-1), (r-1, c)
        k.dnm = r <= 7 ? ((c >= 1 ? 1u << 8 : 0u) | (c <= 7 ? 1u << 9 : 0u)) : none;       // H slots (r, c-1), (r, c)
        k.lfm = c >= 1 ? ((r >= 1 ? 1u : 0u) | (r <= 7 ? 1u << 8 : 0u)) : none;            // V slots (r-1, c-1), (r, c-1)
        k.rtm = c <= 7 ? ((r >= 1 ? 2u : 0u) | (r <= 7 ? 1u << 9 : 0u)) : none;            // V slots (r-1, c), (r, c)
        k.pv = ((r <= 7 && c <= 7) ? 1u << 9 : 0u) | ((r >= 1 && c <= 7) ? 2u : 0u) | ((r <= 6 && c <= 7) ? 1u << 17 : 0u) |
               ((r <= 7 && c >= 1) ? 1u << 8 : 0u) | ((r <= 7 && c <= 6) ? 1u << 10 : 0u);
        k.sh = (uint32_t)(8 * r + c);
        const int blk = v >= 48 ? 1 : 0, row = v - 48 * blk, kl0 = v - 32 * blk;  // row and self position inside the block's 64-node window
        auto off = [&](int kl, bool exists) -> uint32_t {  // tf32: 32 elements (128 B) per K-block row, 16-byte chunks swizzled by the row
            return exists ? (uint32_t)blk * kAdjBlock + (uint32_t)(kl >> 5) * kAdjKBlock + (uint32_t)row * 128u +
                                (uint32_t)((((kl & 31) >> 2) ^ (row & 7)) << 4) + (uint32_t)(kl & 3) * 4u
                          : 0xFFFFu;
        };
        k.adj01 = off(kl0, true) | (off(kl0 - 9, r >= 1) << 16);
        k.adj23 = off(kl0 + 9, r <= 7) | (off(kl0 - 1, c >= 1) << 16);
        k.adj4 = off(kl0 + 1, c <= 7);
        k.row_off = sw32_chunk(v, 0);  // chunk 0 of row v in the K-major SWIZZLE_32B layer-1 operand; chunk 1 = ^ 16
        k.pad0 = k.pad1 = 0u;
        sm.nc[v] = k;
    } else if (gtid >= 128 && gtid < 192) {
        const int i = gtid - 128, a = i >> 3, b2 = i & 7;
        float cf = 0.f;
        if (a >= 1 && a <= 5 && b2 >= 1 && b2 <= 5) cf = dinv_sel3((1 << (a - 1)) - 1) * dinv_sel3((1 << (b2 - 1)) - 1);  // popcount(2^k - 1) = k
        sm.lut[i] = make_float2(cf, __uint_as_float((__float_as_uint(cf) + 0x1000u) & 0xFFFFE000u));  // second: rounded to tf32
    }
    if (tid < 32) reinterpret_cast<uint32_t *>(gs.deg)[tid] = 0x01010101u;
    if (gtid >= 256 && gtid < 256 + kH) {
        const int n = gtid - 256;
        uint4 c0, c1;
        if (prepared) {  // layer-1 operand as built by aq_prepare_inference
            c0 = __ldg(reinterpret_cast<const uint4 *>(prepared + kPrepW1 + sw32_chunk(n, 0)));
            c1 = __ldg(reinterpret_cast<const uint4 *>(prepared + kPrepW1 + sw32_chunk(n, 1)));
        } else {
            float w[kF];
#pragma unroll
            for (int f = 0; f < kF; ++f) w[f] = __ldg(params + kOffW1 + n * kF + f);
            const float bias = __ldg(params + kOffB1 + n);
            const float bias_hi = __bfloat162float(__float2bfloat16_rn(bias));
            c0.x = pack_bf16(w[0], w[1]); c0.y = pack_bf16(w[2], w[3]); c0.z = pack_bf16(w[4], w[5]); c0.w = pack_bf16(w[0], w[1]);
            c1.x = pack_bf16(w[2], w[3]); c1.y = pack_bf16(w[4], w[5]); c1.z = pack_bf16(bias_hi, bias - bias_hi); c1.w = 0u;
        }
        *reinterpret_cast<uint4 *>(sm.w1 + sw32_chunk(n, 0)) = c0;
        *reinterpret_cast<uint4 *>(sm.w1 + sw32_chunk(n, 1)) = c1;
    }
    if (gtid < kG3) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&sm.mbar_t[gtid])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&sm.mbar[gtid])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (gtid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem_base = sm.tmem_base;
    const uint32_t lane_off = (uint32_t)(((tid >> 5) & 3) * 32) << 16;  // this warp's TMEM lane quadrant (warp index mod 4)
    if (half == 0) {  // W2 (group 0) / W3 (group 1): row `lane` -> 64 TMEM columns, two bf16 per column (k = 2c, 2c + 1)
        const uint32_t dst = tmem_base + lane_off + (grp ? kTmemW3 : kTmemW2);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            uint32_t r[32];
            if (prepared) {
                const unsigned char *src = prepared + (grp ? kPrepW3 : kPrepW2);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + sw128_chunk(lane, hh * 8 + j, kWKBlock)));
                    r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
                }
            } else {
                const float4 *W = reinterpret_cast<const float4 *>(params + (grp ? kOffW3 : kOffW2) + lane * kH + hh * 64);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 v = __ldg(W + j);
                    r[2 * j] = pack_bf16(v.x, v.y); r[2 * j + 1] = pack_bf16(v.z, v.w);
                }
            }
            tmem_st32_3(dst + hh * 32, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    const uint32_t tmem_t = tmem_base + (uint32_t)grp * kTmemGroup;     // transform accumulator (also the aggregation's A operand)
    const uint32_t tmem_a = tmem_t + kNodesPad;                         // aggregation accumulator
    const uint32_t bar = smem_u32(&sm.mbar[grp]);
    const uint32_t fm_addr = smem_u32(gs.fm), adj_addr = smem_u32(&gs.adj[0][0]), l1_addr = smem_u32(gs.l1op);
    const uint32_t w1_addr = smem_u32(sm.w1);
    const uint32_t row_addr = fm_addr + (uint32_t)(lane >> 3) * 512u + (uint32_t)(lane & 7) * 64u;  // this thread's feature row
    const int swz = (lane & 7) >> 1;
    const float bias2 = __ldg(params + kOffB2 + lane), bias3 = __ldg(params + kOffB3 + lane);
    uint32_t phase = 0;
    // The MMAs of a group are issued by its first warp from WARP-UNIFORM values (everything below derives from a shuffled warp
    // index), so that descriptors live in uniform registers and one tcgen05.mma costs a few instructions instead of a
    // per-thread register -> uniform register broadcast loop (~80 cycles per MMA, measured).
    const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int grp_u = warp_u >> 3;
    const bool issuer_warp = (warp_u & 7) == 0;
    const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t tmem_t_u = tmem_base_u + (uint32_t)grp_u * kTmemGroup, tmem_a_u = tmem_t_u + kNodesPad;
    const uint32_t smem_u = smem_u32(&sm);
    const uint32_t fm_u = smem_u + (uint32_t)offsetof(Tc3Smem, g) + (uint32_t)grp_u * (uint32_t)sizeof(Tc3Group);
    const uint32_t adj_u = fm_u + (uint32_t)offsetof(Tc3Group, adj), l1_u = fm_u + (uint32_t)offsetof(Tc3Group, l1op);
    const uint32_t w1_u = smem_u + (uint32_t)offsetof(Tc3Smem, w1);
    const uint32_t bar_u = smem_u + (uint32_t)offsetof(Tc3Smem, mbar) + (uint32_t)grp_u * 8u;
    const uint32_t bar_t_u = smem_u + (uint32_t)offsetof(Tc3Smem, mbar_t) + (uint32_t)grp_u * 8u;
    uint32_t phase_t = 0;

    const int64_t stride = (int64_t)gridDim.x * kG3;
    uint4 pre_a = make_uint4(0u, 0u, 0u, 0u);
    uint32_t pre_b = 0u;
    if ((int64_t)blockIdx.x * kG3 + grp < B && tid < kV) {  // states are fetched one board ahead of their node phase
        pre_a = __ldg(reinterpret_cast<const uint4 *>(states + (int64_t)blockIdx.x * kG3 + grp));
        pre_b = __ldg(reinterpret_cast<const uint32_t *>(states + (int64_t)blockIdx.x * kG3 + grp) + 4);
    }
    // Node phase of board bn: fills the adjacency tile at adj_dst and the layer-1 node operand.  It contains one group barrier
    // (all threads of the group call it).
    auto node_phase = [&](int64_t bn, uint32_t adj_dst) {
        // ---- part 1: open directions of node v from two bitboard windows; degree -> shared memory ------------------------
        uint32_t wH = 0u, wV = 0u, meta = 0u;
        int m = 0, dv = 1;
        if (tid < kV) {
            const u64 h = ((u64)pre_a.y << 32) | pre_a.x, vw = ((u64)pre_a.w << 32) | pre_a.z;
            meta = pre_b;
            if (bn + stride < B) {
                pre_a = __ldg(reinterpret_cast<const uint4 *>(states + bn + stride));
                pre_b = __ldg(reinterpret_cast<const uint32_t *>(states + bn + stride) + 4);
            }
            const uint4 k0 = *reinterpret_cast<const uint4 *>(&sm.nc[tid].upm);
            const uint32_t sh = sm.nc[tid].sh;
            wH = (uint32_t)((((unsigned __int128)h) << 9) >> sh);
            wV = (uint32_t)((((unsigned __int128)vw) << 9) >> sh);
            const uint32_t eH = wH | 0x80000000u, eV = wV | 0x80000000u;
            m = ((eH & k0.x) == 0u ? 1 : 0) | ((eH & k0.y) == 0u ? 2 : 0) | ((eV & k0.z) == 0u ? 4 : 0) | ((eV & k0.w) == 0u ? 8 : 0);
            dv = 1 + __popc(m);
            gs.deg[16 + tid] = (uint8_t)dv;
        }
        group_sync3(grp);
        // ---- part 2: A_hat row of v -> adjacency tile (tf32 from a table), and the layer-1 node operand row
        //      [hi(A_hat x0) (6) | lo(A_hat x0) (6) | 1 | 1 | 0 | 0]; the six planes of pieces_array (game_logic.py:56-93) at v and
        //      its neighbours are read from the same windows ---------------------------------------------------------------
        if (tid < kV) {
            const int v = tid;
            const uint4 k1 = *reinterpret_cast<const uint4 *>(&sm.nc[tid].pv);
            const uint2 k2 = *reinterpret_cast<const uint2 *>(&sm.nc[tid].adj4);
            const int du = gs.deg[16 + v - 9], dd = gs.deg[16 + v + 9], dl = gs.deg[16 + v - 1], dr = gs.deg[16 + v + 1];
            const float2 e0 = sm.lut[dv * 9];
            const float2 eu = sm.lut[(m & 1) ? dv * 8 + du : 0], ed = sm.lut[(m & 2) ? dv * 8 + dd : 0];
            const float2 el = sm.lut[(m & 4) ? dv * 8 + dl : 0], er = sm.lut[(m & 8) ? dv * 8 + dr : 0];
            const float c0 = e0.x, cu = eu.x, cd = ed.x, cl = el.x, cr = er.x;
            {
                const uint32_t o0 = k1.z & 0xFFFFu, o1 = k1.z >> 16, o2 = k1.w & 0xFFFFu, o3 = k1.w >> 16, o4 = k2.x;
                sts32_3(adj_dst + o0, __float_as_uint(e0.y));
                if (o1 != 0xFFFFu) sts32_3(adj_dst + o1, __float_as_uint(eu.y));
                if (o2 != 0xFFFFu) sts32_3(adj_dst + o2, __float_as_uint(ed.y));
                if (o3 != 0xFFFFu) sts32_3(adj_dst + o3, __float_as_uint(el.y));
                if (o4 != 0xFFFFu) sts32_3(adj_dst + o4, __float_as_uint(er.y));
            }
            float s[kF];
            {
                const int dp = (int)(meta & 0xFF) - v, de = (int)((meta >> 16) & 0xFF) - v;
                const float pw = (float)((meta >> 8) & 0xFF), ew = (float)(meta >> 24);
                const uint32_t pH = wH & k1.x, pV = wV & k1.x;
                auto onehot = [&](int d) {  // same summation order as a fused multiply-add chain over {self, up, down, left, right}
                    float t = d == 0 ? c0 : 0.f;
                    t += d == -9 ? cu : 0.f; t += d == 9 ? cd : 0.f; t += d == -1 ? cl : 0.f; t += d == 1 ? cr : 0.f;
                    return t;
                };
                auto plane = [&](uint32_t w) {
                    float t = (w & (1u << 9)) ? c0 : 0.f;
                    t += (w & 2u) ? cu : 0.f; t += (w & (1u << 17)) ? cd : 0.f; t += (w & (1u << 8)) ? cl : 0.f; t += (w & (1u << 10)) ? cr : 0.f;
                    return t;
                };
                auto scaled = [&](float x) { return fmaf(cr, x, fmaf(cl, x, fmaf(cd, x, fmaf(cu, x, c0 * x)))); };
                s[0] = onehot(dp); s[1] = scaled(pw); s[2] = onehot(de); s[3] = scaled(ew); s[4] = plane(pH); s[5] = plane(pV);
            }
            uint4 c0v, c1v;
            c0v.x = cvt2_bf16(s[0], s[1]); c0v.y = cvt2_bf16(s[2], s[3]); c0v.z = cvt2_bf16(s[4], s[5]);
            c0v.w = cvt2_bf16(s[0] - __uint_as_float(c0v.x << 16), s[1] - __uint_as_float(c0v.x & 0xFFFF0000u));
            c1v.x = cvt2_bf16(s[2] - __uint_as_float(c0v.y << 16), s[3] - __uint_as_float(c0v.y & 0xFFFF0000u));
            c1v.y = cvt2_bf16(s[4] - __uint_as_float(c0v.z << 16), s[5] - __uint_as_float(c0v.z & 0xFFFF0000u));
            c1v.z = 0x3F803F80u; c1v.w = 0u;  // 1, 1, 0, 0
            sts128_3(l1_addr + k2.y, c0v);
            sts128_3(l1_addr + (k2.y ^ 16u), c1v);
        }
    };
    uint32_t par = 0;  // adjacency buffer of the current board
    {
        const int64_t b0 = (int64_t)blockIdx.x * kG3 + grp;
        if (b0 < B) node_phase(b0, adj_addr);
    }
#if TC3_TIMING
    long long t_last = clock64();
#endif
    for (int64_t b = (int64_t)blockIdx.x * kG3 + grp; b < B; b += stride) {
        // ---- layer 1: one K = 16 MMA into the transform accumulator ------------------------------------------------------------
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        group_sync3(grp);
        TC3_T(3);
        if (issuer_warp) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (elect_one3()) {
                mma_bf16(tmem_t_u, desc_sw32(w1_u), desc_sw32(l1_u), kIdescL1, 0u);
                mma_commit(bar_u);
            }
            __syncwarp();
        }
        mbar_spin(bar, phase);
        phase ^= 1u;
        TC3_T(4);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        epilogue_store3(tmem_t + lane_off, half, row_addr, swz, 0.f);   // ReLU -> bf16 -> X1^T (the bias is inside the MMA)
        float pool = 0.f;
#pragma unroll 1
        for (int layer = 1; layer < kLayers; ++layer) {
            // ---- transform (bf16) and aggregation (tf32, reading the transform's accumulator in place), one commit ----------
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            group_sync3(grp);
            TC3_T(5);
            {
                const uint32_t par_u = __shfl_sync(0xffffffffu, par, 0);
                if (issuer_warp) {
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    if (elect_one3()) {
                        const uint32_t w_tmem = tmem_base_u + (layer == 1 ? kTmemW2 : kTmemW3);
#pragma unroll
                        for (int k = 0; k < 8; ++k)  // K = 128 features = 8 x 16: 8 TMEM columns of A, two 8-feature atoms of B per step
                            mma_ts_f16(tmem_t_u, w_tmem + k * 8, desc_fm_mn3(fm_u + k * 1024), kIdescT, k > 0 ? 1u : 0u);
                        // The aggregation reads the transform's accumulator as its A operand.  Consecutive MMAs are NOT interlocked on
                        // such a TMEM read-after-write (measured: back-to-back issue gives wrong sums), so the issuing thread waits for
                        // the transform's completion before it issues the aggregation.
                        mma_commit(bar_t_u);
                        mbar_spin(bar_t_u, phase_t);
                        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                        const uint32_t adj_cur = adj_u + par_u * kAdjBoard;
#pragma unroll
                        for (int blk = 0; blk < 2; ++blk) {
#pragma unroll
                            for (int s = 0; s < 8; ++s)  // 64 in-nodes = 8 K steps of 8 tf32: 8 accumulator columns of A, 32 B of B per step
                                mma_ts_tf32(tmem_a_u + blk * 48, tmem_t_u + blk * 32 + s * 8,
                                            desc_sw128(adj_cur + (uint32_t)blk * kAdjBlock + (uint32_t)(s >> 2) * kAdjKBlock + (uint32_t)(s & 3) * 32u),
                                            kIdescA, s ? 1u : 0u);
                        }
                        mma_commit(bar_u);
                    }
                    __syncwarp();
                }
                phase_t ^= 1u;
            }
            TC3_T(6);
            // while the last layer's MMAs run: the node phase of this group's next board (other adjacency buffer)
            if (layer + 1 == kLayers && b + stride < B) node_phase(b + stride, adj_addr + (par ^ 1u) * kAdjBoard);
            TC3_T(7);
            mbar_spin(bar, phase);
            phase ^= 1u;
            TC3_T(8);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (layer + 1 < kLayers) {  // + bias -> ReLU -> bf16 -> X^T row of the next layer
                epilogue_store3(tmem_a + lane_off, half, row_addr, swz, bias2);
            } else {                    // last layer feeds only the mean pool
                float za[32], zb[16];
                if (half == 0) {
                    tmem_ld32(tmem_a + lane_off, za);
                    tmem_ld16(tmem_a + lane_off + 32, zb);
#pragma unroll
                    for (int i = 0; i < 32; ++i) pool += fmaxf(za[i] + bias3, 0.f);
#pragma unroll
                    for (int i = 0; i < 16; ++i) pool += fmaxf(zb[i] + bias3, 0.f);
                } else {
                    tmem_ld16(tmem_a + lane_off + 48, zb);
                    tmem_ld32(tmem_a + lane_off + 64, za);
#pragma unroll
                    for (int i = 0; i < 16; ++i) pool += fmaxf(zb[i] + bias3, 0.f);
#pragma unroll
                    for (int i = 0; i < kV - 64; ++i) pool += fmaxf(za[i] + bias3, 0.f);
                }
            }
        }
        // ---- global_mean_pool: the two column halves of a feature meet in shared memory --------------------------------------
        if (half == 1) gs.xch[lane] = pool;
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        group_sync3(grp);
        if (half == 0) pooled_out[b * kH + lane] = (pool + gs.xch[lane]) / (float)kV;
        par ^= 1u;
        TC3_T(9);
        // (gs.xch is rewritten only after the next board's group barriers)
    }
    // ---- teardown ---------------------------------------------------------------------------------------------------
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (gtid < 32) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

}  // namespace

// Inference trunk, version 3.  Same contract as aq_gcn_forward_tc(saved == nullptr).
int aq_gcn_forward_tc3(const float *params, const void *prepared, const AqState *states, int64_t B, float *pooled, cudaStream_t st) {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const size_t smem = sizeof(Tc3Smem) + 1024;
    const int64_t want = (B + kG3 - 1) / kG3;
    const unsigned grid = (unsigned)(want < sms ? want : sms);
    cudaError_t e = cudaFuncSetAttribute(gcn_forward_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return aq_set_error((int)e, "gcn_forward_tc3 smem");
    gcn_forward_tc3_kernel<<<grid, kG3 * kGT, smem, st>>>(params, reinterpret_cast<const unsigned char *>(prepared), states, B, pooled);
    return aq_check_launch("gcn_forward_tc3_kernel");
}
