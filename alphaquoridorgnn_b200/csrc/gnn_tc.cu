// bf16 tensor-core path of the GCN trunk (inference): graph build + 3 GCN layers + mean pool with
// the two 128x128 node transforms on the 5th-gen tensor cores.
//
//   * one CTA per SM, persistent over boards, two 8-warp groups each owning one board at a time;
//     W2, W3 live in shared memory as bf16 in the canonical UMMA K-major SWIZZLE_128B layout for
//     the whole kernel and are shared by both groups;
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

// Two independent 8-warp groups per CTA, each working on its own board (own A tile, Z buffer, TMEM
// accumulator and mbarrier) and sharing the bf16 weight tiles: while one group waits for its MMAs
// the other runs its CUDA-core phases, and the SM holds 16 warps to hide shared-memory latency.
constexpr int kGroups = 2;
constexpr int kGroupThreads = 256;
constexpr int kTcThreads = kGroups * kGroupThreads;
constexpr int kZStride = 132;                // fp32 Z rows padded: conflict-free per-row float4 stores
constexpr uint32_t kTileBytes = 128 * 256;   // 128 rows x 128 bf16 = two K-blocks of 128 rows x 128 B
constexpr uint32_t kKBlockBytes = 128 * 128;
constexpr uint32_t kTmemCols = 128 * kGroups;

// tcgen05 instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1),
// both K-major (bits 15,16 = 0), N>>3 at bits 17-22, M>>4 at bits 24-28.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

// Rows 81..127 of each K-block of the A tile are read by the tensor core but their accumulator rows
// are never used, so those 47 x 128 B = 6016 B per K-block hold the group's small per-board arrays.
constexpr uint32_t kGap0 = kV * 128;                        // K-block 0 gap: node records | x0 | open
constexpr uint32_t kOffRec = kGap0;                         // 81 x 32 B: {c0,cu,cd,cl | cr, packed neighbour rows, -, -}
constexpr uint32_t kOffX0 = kOffRec + kV * 32;              // 488 floats
constexpr uint32_t kOffOpen = kOffX0 + 1952;                // 96 bytes
static_assert(kOffOpen + 96 <= kKBlockBytes, "K-block 0 gap overflow");
constexpr uint32_t kOffRed = kKBlockBytes + kGap0;          // K-block 1 gap: pool partials 8 x 128 floats
static_assert(kOffRed + 8 * kH * 4 <= 2 * kKBlockBytes, "K-block 1 gap overflow");

struct TcGroupSmem {
    unsigned char a[kTileBytes];  // 1024-byte aligned (SWIZZLE_128B atoms are 8 rows x 128 B)
    float z[kV * kZStride];
    unsigned char pad[1024 - (kV * kZStride * 4) % 1024];
};
static_assert(sizeof(TcGroupSmem) % 1024 == 0, "group smem must keep 1024-byte alignment");

struct TcSmem {
    unsigned char w2[kTileBytes];
    unsigned char w3[kTileBytes];
    TcGroupSmem g[kGroups];
    unsigned char w1[128 * 32];   // layer-1 B operand: bf16 [128 n][16 k], K-major SWIZZLE_32B
    unsigned long long mbar[kGroups];
    uint32_t tmem_base;
};
static_assert(sizeof(TcSmem) + 1024 <= 227 * 1024, "TcSmem exceeds shared memory");

__device__ __forceinline__ void group_sync(int grp) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(grp + 1), "r"(kGroupThreads) : "memory");
}

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

// K-major SWIZZLE_32B operand (rows of 32 B, 8-row atoms of 256 B): chunk c of row r sits at
// r*32 + ((c ^ ((r >> 2) & 1)) << 4)  (Swizzle<1,4,3>: address bit 7 XORed into bit 4); layout type 6, SBO = 256 B
__device__ __forceinline__ uint32_t sw32_chunk(int row, int c) {
    return (uint32_t)row * 32u + (uint32_t)((c ^ ((row >> 2) & 1)) << 4);
}
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
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

// relu on a packed bf16x2 word
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t x) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162 *>(&x);
    v = __hmax2(v, __floats2bfloat162_rn(0.f, 0.f));
    return *reinterpret_cast<uint32_t *>(&v);
}

__global__ void __launch_bounds__(kTcThreads, 1)
gcn_forward_tc_kernel(const float *__restrict__ params, const AqState *__restrict__ states, int64_t B,
                      float *__restrict__ pooled_out) {
    extern __shared__ unsigned char smem_raw[];
    TcSmem &sm = *reinterpret_cast<TcSmem *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int gtid = threadIdx.x;
    const int grp = gtid / kGroupThreads, tid = gtid % kGroupThreads, lane = tid & 31, warp = tid >> 5;

    for (int c = gtid; c < 2 * 128 * 16; c += kTcThreads) {  // both 128x128 weight tiles, 16-byte chunks
        const int which = c >> 11, cc = c & 2047;
        const int n = cc >> 4, j = cc & 15;
        const float *W = params + (which ? kOffW3 : kOffW2);
        const float4 lo = __ldg(reinterpret_cast<const float4 *>(W + n * kH + j * 8));
        const float4 hi = __ldg(reinterpret_cast<const float4 *>(W + n * kH + j * 8) + 1);
        uint4 v;
        v.x = pack_bf16(lo.x, lo.y); v.y = pack_bf16(lo.z, lo.w);
        v.z = pack_bf16(hi.x, hi.y); v.w = pack_bf16(hi.z, hi.w);
        *reinterpret_cast<uint4 *>((which ? sm.w3 : sm.w2) + sw128_chunk(n, j)) = v;
    }
    // Layer-1 B operand, K = 16: columns [W1 (6) | W1 (6) | b1_hi | b1_lo | 0 | 0].  The A operand
    // carries [hi(A_hat x0) (6) | lo(A_hat x0) (6) | 1 | 1 | 0 | 0], so the product is
    // (hi + lo) . bf16(W1) + b1: the input keeps ~16 mantissa bits and the bias comes for free.
    if (gtid < kH) {
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
    if (gtid < kGroups) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&sm.mbar[gtid])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (gtid < 32) {  // TMEM: 128 lanes x (128 fp32 columns per group)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");

    TcGroupSmem &gs = sm.g[grp];
    float4 *rec = reinterpret_cast<float4 *>(gs.a + kOffRec);
    float *x0 = reinterpret_cast<float *>(gs.a + kOffX0);
    uint8_t *open_s = gs.a + kOffOpen;
    float *red = reinterpret_cast<float *>(gs.a + kOffRed);
    const uint32_t tmem = sm.tmem_base + (uint32_t)grp * 128u;  // this group's accumulator columns
    const uint32_t bar = smem_u32(&sm.mbar[grp]);
    const uint32_t a_addr = smem_u32(gs.a), w1_addr = smem_u32(sm.w1), w2_addr = smem_u32(sm.w2), w3_addr = smem_u32(sm.w3);
    // aggregation mapping: half-warp per node, lane owns float4 columns l16 and l16 + 16
    const int sub = lane >> 4, l16 = lane & 15;
    const float4 b2a = __ldg(reinterpret_cast<const float4 *>(params + kOffB2) + l16);
    const float4 b2b = __ldg(reinterpret_cast<const float4 *>(params + kOffB2) + 16 + l16);
    const float4 b3a = __ldg(reinterpret_cast<const float4 *>(params + kOffB3) + l16);
    const float4 b3b = __ldg(reinterpret_cast<const float4 *>(params + kOffB3) + 16 + l16);
    const int q = warp & 3, half = warp >> 2, erow = q * 32 + lane;  // epilogue: TMEM lane quadrant / column half / row
    uint32_t phase = 0;

    for (int64_t b = (int64_t)blockIdx.x * kGroups + grp; b < B; b += (int64_t)gridDim.x * kGroups) {
        // ---- inputs: node features + open-direction masks ------------------------------------------
        {
            const AqState s = load_state(states + b);
            board_inputs_from_state(s, x0, open_s, tid);
        }
        group_sync(grp);
        // ---- per-node record: A_hat coefficients and neighbour rows (a closed direction points at
        //      the node itself with coefficient 0) ------------------------------------------------------
        if (tid < kV) {
            const int v = tid, m = open_s[v];
            const float dv = dinv_of(m);
            const int iu = (m & 1) ? v - 9 : v, id = (m & 2) ? v + 9 : v, il = (m & 4) ? v - 1 : v, ir = (m & 8) ? v + 1 : v;
            const float cu = (m & 1) ? dv * dinv_of(open_s[iu]) : 0.f, cd = (m & 2) ? dv * dinv_of(open_s[id]) : 0.f;
            const float cl = (m & 4) ? dv * dinv_of(open_s[il]) : 0.f, cr = (m & 8) ? dv * dinv_of(open_s[ir]) : 0.f;
            rec[2 * v] = make_float4(dv * dv, cu, cd, cl);
            rec[2 * v + 1] = make_float4(cr, __int_as_float(iu | (id << 8) | (il << 16) | (ir << 24)), 0.f, 0.f);
        }
        group_sync(grp);
        // ---- layer 1 A operand: row v = [hi(A_hat x0) | lo(A_hat x0) | 1 | 1 | 0 | 0] as bf16 ----------
        for (int i = tid; i < kV * 8; i += kGroupThreads) {
            const int v = i >> 3, f = i & 7;
            unsigned short hi_bits, lo_bits;
            if (f < kF) {
                const float4 r0 = rec[2 * v], r1 = rec[2 * v + 1];
                const int nb = __float_as_int(r1.y);
                float s = r0.x * x0[v * kF + f];
                s = fmaf(r0.y, x0[(nb & 0xFF) * kF + f], s);
                s = fmaf(r0.z, x0[((nb >> 8) & 0xFF) * kF + f], s);
                s = fmaf(r0.w, x0[((nb >> 16) & 0xFF) * kF + f], s);
                s = fmaf(r1.x, x0[((nb >> 24) & 0xFF) * kF + f], s);
                const __nv_bfloat16 h = __float2bfloat16_rn(s);
                const __nv_bfloat16 l = __float2bfloat16_rn(s - __bfloat162float(h));
                hi_bits = *reinterpret_cast<const unsigned short *>(&h);
                lo_bits = *reinterpret_cast<const unsigned short *>(&l);
                // columns f (chunk 0) and 6 + f (chunk 0 for f < 2, chunk 1 otherwise)
                *reinterpret_cast<unsigned short *>(gs.a + sw128_chunk(v, 0) + f * 2) = hi_bits;
                const int c = 6 + f;
                *reinterpret_cast<unsigned short *>(gs.a + sw128_chunk(v, c >> 3) + (c & 7) * 2) = lo_bits;
            } else if (f == 6) {  // columns 12..15 = 1, 1, 0, 0
                *reinterpret_cast<uint2 *>(gs.a + sw128_chunk(v, 1) + 8) = make_uint2(0x3F803F80u, 0u);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        group_sync(grp);
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            mma_bf16(tmem, umma_desc(a_addr), umma_desc_sw32(w1_addr), kIdesc, 0u);  // one K = 16 step
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        // ---- layer 1 epilogue: ReLU, bf16, straight into the layer-2 A tile ------------------------------
        if (q * 32 < kV) {
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
                const int col0 = half * 64 + cb * 32;
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
                if (erow < kV) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 pk;
                        pk.x = relu_bf16x2(pack_bf16(v[8 * i + 0], v[8 * i + 1]));
                        pk.y = relu_bf16x2(pack_bf16(v[8 * i + 2], v[8 * i + 3]));
                        pk.z = relu_bf16x2(pack_bf16(v[8 * i + 4], v[8 * i + 5]));
                        pk.w = relu_bf16x2(pack_bf16(v[8 * i + 6], v[8 * i + 7]));
                        *reinterpret_cast<uint4 *>(gs.a + sw128_chunk(erow, (col0 >> 3) + i)) = pk;
                    }
                }
            }
        }
        float pool[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pool[i] = 0.f;
#pragma unroll 1
        for (int layer = 1; layer < kLayers; ++layer) {
            // make the generic-proxy writes of the A tile visible to the tensor core (async proxy)
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            group_sync(grp);
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t w_addr = layer == 1 ? w2_addr : w3_addr;
#pragma unroll
                for (int k = 0; k < 8; ++k) {  // K = 128 = 8 x UMMA_K(16); 4 steps of 32 B per 128 B swizzle span
                    const uint32_t off = (uint32_t)(k >> 2) * kKBlockBytes + (uint32_t)(k & 3) * 32u;
                    mma_bf16(tmem, umma_desc(a_addr + off), umma_desc(w_addr + off), kIdesc, k > 0 ? 1u : 0u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            // ---- epilogue: TMEM -> registers -> fp32 Z (row = TMEM lane) ---------------------------
            if (q * 32 < kV) {  // quadrant 3 (rows 96..127) holds no board rows
#pragma unroll
                for (int cb = 0; cb < 2; ++cb) {
                    const int col0 = half * 64 + cb * 32;
                    float v[32];
                    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
                    if (erow < kV) {
                        float4 *dst = reinterpret_cast<float4 *>(gs.z + erow * kZStride + col0);
#pragma unroll
                        for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            group_sync(grp);
            // ---- aggregation + bias + ReLU: half-warp per node (two nodes per warp instruction), lane owns
            //      float4 columns l16 and l16+16; all five rows are loaded unconditionally --------------------
            {
                const float4 ba = layer == 1 ? b2a : b3a, bb = layer == 1 ? b2b : b3b;
                const bool last = layer + 1 == kLayers;
                for (int v = 2 * warp + sub; v < kV; v += 2 * (kGroupThreads / 32)) {
                    const float4 r0 = rec[2 * v], r1 = rec[2 * v + 1];
                    const int nb = __float_as_int(r1.y);
                    const float *z0 = gs.z + v * kZStride + l16 * 4;
                    const float *zu = gs.z + (nb & 0xFF) * kZStride + l16 * 4;
                    const float *zd = gs.z + ((nb >> 8) & 0xFF) * kZStride + l16 * 4;
                    const float *zl = gs.z + ((nb >> 16) & 0xFF) * kZStride + l16 * 4;
                    const float *zr = gs.z + ((nb >> 24) & 0xFF) * kZStride + l16 * 4;
                    const float4 a0 = *reinterpret_cast<const float4 *>(z0), e0 = *reinterpret_cast<const float4 *>(z0 + 64);
                    const float4 au = *reinterpret_cast<const float4 *>(zu), eu = *reinterpret_cast<const float4 *>(zu + 64);
                    const float4 ad = *reinterpret_cast<const float4 *>(zd), ed = *reinterpret_cast<const float4 *>(zd + 64);
                    const float4 al = *reinterpret_cast<const float4 *>(zl), el = *reinterpret_cast<const float4 *>(zl + 64);
                    const float4 ar = *reinterpret_cast<const float4 *>(zr), er = *reinterpret_cast<const float4 *>(zr + 64);
                    const float c0 = r0.x, cu = r0.y, cd = r0.z, cl = r0.w, cr = r1.x;
                    float s[8];
                    s[0] = fmaf(cr, ar.x, fmaf(cl, al.x, fmaf(cd, ad.x, fmaf(cu, au.x, fmaf(c0, a0.x, ba.x)))));
                    s[1] = fmaf(cr, ar.y, fmaf(cl, al.y, fmaf(cd, ad.y, fmaf(cu, au.y, fmaf(c0, a0.y, ba.y)))));
                    s[2] = fmaf(cr, ar.z, fmaf(cl, al.z, fmaf(cd, ad.z, fmaf(cu, au.z, fmaf(c0, a0.z, ba.z)))));
                    s[3] = fmaf(cr, ar.w, fmaf(cl, al.w, fmaf(cd, ad.w, fmaf(cu, au.w, fmaf(c0, a0.w, ba.w)))));
                    s[4] = fmaf(cr, er.x, fmaf(cl, el.x, fmaf(cd, ed.x, fmaf(cu, eu.x, fmaf(c0, e0.x, bb.x)))));
                    s[5] = fmaf(cr, er.y, fmaf(cl, el.y, fmaf(cd, ed.y, fmaf(cu, eu.y, fmaf(c0, e0.y, bb.y)))));
                    s[6] = fmaf(cr, er.z, fmaf(cl, el.z, fmaf(cd, ed.z, fmaf(cu, eu.z, fmaf(c0, e0.z, bb.z)))));
                    s[7] = fmaf(cr, er.w, fmaf(cl, el.w, fmaf(cd, ed.w, fmaf(cu, eu.w, fmaf(c0, e0.w, bb.w)))));
                    if (!last) {
                        // next layer's A tile: columns 4*l16..+3 = half (l16 & 1) of chunk l16/2, and the same 64 columns on
                        uint2 p0, p1;
                        p0.x = relu_bf16x2(pack_bf16(s[0], s[1])); p0.y = relu_bf16x2(pack_bf16(s[2], s[3]));
                        p1.x = relu_bf16x2(pack_bf16(s[4], s[5])); p1.y = relu_bf16x2(pack_bf16(s[6], s[7]));
                        *reinterpret_cast<uint2 *>(gs.a + sw128_chunk(v, l16 >> 1) + (l16 & 1) * 8) = p0;
                        *reinterpret_cast<uint2 *>(gs.a + sw128_chunk(v, 8 + (l16 >> 1)) + (l16 & 1) * 8) = p1;
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) pool[i] += fmaxf(s[i], 0.f);
                    }
                }
            }
        }
        // ---- global_mean_pool: combine the two half-warps, then the eight warps ---------------------------
#pragma unroll
        for (int i = 0; i < 8; ++i) pool[i] += __shfl_xor_sync(0xffffffffu, pool[i], 16);
        if (sub == 0) {
            reinterpret_cast<float4 *>(red + warp * kH)[l16] = make_float4(pool[0], pool[1], pool[2], pool[3]);
            reinterpret_cast<float4 *>(red + warp * kH)[16 + l16] = make_float4(pool[4], pool[5], pool[6], pool[7]);
        }
        group_sync(grp);
        if (tid < kH) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w * kH + tid];
            pooled_out[b * kH + tid] = s / (float)kV;
        }
        group_sync(grp);
    }
    // ---- teardown ------------------------------------------------------------------------------------
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (gtid < 32) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(sm.tmem_base), "r"(kTmemCols) : "memory");
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
    const int64_t want = (B + kGroups - 1) / kGroups;
    const unsigned grid = (unsigned)(want < sms ? want : sms);
    gcn_forward_tc_kernel<<<grid, kTcThreads, smem, st>>>(params, states, B, pooled);
    return aq_check_launch("gcn_forward_tc_kernel");
}
