// Inline-PTX helpers shared by the tcgen05 kernels: shared-memory matrix descriptors (K-major SWIZZLE_128B /
// SWIZZLE_32B), tcgen05.mma / commit / ld wrappers, mbarrier wait, bf16 packing.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>

namespace aqtc {

constexpr int kGroupThreads = 128;   // one group = 4 warps = the 128 TMEM lanes

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void group_sync(int grp) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(grp + 1), "r"(kGroupThreads) : "memory");
}

// byte offset of 16-byte chunk j (0..15) of `row` inside a K-major SWIZZLE_128B tile with the given K-block size
__device__ __forceinline__ uint32_t sw128_chunk(int row, int j, uint32_t kblock) {
    return (uint32_t)(j >> 3) * kblock + (uint32_t)row * 128u + (uint32_t)(((j & 7) ^ (row & 7)) << 4);
}
// shared-memory matrix descriptors: start>>4, LBO=1 (unused for swizzled K-major), SBO, version 1, layout type
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {  // SBO = 1024 B, type 2 = SWIZZLE_128B
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// K-major SWIZZLE_32B (rows of 32 B, 8-row atoms of 256 B): chunk c of row r at r*32 + ((c ^ ((r>>2)&1)) << 4)
__device__ __forceinline__ uint32_t sw32_chunk(int row, int c) {
    return (uint32_t)row * 32u + (uint32_t)((c ^ ((row >> 2) & 1)) << 4);
}
__device__ __forceinline__ uint64_t desc_sw32(uint32_t saddr) {   // SBO = 256 B, type 6 = SWIZZLE_32B
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// one lane of a fully converged warp (elect.sync); used inside warp-uniform branches so that MMA operands stay in uniform registers
__device__ __forceinline__ bool elect_one_lane() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");  // suspend-time hint (ns): sleep, do not spin
    }
}
// 32 accumulator columns of this thread's TMEM lane
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
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&t);
}
// relu(x) -> bf16 -> 16-bit shared store (cvt.rn.relu folds the ReLU into the conversion)
__device__ __forceinline__ void st_relu_bf16(uint32_t saddr, float x) {
    asm volatile("{\n\t.reg .b16 t;\n\tcvt.rn.relu.bf16.f32 t, %1;\n\tst.shared.b16 [%0], t;\n\t}\n" ::"r"(saddr), "f"(x) : "memory");
}
__device__ __forceinline__ unsigned short bf16_bits(float x) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    return *reinterpret_cast<const unsigned short *>(&h);
}


__device__ __forceinline__ uint4 pack8_bf16(const float *f) {
    uint4 v;
    v.x = pack_bf16(f[0], f[1]); v.y = pack_bf16(f[2], f[3]); v.z = pack_bf16(f[4], f[5]); v.w = pack_bf16(f[6], f[7]);
    return v;
}

// "Prepared" inference weights: the bf16 operand tiles exactly as the kernels keep them in shared memory, built
// once by aq_prepare_inference so that every CTA only copies them (no per-CTA fp32 -> bf16 conversion).
constexpr uint32_t kPrepW2 = 0;                    // trunk W2 tile  [128][128] K-major SWIZZLE_128B   (32 KB)
constexpr uint32_t kPrepW3 = 32768;                // trunk W3 tile                                    (32 KB)
constexpr uint32_t kPrepW1 = 65536;                // trunk layer-1 operand [128][16] SWIZZLE_32B      ( 4 KB)
constexpr uint32_t kPrepHeadB1 = 69632;            // heads [Wp0 ; Wv0] tile [128][128]                (32 KB)
constexpr uint32_t kPrepHeadB2 = 102400;           // heads Wp2 tile [224][64]                         (28 KB)
constexpr uint32_t kPrepBytes = 131072;

}  // namespace aqtc
