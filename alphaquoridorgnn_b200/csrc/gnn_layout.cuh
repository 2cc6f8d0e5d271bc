// Flat parameter / saved-activation layouts shared by the forward, backward and optimiser
// kernels.  Parameter order = state_dict order of GraphPolicyValueNetwork
// (pv_network_gnn.py:24-51): gcn_layers.{0,1,2}.{lin.weight,bias}, policy_head.{0,2}, value_head.{0,2}.
#pragma once
#include <cstdint>

namespace aq {

constexpr int kV = 81;     // nodes per board
constexpr int kF = 6;      // NUM_FEATURES
constexpr int kH = 128;    // HIDDEN_DIM
constexpr int kHH = 64;    // HIDDEN_DIM // 2
constexpr int kP = 209;    // POLICY_OUTPUT_SIZE
constexpr int kLayers = 3; // NUM_GCN_LAYERS

constexpr int kOffW1 = 0;                       // [128][6]
constexpr int kOffB1 = kOffW1 + kH * kF;        // 768
constexpr int kOffW2 = kOffB1 + kH;             // 896   [128][128]
constexpr int kOffB2 = kOffW2 + kH * kH;        // 17280
constexpr int kOffW3 = kOffB2 + kH;             // 17408
constexpr int kOffB3 = kOffW3 + kH * kH;        // 33792
constexpr int kOffWP0 = kOffB3 + kH;            // 33920 [64][128]
constexpr int kOffBP0 = kOffWP0 + kHH * kH;     // 42112
constexpr int kOffWP2 = kOffBP0 + kHH;          // 42176 [209][64]
constexpr int kOffBP2 = kOffWP2 + kP * kHH;     // 55552
constexpr int kOffWV0 = kOffBP2 + kP;           // 55761 [64][128]
constexpr int kOffBV0 = kOffWV0 + kHH * kH;     // 63953
constexpr int kOffWV2 = kOffBV0 + kHH;          // 64017 [1][64]
constexpr int kOffBV2 = kOffWV2 + kHH;          // 64081
constexpr int kNumParams = kOffBV2 + 1;         // 64082

// Saved activations for backward, struct-of-arrays over the batch (all float32).
struct SavedLayout {
    int64_t B;
    __host__ __device__ int64_t x(int layer) const { return (int64_t)layer * B * kV * kH; }  // X1,X2,X3 [B][81][128]
    __host__ __device__ int64_t pooled() const { return 3 * B * kV * kH; }                   // [B][128]
    __host__ __device__ int64_t hp() const { return pooled() + B * kH; }                     // [B][64]
    __host__ __device__ int64_t hv() const { return hp() + B * kHH; }                        // [B][64]
    __host__ __device__ int64_t policy() const { return hv() + B * kHH; }                    // [B][209]
    __host__ __device__ int64_t value() const { return policy() + B * kP; }                  // [B]
    __host__ __device__ int64_t coef() const { return value() + B; }                         // [B][81][5] self,U,D,L,R
    __host__ __device__ int64_t ax0() const { return coef() + B * kV * 5; }                  // [B][81][6]  A_hat X0
    __host__ __device__ int64_t total() const { return ax0() + B * kV * kF; }
};

// Tensor-core training path (precision 1): the x(l) regions of SavedLayout hold bf16 data instead of fp32 --
//   tc_xfm(l): post-ReLU activations of layer l+1, FEATURE-major bf16 [B][128][96] (zero beyond node 80);
//   tc_a1t:    the layer-1 node operand transposed, bf16 [B][16][96], stored behind tc_xfm(0) inside x(0);
//   coef():    A_hat coefficients as [B][81][4] floats {c0, cu, cd, cl}.
// x(l) has room for B*81*128 floats = B*41,472 bytes; the two views need B*(24,576 + 3,072) bytes.
__host__ __device__ inline unsigned short *tc_xfm(float *saved, int64_t B, int layer) {
    return reinterpret_cast<unsigned short *>(saved + SavedLayout{B}.x(layer));
}
__host__ __device__ inline unsigned short *tc_a1t(float *saved, int64_t B) {
    return reinterpret_cast<unsigned short *>(saved + SavedLayout{B}.x(0)) + B * kH * 96;
}

// Version-2 tensor-core training path (gnn_tc2.cu forward with kSave, gnn_tc2_bwd.cu): the x(0) and x(1) regions hold, per board,
//   x(0): [B][24576 B] X1^T tiles (bf16, feature-major SWIZZLE_64B, exactly the shared-memory bytes), then per board 6144 B:
//         the layer-1 node operand transposed, bf16 [16][96] as a K-major SWIZZLE_64B tile (3072 B), and the A_hat coefficients
//         [81][8] floats {self, up, down, left, right, 0, 0, 0} rounded to tf32 (2592 B);
//   x(1): [B][24576 B] X2^T tiles, then per board the ReLU mask of layer 3: [128 features][4 words] (81 bits used).
// x(l) has room for B * 41,472 bytes.
struct Tc2Saved {
    int64_t B;
    static constexpr int64_t kTile = 3 * 16 * 512;
    __host__ __device__ unsigned char *base(float *saved, int l) const { return reinterpret_cast<unsigned char *>(saved + SavedLayout{B}.x(l)); }
    __host__ __device__ unsigned char *xt(float *saved, int l, int64_t b) const { return base(saved, l) + b * kTile; }
    __host__ __device__ unsigned char *a1t(float *saved, int64_t b) const { return base(saved, 0) + B * kTile + b * 6144; }
    __host__ __device__ float *coef(float *saved, int64_t b) const { return reinterpret_cast<float *>(a1t(saved, b) + 3072); }
    __host__ __device__ unsigned char *mask3(float *saved, int64_t b) const { return base(saved, 1) + B * kTile + b * 2048; }
    // byte offset of element (row k < 16, node v < 96) in the K-major SWIZZLE_64B [16][96] tile (three 32-node blocks of 1024 B)
    __host__ __device__ static uint32_t a1t_off(int k, int v) {
        return (uint32_t)(v >> 5) * 1024u + (uint32_t)(k >> 3) * 512u + (uint32_t)(k & 7) * 64u +
               (uint32_t)((((v & 31) >> 3) ^ ((k & 7) >> 1)) << 4) + (uint32_t)(v & 7) * 2u;
    }
};

}  // namespace aq
