// Building blocks of the fp32 (FFMA) GCN kernels, shared by forward and backward:
//   * one CTA of 256 threads owns one 81-node board at a time, everything lives in shared memory;
//   * the [81x128]x[128x128] node transform is a register-tiled FFMA GEMM: warp w owns rows
//     11w..11w+10, lane l owns columns 4l..4l+3 (44 accumulators per thread);
//   * the A_hat aggregation is a 5-point stencil over the board with coefficients
//     dinv_i*dinv_j (gcn_norm with self loops), warp-per-node, 32 lanes x float4 = one 128-wide row.
#pragma once
#include "aq_common.cuh"
#include "gnn_layout.cuh"

namespace aq {

constexpr int kGcnThreads = 256;
constexpr int kRowsPerWarp = 11;  // 8 warps x 11 rows = 88 >= 81

// Weight tile [128][128] in shared memory, float4 columns XOR-swizzled by the row so that both the
// transposing fill and the GEMM reads are (nearly) bank-conflict free.
AQ_DEV int wt_index(int row, int col) { return row * kH + ((((col >> 2) ^ (row & 31)) << 2) | (col & 3)); }

// dst[row=k][col=n] = W[n][k]   (W row-major [128][128] in global memory)
__device__ __forceinline__ void load_weight_transposed(float *dst, const float *__restrict__ W, int tid) {
    for (int i = tid; i < kH * kH; i += kGcnThreads) {
        const int n = i >> 7, k = i & 127;  // coalesced along k
        dst[wt_index(k, n)] = __ldg(W + i);
    }
}
// dst[row=n][col=k] = W[n][k]
__device__ __forceinline__ void load_weight_natural(float *dst, const float *__restrict__ W, int tid) {
    for (int i = tid; i < kH * kH / 4; i += kGcnThreads) {
        const int n = i >> 5, c = i & 31;
        const float4 w = __ldg(reinterpret_cast<const float4 *>(W) + i);
        *reinterpret_cast<float4 *>(dst + n * kH + ((c ^ (n & 31)) << 2)) = w;
    }
}

// acc[i][:] = sum_k A[r0+i][k] * Wt[k][4*lane .. 4*lane+3]
__device__ __forceinline__ void gemm_rows(const float *__restrict__ A, const float *__restrict__ Wt, int warp, int lane,
                                          float (&acc)[kRowsPerWarp][4]) {
#pragma unroll
    for (int i = 0; i < kRowsPerWarp; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    const float4 *A4 = reinterpret_cast<const float4 *>(A) + warp * kRowsPerWarp * (kH / 4);
    const float4 *W4 = reinterpret_cast<const float4 *>(Wt);
#pragma unroll 2
    for (int k0 = 0; k0 < kH; k0 += 4) {
        // rows k0..k0+3 share (k & 31) high bits only when k0 % 4 == 0: swizzle per row
        const float4 w0 = W4[(k0 + 0) * 32 + (lane ^ ((k0 + 0) & 31))];
        const float4 w1 = W4[(k0 + 1) * 32 + (lane ^ ((k0 + 1) & 31))];
        const float4 w2 = W4[(k0 + 2) * 32 + (lane ^ ((k0 + 2) & 31))];
        const float4 w3 = W4[(k0 + 3) * 32 + (lane ^ ((k0 + 3) & 31))];
#pragma unroll
        for (int i = 0; i < kRowsPerWarp; ++i) {
            const float4 a = A4[i * (kH / 4) + (k0 >> 2)];  // warp-wide broadcast
            acc[i][0] = fmaf(a.x, w0.x, acc[i][0]); acc[i][1] = fmaf(a.x, w0.y, acc[i][1]);
            acc[i][2] = fmaf(a.x, w0.z, acc[i][2]); acc[i][3] = fmaf(a.x, w0.w, acc[i][3]);
            acc[i][0] = fmaf(a.y, w1.x, acc[i][0]); acc[i][1] = fmaf(a.y, w1.y, acc[i][1]);
            acc[i][2] = fmaf(a.y, w1.z, acc[i][2]); acc[i][3] = fmaf(a.y, w1.w, acc[i][3]);
            acc[i][0] = fmaf(a.z, w2.x, acc[i][0]); acc[i][1] = fmaf(a.z, w2.y, acc[i][1]);
            acc[i][2] = fmaf(a.z, w2.z, acc[i][2]); acc[i][3] = fmaf(a.z, w2.w, acc[i][3]);
            acc[i][0] = fmaf(a.w, w3.x, acc[i][0]); acc[i][1] = fmaf(a.w, w3.y, acc[i][1]);
            acc[i][2] = fmaf(a.w, w3.z, acc[i][2]); acc[i][3] = fmaf(a.w, w3.w, acc[i][3]);
        }
    }
}

// A_hat coefficients of one board into coef[81][5] = {self, U, D, L, R}; closed edges get 0.
// open[v] bit d = direction d open.  dinv = (1 + popcount)^-1/2.
__device__ __forceinline__ float dinv_of(int m) {
    const int deg = 1 + __popc(m & 15);
    return deg == 1 ? 1.0f : deg == 2 ? 0.70710678118654752f : deg == 3 ? 0.57735026918962576f : deg == 4 ? 0.5f
                                                                                              : 0.44721359549995794f;
}

__device__ __forceinline__ void board_coefficients(const uint8_t *open_s, float *coef, int tid) {
    if (tid < kV) {
        const int v = tid, m = open_s[v];
        const float dv = dinv_of(m);
        coef[v * 5 + 0] = dv * dv;
        coef[v * 5 + 1] = (m & 1) ? dv * dinv_of(open_s[v - 9]) : 0.f;
        coef[v * 5 + 2] = (m & 2) ? dv * dinv_of(open_s[v + 9]) : 0.f;
        coef[v * 5 + 3] = (m & 4) ? dv * dinv_of(open_s[v - 1]) : 0.f;
        coef[v * 5 + 4] = (m & 8) ? dv * dinv_of(open_s[v + 1]) : 0.f;
    }
}

// out[v][:] = epilogue( sum_{u in N(v) + v} coef * in[u][:] ), warp per node, lane = float4 column.
// kBiasRelu: out = relu(agg + bias)   (GCNConv adds the bias after aggregation, ReLU outside)
template <bool kBiasRelu>
__device__ __forceinline__ void aggregate(const float *__restrict__ in, float *__restrict__ out,
                                          const float *__restrict__ coef, const float *__restrict__ bias, int warp,
                                          int lane) {
    const float4 *in4 = reinterpret_cast<const float4 *>(in);
    float4 *out4 = reinterpret_cast<float4 *>(out);
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kBiasRelu) bb = reinterpret_cast<const float4 *>(bias)[lane];
    for (int v = warp; v < kV; v += kGcnThreads / 32) {
        const float c0 = coef[v * 5 + 0], cu = coef[v * 5 + 1], cd = coef[v * 5 + 2], cl = coef[v * 5 + 3],
                    cr = coef[v * 5 + 4];
        float4 a = in4[v * 32 + lane];
        float4 s = make_float4(c0 * a.x, c0 * a.y, c0 * a.z, c0 * a.w);
        if (cu != 0.f) { a = in4[(v - 9) * 32 + lane]; s.x = fmaf(cu, a.x, s.x); s.y = fmaf(cu, a.y, s.y); s.z = fmaf(cu, a.z, s.z); s.w = fmaf(cu, a.w, s.w); }
        if (cd != 0.f) { a = in4[(v + 9) * 32 + lane]; s.x = fmaf(cd, a.x, s.x); s.y = fmaf(cd, a.y, s.y); s.z = fmaf(cd, a.z, s.z); s.w = fmaf(cd, a.w, s.w); }
        if (cl != 0.f) { a = in4[(v - 1) * 32 + lane]; s.x = fmaf(cl, a.x, s.x); s.y = fmaf(cl, a.y, s.y); s.z = fmaf(cl, a.z, s.z); s.w = fmaf(cl, a.w, s.w); }
        if (cr != 0.f) { a = in4[(v + 1) * 32 + lane]; s.x = fmaf(cr, a.x, s.x); s.y = fmaf(cr, a.y, s.y); s.z = fmaf(cr, a.z, s.z); s.w = fmaf(cr, a.w, s.w); }
        if (kBiasRelu) {
            s.x = fmaxf(s.x + bb.x, 0.f); s.y = fmaxf(s.y + bb.y, 0.f);
            s.z = fmaxf(s.z + bb.z, 0.f); s.w = fmaxf(s.w + bb.w, 0.f);
        }
        out4[v * 32 + lane] = s;
    }
}

// Open-direction mask of one square straight from the wall bitboards (is_wall_blocking,
// game_logic.py:145-167): an H wall in slot (x,y) blocks the vertical moves between rows x,x+1 in
// columns y,y+1; a V wall blocks the horizontal moves between columns y,y+1 in rows x,x+1.
__device__ __forceinline__ int node_open_mask(u64 h, u64 vw, int r, int c) {
    // pair of wall slots whose wall touches column c (slots c-1 and c) / row r (slots r-1 and r)
    const u64 colpair = (c < 8 ? 1ull << c : 0ull) | (c > 0 ? 1ull << (c - 1) : 0ull);
    const u64 rowpair = (r < 8 ? 1ull << (8 * r) : 0ull) | (r > 0 ? 1ull << (8 * (r - 1)) : 0ull);
    const bool up = r > 0 && ((h >> (8 * (r - 1))) & colpair) == 0;
    const bool down = r < 8 && ((h >> (8 * r)) & colpair & 0xFFull) == 0;
    const bool left = c > 0 && ((vw >> (c - 1)) & rowpair) == 0;
    const bool right = c < 8 && ((vw >> c) & rowpair) == 0;
    return (int)up | ((int)down << 1) | ((int)left << 2) | ((int)right << 3);
}

// Node features of one board into x0[81][6] and open masks into open_s[81], from a packed state.
__device__ __forceinline__ void board_inputs_from_state(const AqState &s, float *x0, uint8_t *open_s, int tid) {
    if (tid < kV) {
        const int v = tid;
        open_s[v] = (uint8_t)node_open_mask(s.hwalls, s.vwalls, v / 9, v % 9);
        // the six planes of pieces_array (game_logic.py:56-93); wall planes sit on the slot's top-left tile
        const int r = v / 9, c = v % 9;
        const bool slot_ok = r < 8 && c < 8;
        const int slot = r * 8 + c;
        x0[v * kF + 0] = v == s.ppos ? 1.f : 0.f;
        x0[v * kF + 1] = (float)s.pwalls;
        x0[v * kF + 2] = v == s.epos ? 1.f : 0.f;
        x0[v * kF + 3] = (float)s.ewalls;
        x0[v * kF + 4] = (slot_ok && ((s.hwalls >> slot) & 1)) ? 1.f : 0.f;
        x0[v * kF + 5] = (slot_ok && ((s.vwalls >> slot) & 1)) ? 1.f : 0.f;
    }
}

}  // namespace aq
