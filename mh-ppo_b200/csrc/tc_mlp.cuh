// tc_mlp.cuh -- the 4-layer policy / value MLP on the 5th-generation tensor cores.
//
// One CTA (128 threads) owns a tile of 128 samples; thread = sample = accumulator row = TMEM lane.  Every dense layer is
// D[128 x J] = A[128 x K] * W[J x K]^T with `tcgen05.mma kind::tf32`, operands in shared memory in the canonical K-major
// layout (tc.cuh), accumulator in TMEM.  Operands are split hi + lo (hi = top 19 bits, lo = remainder) and each layer is
// three MMA chains A_hi*W_hi + A_lo*W_hi + A_hi*W_lo into the same accumulator ("3xTF32"), which restores ~fp32 accuracy
// (measured 2e-6 of the largest entry, tests/test_tc_gpu.py) -- plain tf32 (1e-3) would break the 1e-5 loss-parity bar.
// The epilogue of a layer reads the accumulator row with tcgen05.ld, adds the bias, applies ReLU and writes the next
// layer's A operand (hi and lo tiles) straight back into shared memory; weights are staged once per CTA.
#pragma once
#include "mlp.cuh"
#include "tc.cuh"

namespace mhppo { namespace tcm {

constexpr int OUTP = 16;      // output layer padded to the smallest legal N of an M = 128 MMA

__device__ __forceinline__ float hi_part(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// shared-memory image of one net: hi/lo B tiles per layer + biases
template <int KP>
struct NetTiles {
    static constexpr int W1 = H1 * KP, W2 = H2 * H1, W3 = H3 * H2, W4 = OUTP * H3;
    static constexpr int FLOATS = 2 * (W1 + W2 + W3 + W4) + H1 + H2 + H3 + OUTP;
    float *w1h, *w1l, *w2h, *w2l, *w3h, *w3l, *w4h, *w4l, *b1, *b2, *b3, *b4;
    __device__ void carve(float *p) {
        w1h = p; p += W1; w1l = p; p += W1; w2h = p; p += W2; w2l = p; p += W2; w3h = p; p += W3; w3l = p; p += W3;
        w4h = p; p += W4; w4l = p; p += W4; b1 = p; p += H1; b2 = p; p += H2; b3 = p; p += H3; b4 = p;
    }
    // flat layout (mlp.cuh): Wt[k][j]; B tile element (row = j, k) = Wt[k][j]
    __device__ void stage(const float *__restrict__ g) {
        auto fill = [&](float *h, float *l, const float *wt, int K, int J, int Jsrc, int Jpad) {
            for (int i = threadIdx.x; i < Jpad * K; i += blockDim.x) {
                const int j = i / K, k = i % K;
                const float v = (j < Jsrc) ? wt[k * J + j] : 0.f, hv = hi_part(v);
                h[tc::tile_index(j, k, K)] = hv; l[tc::tile_index(j, k, K)] = v - hv;
            }
        };
        fill(w1h, w1l, g, KP, H1, H1, H1);
        fill(w2h, w2l, g + off_w2(KP), H1, H2, H2, H2);
        fill(w3h, w3l, g + off_w3(KP), H2, H3, H3, H3);
        fill(w4h, w4l, g + off_w4(KP), H3, OP, OP, OUTP);
        for (int i = threadIdx.x; i < H1; i += blockDim.x) b1[i] = g[off_b1(KP) + i];
        for (int i = threadIdx.x; i < H2; i += blockDim.x) b2[i] = g[off_b2(KP) + i];
        for (int i = threadIdx.x; i < H3; i += blockDim.x) b3[i] = g[off_b3(KP) + i];
        for (int i = threadIdx.x; i < OUTP; i += blockDim.x) b4[i] = (i < OP) ? g[off_b4(KP) + i] : 0.f;
    }
};

// write K values of this thread's row into a hi/lo operand tile pair (canonical layout, 16-byte stores)
template <int K>
__device__ __forceinline__ void put_row(float *hi, float *lo, int row, const float *v) {
    const int base = (row >> 3) * (K * 8) + (row & 7) * 4;
#pragma unroll
    for (int c = 0; c < K / 4; ++c) {
        float4 h, l;
        h.x = hi_part(v[4 * c + 0]); h.y = hi_part(v[4 * c + 1]); h.z = hi_part(v[4 * c + 2]); h.w = hi_part(v[4 * c + 3]);
        l.x = v[4 * c + 0] - h.x; l.y = v[4 * c + 1] - h.y; l.z = v[4 * c + 2] - h.z; l.w = v[4 * c + 3] - h.w;
        *reinterpret_cast<float4 *>(hi + base + c * 32) = h;
        *reinterpret_cast<float4 *>(lo + base + c * 32) = l;
    }
}

// NC consecutive columns [col0, col0 + NC) of this thread's row of a [128 x K] operand tile pair
template <int NC>
__device__ __forceinline__ void put_cols(float *hi, float *lo, int row, int K, int col0, const float *v) {
    const int base = (row >> 3) * (K * 8) + (row & 7) * 4 + (col0 >> 2) * 32;
#pragma unroll
    for (int c = 0; c < NC / 4; ++c) {
        float4 h, l;
        h.x = hi_part(v[4 * c + 0]); h.y = hi_part(v[4 * c + 1]); h.z = hi_part(v[4 * c + 2]); h.w = hi_part(v[4 * c + 3]);
        l.x = v[4 * c + 0] - h.x; l.y = v[4 * c + 1] - h.y; l.z = v[4 * c + 2] - h.z; l.w = v[4 * c + 3] - h.w;
        *reinterpret_cast<float4 *>(hi + base + c * 32) = h;
        *reinterpret_cast<float4 *>(lo + base + c * 32) = l;
    }
}

// one thread issues the 3xTF32 chains of a layer and commits them to `bar`
template <int K, int J>
__device__ __forceinline__ void issue_layer(uint32_t d_tmem, const float *ah, const float *al, const float *bh, const float *bl, uint64_t *bar) {
    constexpr uint32_t idesc = tc::make_idesc_tf32(128, J);
    const uint32_t a0 = tc::smem_u32(ah), a1 = tc::smem_u32(al), b0 = tc::smem_u32(bh), b1 = tc::smem_u32(bl);
    bool acc = false;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t pa = (pass == 1) ? a1 : a0, pb = (pass == 2) ? b1 : b0;
#pragma unroll
        for (int k0 = 0; k0 < K; k0 += 8) {
            tc::mma_tf32(d_tmem, tc::make_desc(pa + (k0 / 4) * 128, K), tc::make_desc(pb + (k0 / 4) * 128, K), idesc, acc);
            acc = true;
        }
    }
    tc::mma_commit(bar);
}

// Forward pass of one net for the CTA's 128-sample tile.  On entry the A tile pair (ah, al) holds the KP input features
// of every row (put_row<KP>) and the writes are fenced + block-synchronised.  Returns this thread's raw outputs (OP).
// ah/al must hold 128 x 64 floats each.  `phase` is the running parity of `bar`.
template <int KP>
__device__ __forceinline__ float4 forward_tile(const NetTiles<KP> &w, float *ah, float *al, uint32_t tmem, uint64_t *bar, uint32_t &phase,
                                               bool &ok) {
    const int row = threadIdx.x, warp = threadIdx.x >> 5;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    float v[32];
    auto sync_for_mma = [&]() { tc::fence_async_smem(); tc::fence_before(); __syncthreads(); tc::fence_after(); };
    // layer 1: KP -> 32
    if (threadIdx.x == 0) issue_layer<KP, H1>(tmem, ah, al, w.w1h, w.w1l, bar);
    ok &= tc::mbar_wait(bar, phase); phase ^= 1; tc::fence_after();
    tc::tmem_ld32(trow, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + w.b1[j], 0.f);
    put_row<H1>(ah, al, row, v);
    sync_for_mma();
    // layer 2: 32 -> 64
    if (threadIdx.x == 0) issue_layer<H1, H2>(tmem, ah, al, w.w2h, w.w2l, bar);
    ok &= tc::mbar_wait(bar, phase); phase ^= 1; tc::fence_after();
    {
        float u[64];
        tc::tmem_ld32(trow, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) u[j] = fmaxf(v[j] + w.b2[j], 0.f);
        tc::tmem_ld32(trow + 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) u[32 + j] = fmaxf(v[j] + w.b2[32 + j], 0.f);
        put_row<H2>(ah, al, row, u);
    }
    sync_for_mma();
    // layer 3: 64 -> 32
    if (threadIdx.x == 0) issue_layer<H2, H3>(tmem, ah, al, w.w3h, w.w3l, bar);
    ok &= tc::mbar_wait(bar, phase); phase ^= 1; tc::fence_after();
    tc::tmem_ld32(trow, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + w.b3[j], 0.f);
    put_row<H3>(ah, al, row, v);
    sync_for_mma();
    // layer 4: 32 -> out (padded to 16 columns)
    if (threadIdx.x == 0) issue_layer<H3, OUTP>(tmem, ah, al, w.w4h, w.w4l, bar);
    ok &= tc::mbar_wait(bar, phase); phase ^= 1; tc::fence_after();
    tc::tmem_ld32(trow, v);          // columns >= 16 hold stale data of layer 3 and are ignored
    tc::fence_before();
    __syncthreads();                 // every row has read its outputs before the next tile overwrites A / TMEM
    tc::fence_after();
    return make_float4(v[0] + w.b4[0], v[1] + w.b4[1], v[2] + w.b4[2], v[3] + w.b4[3]);
}

}}  // namespace mhppo::tcm
