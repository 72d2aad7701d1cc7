// mlp.cuh -- the policy / value networks of MH-PPO on the SM (Model_PPO, PY:42-93).
//
// Every net is the same 4-layer MLP  in -> 32 -> 64 -> 32 -> out  with ReLU, and one of three heads:
// type 0 linear (critics), type 1 `std*tanh(z)+mean` (Gaussian mean of the cross / wait actors,
// std = 3, mean = -1 from PY:1044-1045), type 2 softmax over the two logits (choice actor).
//
// Flat parameter layout used by every kernel and by Adam (host packs/unpacks state_dicts,
// mh-ppo_b200/policy.py):  W1t[KP][32] b1[32] W2t[32][64] b2[64] W3t[64][32] b3[32] W4t[32][4] b4[4]
// -- weights TRANSPOSED ([in][out], out fastest) so a thread that owns one sample streams the
// input index and reads 4 output weights per 128-bit shared-memory broadcast; the input width is
// zero-padded to KP (13 -> 16, 18/30 -> 32, 54 -> 56) and the output width to 4.
//
// Execution model of the FFMA kernels here: one thread owns one sample; activations live in shared
// memory as [sample][feature] rows with an odd stride (conflict-free for lane = sample), weights
// are staged once per CTA.  These are the exact-fp32 kernels (the 1e-5 loss-parity bar is checked
// with them, MHPPO_MLP=ffma) and they serve the wider choice nets and the rollout inference; the
// 13-input nets of the update run on the tensor cores by default (tc_mlp.cuh, tc_grad.cuh), batching
// 128 samples per CTA tile.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mhppo {

// Gaussian head of the cross / wait actors (Model_PPO head type 1, PY:88-90: tanh(z) * std + mean with mean / std derived
// from car_b, PY:1044-1045) and the exploration noise N(mu, variance) of PY:726-729 / 451-453 / 795-800:
// log_prob = -(a - mu)^2 * inv_2var - logp_c, inv_2var = 1 / (2 variance), logp_c = ln(2 pi variance) / 2.
// Filled on the host by mhppo_set_gaussian_head (defaults: mean -1, std 3, variance 0.5, acc_hi 2).
struct HeadCfg { float mean, std, sigma, inv_2var, logp_c, acc_hi; };


constexpr int H1 = 32, H2 = 64, H3 = 32, OP = 4;
constexpr int kMlpBlock = 128;

__host__ __device__ constexpr int net_params(int KP) { return KP * H1 + H1 + H1 * H2 + H2 + H2 * H3 + H3 + H3 * OP + OP; }
__host__ __device__ constexpr int off_b1(int KP) { return KP * H1; }
__host__ __device__ constexpr int off_w2(int KP) { return off_b1(KP) + H1; }
__host__ __device__ constexpr int off_b2(int KP) { return off_w2(KP) + H1 * H2; }
__host__ __device__ constexpr int off_w3(int KP) { return off_b2(KP) + H2; }
__host__ __device__ constexpr int off_b3(int KP) { return off_w3(KP) + H2 * H3; }
__host__ __device__ constexpr int off_w4(int KP) { return off_b3(KP) + H3; }
__host__ __device__ constexpr int off_b4(int KP) { return off_w4(KP) + H3 * OP; }

// row strides of the activation tiles (odd => lane = sample is bank-conflict free)
template <int KP> struct Strides { static constexpr int X = KP + 1, A1 = H1 + 1, A2 = H2 + 1, A3 = H3 + 1; };

// one dense layer for the sample owned by this thread: out[j] = b[j] + sum_k in[k] * Wt[k][j]
// `in` is this thread's row in shared memory; Wt/b are the CTA's staged weights.
// `out` may alias `in`: every input is read before the first output is written (the accumulators are registers).
template <int K, int J, bool RELU>
__device__ __forceinline__ void dense_fwd(const float *in, const float *__restrict__ Wt, const float *__restrict__ b, float *out) {
    float acc[J];
#pragma unroll
    for (int j = 0; j < J; ++j) acc[j] = b[j];
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
        const float a = in[k];
        const float4 *w = reinterpret_cast<const float4 *>(Wt + k * J);
#pragma unroll
        for (int j4 = 0; j4 < J / 4; ++j4) {
            const float4 ww = w[j4];
            acc[4 * j4 + 0] = fmaf(a, ww.x, acc[4 * j4 + 0]); acc[4 * j4 + 1] = fmaf(a, ww.y, acc[4 * j4 + 1]);
            acc[4 * j4 + 2] = fmaf(a, ww.z, acc[4 * j4 + 2]); acc[4 * j4 + 3] = fmaf(a, ww.w, acc[4 * j4 + 3]);
        }
    }
#pragma unroll
    for (int j = 0; j < J; ++j) out[j] = RELU ? fmaxf(acc[j], 0.f) : acc[j];
}

// whole forward pass for this thread's sample; x/a1/a2/a3 are its rows; returns the 4 (padded) raw outputs
template <int KP>
__device__ __forceinline__ float4 mlp_fwd_rows(const float *__restrict__ sw, const float *x, float *a1, float *a2, float *a3) {
    dense_fwd<KP, H1, true>(x, sw, sw + off_b1(KP), a1);
    dense_fwd<H1, H2, true>(a1, sw + off_w2(KP), sw + off_b2(KP), a2);
    dense_fwd<H2, H3, true>(a2, sw + off_w3(KP), sw + off_b3(KP), a3);
    float o[OP];
    dense_fwd<H3, OP, false>(a3, sw + off_w4(KP), sw + off_b4(KP), o);
    return make_float4(o[0], o[1], o[2], o[3]);
}

// forward-only variant: ONE row of kRowFwd floats per sample, every layer computed in place (65 floats instead of 148 per
// sample: 2 x 256-thread CTAs per SM instead of 1 x 128)
constexpr int kRowFwd = H2 + 1;
constexpr int kFwdBlock = 256;
template <int KP>
__device__ __forceinline__ float4 mlp_fwd_inplace(const float *__restrict__ sw, float *row) {
    dense_fwd<KP, H1, true>(row, sw, sw + off_b1(KP), row);
    dense_fwd<H1, H2, true>(row, sw + off_w2(KP), sw + off_b2(KP), row);
    dense_fwd<H2, H3, true>(row, sw + off_w3(KP), sw + off_b3(KP), row);
    float o[OP];
    dense_fwd<H3, OP, false>(row, sw + off_w4(KP), sw + off_b4(KP), o);
    return make_float4(o[0], o[1], o[2], o[3]);
}

// stage a net's flat parameters into shared memory (whole CTA)
template <int KP>
__device__ __forceinline__ void stage_net(float *sw, const float *__restrict__ g) {
    for (int i = threadIdx.x; i < net_params(KP); i += blockDim.x) sw[i] = g[i];
}

// The same staging as ONE bulk asynchronous copy per net (TMA, `cp.async.bulk`): a single thread issues it, the bytes land
// while the CTA does other work, completion is counted on an mbarrier.  Needs 16-byte aligned source / destination and a
// size that is a multiple of 16 (net_params(16) * 4 = 19 472).
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_bar_init(uint64_t *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_expect(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t polls = 0; !done; ++polls) {
        if (polls == (1u << 26)) __trap();          // a lost copy is an error, not a hang
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(smem_addr(bar)), "r"(parity)
                     : "memory");
    }
}

}  // namespace mhppo
