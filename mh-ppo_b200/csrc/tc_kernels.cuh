// tc_kernels.cuh -- tensor-core (tcgen05) versions of the forward-only MLP kernels of the PPO loop:
//   k_value_stats_tc : critic forward + advantage statistics (train_model_c, PY:784-787)
//   k_policy_act_tc  : per-step continuous action of the rollout (PY:434-453)
// Same inputs / outputs / sample mapping as the FFMA kernels in ppo_update.cuh / ppo_rollout.cuh, which remain as the
// exact-fp32 cross-check (MHPPO_MLP=ffma) and serve the wider choice net.
#pragma once
#include "ppo_rollout.cuh"
#include "ppo_update.cuh"
#include "tc_mlp.cuh"

namespace mhppo {

constexpr int kTcCols = 64;       // TMEM columns per CTA (widest layer)

struct TcShared {
    uint64_t bar;
    uint32_t tmem_base;
    uint32_t fail;
};

__device__ __forceinline__ uint32_t tc_prologue(TcShared *sh) {
    if (threadIdx.x == 0) { tc::mbar_init(&sh->bar, 1); sh->fail = 0; }
    if ((threadIdx.x >> 5) == 0) tc::tmem_alloc(&sh->tmem_base, kTcCols);
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    return sh->tmem_base;
}
__device__ __forceinline__ void tc_epilogue(uint32_t tmem) {
    tc::fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tc::tmem_dealloc(tmem, kTcCols);
}

__global__ void __launch_bounds__(128) k_value_stats_tc(SampleSet ss, const float *__restrict__ critic, const float *__restrict__ rtg,
                                                        float *__restrict__ V, double *__restrict__ partial /* [grid][3] */,
                                                        int *__restrict__ fail_flag) {
    constexpr int KP = 16;
    extern __shared__ __align__(1024) float smem[];
    __shared__ TcShared sh;
    __shared__ double red[4];
    tcm::NetTiles<KP> w;
    w.carve(smem);
    float *ah = smem + ((tcm::NetTiles<KP>::FLOATS + 255) & ~255), *al = ah + 128 * H2;
    w.stage(critic);
    const uint32_t tmem = tc_prologue(&sh);
    uint32_t phase = 0;
    bool ok = true;
    double sA = 0.0, sAA = 0.0, cnt = 0.0;
    for (int64_t base = (int64_t)blockIdx.x * 128; base < ss.Q; base += (int64_t)gridDim.x * 128) {
        int64_t s = 0;
        const bool sel = map_sample(ss, base + threadIdx.x, s);
        float x[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) x[k] = (sel && k < ss.D) ? ss.x[(int64_t)k * ss.S + s] : 0.f;
        tcm::put_row<KP>(ah, al, threadIdx.x, x);
        tc::fence_async_smem(); tc::fence_before(); __syncthreads(); tc::fence_after();
        const float v = tcm::forward_tile<KP>(w, ah, al, tmem, &sh.bar, phase, ok).x;
        if (sel) {
            V[s] = v;
            const double A = (double)(rtg[s] - v);
            sA += A; sAA += A * A; cnt += 1.0;
        }
    }
    if (!ok) atomicExch(fail_flag, 1);
    const double t0 = team_sum(sA, red, threadIdx.x, 128, 0), t1 = team_sum(sAA, red, threadIdx.x, 128, 0),
                 t2 = team_sum(cnt, red, threadIdx.x, 128, 0);
    if (threadIdx.x == 0) { partial[blockIdx.x * 3 + 0] = t0; partial[blockIdx.x * 3 + 1] = t1; partial[blockIdx.x * 3 + 2] = t2; }
    tc_epilogue(tmem);
}

// thread = (env, car); per pedestrian the CTA runs the cross net and/or the wait net on its 128-row tile and every row keeps
// the output of the net its decision selected (PY:441-446)
__global__ void __launch_bounds__(128) k_policy_act_tc(RolloutDims d, const float *__restrict__ net_cross, const float *__restrict__ net_wait,
                                                       ActIO io, int *__restrict__ fail_flag) {
    constexpr int KP = 16;
    extern __shared__ __align__(1024) float smem[];
    __shared__ TcShared sh;
    tcm::NetTiles<KP> wc, ww;
    constexpr int NF = (tcm::NetTiles<KP>::FLOATS + 255) & ~255;
    wc.carve(smem); ww.carve(smem + NF);
    float *ah = smem + 2 * NF, *al = ah + 128 * H2;
    wc.stage(net_cross); ww.stage(net_wait);
    const uint32_t tmem = tc_prologue(&sh);
    uint32_t phase = 0;
    bool ok = true;
    // persistent CTA: the two nets are staged once and the CTA walks (env block, car) work items
    const int64_t nblk = (d.N + 127) / 128;
    for (int64_t item = blockIdx.x; item < nblk * d.C; item += gridDim.x) {
    const int64_t n = (item % nblk) * 128 + threadIdx.x;
    const bool live = n < d.N;
    const int64_t nn = live ? n : d.N - 1;
    const int i = (int)(item / nblk);
    const ObsView v = obs_view(d, io.obs, nn);
    float mean = d.head.acc_hi;                           // car_b[1,0], PY:436
    float st[13], x[KP];
    feat_c(v, i, 0, st);                                 // state_c_tensor starts as ped 0's features, PY:437
    x[13] = x[14] = x[15] = 0.f;
    for (int p = 0; p < d.P; ++p) {
        const bool ex = (feat_c(v, i, p, x) || d.legacy) && live;      // PY:440
        const int sel = (io.action_d[(int64_t)(i * d.P + p) * d.N + nn] <= 0) ? 0 : 1;
        const int need_cross = __syncthreads_or(ex && sel == 0), need_wait = __syncthreads_or(ex && sel == 1);
        float out = 0.f;
        if (need_cross) {
            tcm::put_row<KP>(ah, al, threadIdx.x, x);
            tc::fence_async_smem(); tc::fence_before(); __syncthreads(); tc::fence_after();
            const float o = tcm::forward_tile<KP>(wc, ah, al, tmem, &sh.bar, phase, ok).x;
            if (sel == 0) out = o;
        }
        if (need_wait) {
            tcm::put_row<KP>(ah, al, threadIdx.x, x);
            tc::fence_async_smem(); tc::fence_before(); __syncthreads(); tc::fence_after();
            const float o = tcm::forward_tile<KP>(ww, ah, al, tmem, &sh.bar, phase, ok).x;
            if (sel == 1) out = o;
        }
        if (ex) {
            const float m = tanhf(out) * d.head.std + d.head.mean;  // head type 1, PY:88-90
            mean = fminf(mean, m);
            if (m == mean) {                              // PY:449-450
#pragma unroll
                for (int k = 0; k < 13; ++k) st[k] = x[k];
            }
        }
    }
    if (live) {
        const uint32_t iteration = io.iter_dev ? *io.iter_dev : io.iteration;
        const PhiloxBlock b = policy_block(d, n, (uint32_t)(io.t * d.C + i), 1u | (iteration << 8));
        const double z = sqrt(-2.0 * log(1.0 - u53(b.w0, b.w1))) * cos(2.0 * 3.141592653589793 * u53(b.w2, b.w3));
        const float a = mean + d.head.sigma * (float)z;
        const float lp = -((a - mean) * (a - mean)) * d.head.inv_2var - d.head.logp_c;
        io.actions[(int64_t)i * d.N + n] = a;
        io.actions[(int64_t)(d.C + i) * d.N + n] = io.light[(int64_t)i * d.N + n];
        const int64_t S = (int64_t)io.T * d.C * d.N, s = ((int64_t)io.t * d.C + i) * d.N + n;
#pragma unroll
        for (int k = 0; k < 13; ++k) io.obs_c[(int64_t)k * S + s] = st[k];
        io.act[s] = a; io.logp[s] = lp;
    }
    }
    if (!ok) atomicExch(fail_flag, 1);
    tc_epilogue(tmem);
}

}  // namespace mhppo
