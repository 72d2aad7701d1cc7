// env_kernels.cuh -- the sm_100a kernels around env_core.cuh: env step (with in-kernel auto-reset),
// masked reset, and the per-(variant, slots) instantiation tables.
//
// Replaces Crosswalk_hybrid_multi_*.step / .reset of the reference (SC:789-946 and siblings) for
// n_envs environments at once.  Grid: one thread per env, 128-thread CTAs; a warp touches 32
// consecutive envs so every state group is one coalesced 512 B float4 access (env_state.cuh).
// Bound: HBM (state is read and written once per step) as long as the fp64 issue rate keeps up;
// DESIGN.md "Kernels" has the byte accounting and the measured roofline fraction.
#pragma once
#include <cuda_runtime.h>

#include "../../include/mhppo.h"
#include "env_state.cuh"

namespace mhppo {

constexpr int kEnvBlock = 128;

struct RngKey { uint32_t k0, k1; int64_t env_id0; };

struct StepIO {
    mhppo_view actions, obs, rewards, reward_light, term_obs;
    uint8_t *done;
    int autoreset;
};

struct ViewOut {
    float *p; int64_t cs;
    __device__ __forceinline__ void operator()(int k, float v) const { p[(int64_t)k * cs] = v; }
};
struct NullOut {
    __device__ __forceinline__ void operator()(int, float) const {}
};

template <int V, int MC, int MP>
__global__ void __launch_bounds__(kEnvBlock) k_env_step(EnvArena a, EnvConst c, RngKey key, StepIO io) {
    const int64_t n = (int64_t)blockIdx.x * kEnvBlock + threadIdx.x;
    if (n >= a.N) return;
    EnvR<MC, MP> e;
    const uint64_t gid = (uint64_t)(key.env_id0 + n);
    e.rng.env_lo = (uint32_t)gid; e.rng.env_hi = (uint32_t)(gid >> 32); e.rng.k0 = key.k0; e.rng.k1 = key.k1;
    load_env<MC, MP>(a, c, n, e);

    // action vector of this env, read once up front: acc[s] = a[s], light[s] = a[nA/2 + s]
    ActR<MC> act;
    {
        const float *ap = io.actions.ptr + n * io.actions.env_stride;
        const int half = c.nA / 2;
#pragma unroll
        for (int s = 0; s < MC; ++s) {
            act.acc[s] = (s < half) ? ap[(int64_t)s * io.actions.comp_stride] : 0.f;
            act.light[s] = (s < half) ? ap[(int64_t)(half + s) * io.actions.comp_stride] : 0.f;
        }
    }
    float *rp = io.rewards.ptr ? io.rewards.ptr + n * io.rewards.env_stride : nullptr;
    float *lp = io.reward_light.ptr ? io.reward_light.ptr + n * io.reward_light.env_stride : nullptr;
    auto rew_out = [&](int i, double r, double rl) {
        if (rp) rp[(int64_t)i * io.rewards.comp_stride] = (float)r;
        if (lp) lp[(int64_t)i * io.reward_light.comp_stride] = (float)rl;
    };
    const bool done = step_env<V, MC, MP>(c, e, act, rew_out);
    if (io.done) io.done[n] = done ? 1 : 0;

    const bool do_reset = done && io.autoreset;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const bool last = !do_reset || pass == 1;
        const mhppo_view &v = last ? io.obs : io.term_obs;
        if (v.ptr) {
            ViewOut out{v.ptr + n * v.env_stride, v.comp_stride};
            write_obs<V, MC, MP>(c, e, pass == 1, out);
        }
        if (last) break;
        EnvR<MC, MP> fresh;                 // scratch copy: only this rare branch touches local memory
        fresh.rng = e.rng;
        reset_env<V, MC, MP>(c, fresh);
        e = fresh;
    }
    store_env<MC, MP>(a, c, n, e);
}

template <int V, int MC, int MP>
__global__ void __launch_bounds__(kEnvBlock) k_env_reset(EnvArena a, EnvConst c, RngKey key, const uint8_t *mask,
                                                         mhppo_view obs) {
    const int64_t n = (int64_t)blockIdx.x * kEnvBlock + threadIdx.x;
    if (n >= a.N) return;
    if (mask && !mask[n]) return;
    EnvR<MC, MP> e;
    const uint64_t gid = (uint64_t)(key.env_id0 + n);
    e.rng.env_lo = (uint32_t)gid; e.rng.env_hi = (uint32_t)(gid >> 32); e.rng.k0 = key.k0; e.rng.k1 = key.k1;
    e.rng.ctr = f2u(a.env_e[n].w);          // the stream cursor survives resets (episodes share one stream)
    reset_env<V, MC, MP>(c, e);
    if (obs.ptr) { ViewOut out{obs.ptr + n * obs.env_stride, obs.comp_stride}; write_obs<V, MC, MP>(c, e, true, out); }
    else { NullOut nul; write_obs<V, MC, MP>(c, e, true, nul); }   // get_data still updates the running-min delta
    store_env<MC, MP>(a, c, n, e);
}

// ---- instantiation table -----------------------------------------------------------------------
struct EnvKernelEntry {
    int variant, mc, mp;
    void (*step)(EnvArena, EnvConst, RngKey, StepIO);
    void (*reset)(EnvArena, EnvConst, RngKey, const uint8_t *, mhppo_view);
};

#define MHPPO_ENV_ENTRY(V, MC, MP) { V, MC, MP, k_env_step<V, MC, MP>, k_env_reset<V, MC, MP> }

// each env_inst_*.cu defines one of these
const EnvKernelEntry *env_table_stop(int *n);
const EnvKernelEntry *env_table_naif(int *n);
const EnvKernelEntry *env_table_coop(int *n);
const EnvKernelEntry *env_table_4cars(int *n);
const EnvKernelEntry *env_table_4cars2(int *n);
const EnvKernelEntry *env_table_scalable(int *n);

}  // namespace mhppo
