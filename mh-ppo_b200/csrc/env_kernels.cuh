// env_kernels.cuh -- the sm_100a kernels around env_core.cuh: env step (with in-kernel auto-reset),
// masked reset, and the per-(variant, slots) instantiation tables.
//
// Replaces Crosswalk_hybrid_multi_*.step / .reset of the reference (SC:789-946 and siblings) for
// n_envs environments at once.  Grid: one thread per env, 128-thread CTAs; a warp touches 32
// consecutive envs so every state group is one coalesced 512 B float4 access (env_state.cuh).
// Roofline: HBM (state is read and written once per step); the measured limiter is instruction issue
// (DESIGN.md "Kernels" has the byte accounting, the measured roofline fraction and the stall breakdown).
#pragma once
#include <cuda_runtime.h>

#include "env_step.cuh"

namespace mhppo {

#ifndef MH_ENV_BLOCK
#define MH_ENV_BLOCK 128
#endif
constexpr int kEnvBlock = MH_ENV_BLOCK;
#ifndef MH_ENV_MIN_BLOCKS
#define MH_ENV_MIN_BLOCKS 5
#endif
constexpr int kEnvMinBlocks = MH_ENV_MIN_BLOCKS;   // CTAs per SM the register allocation is capped for

// dynamic shared memory of the step kernel: the CTA's car slots
template <int MC>
struct StepShared {
    CarSlots<MC, kEnvBlock> cars;
};

template <int V, int MC, int MP>
__global__ void __launch_bounds__(kEnvBlock, kEnvMinBlocks) k_env_step(const __grid_constant__ EnvArena a, const __grid_constant__ EnvConst c,
                                                                           const __grid_constant__ RngKey key, const __grid_constant__ StepIO io) {
    extern __shared__ __align__(16) unsigned char step_smem[];
    StepShared<MC> &sh = *reinterpret_cast<StepShared<MC> *>(step_smem);
    const int64_t n = io.n_begin + (int64_t)blockIdx.x * kEnvBlock + threadIdx.x;
    if (n >= io.n_end) return;
    env_step_thread<V, MC, MP, kEnvBlock>(a, c, key, io, n, sh.cars, (int)threadIdx.x);
}

template <int V, int MC, int MP>
__global__ void __launch_bounds__(kEnvBlock) k_env_reset(const __grid_constant__ EnvArena a, const __grid_constant__ EnvConst c,
                                                         const __grid_constant__ RngKey key, const uint8_t *mask, mhppo_view obs) {
    const int64_t n = (int64_t)blockIdx.x * kEnvBlock + threadIdx.x;
    if (n >= a.N) return;
    if (mask && !mask[n]) return;
    Rng rng;
    const uint64_t gid = (uint64_t)(key.env_id0 + n);
    rng.env_lo = (uint32_t)gid; rng.env_hi = (uint32_t)(gid >> 32); rng.k0 = key.k0; rng.k1 = key.k1;
    rng.ctr = f2u(a.env_e[n].w);            // the stream cursor survives resets (episodes share one stream)
    reset_and_store<V, MC, MP>(a, c, n, rng, obs);
}

// Crosswalk_hybrid_multi_*.get_state (SC:960-969): the observation of the CURRENT state, no step.  Like the reference it
// runs get_data of every car and pedestrian (all car slots are handed to pedestrian.get_data, as in reset), which also
// folds the current gap into each pedestrian's running-min `delta` (SC:457), so the state is written back.
template <int V, int MC, int MP>
__global__ void __launch_bounds__(kEnvBlock) k_env_observe(const __grid_constant__ EnvArena a, const __grid_constant__ EnvConst c, mhppo_view obs) {
    const int64_t n = (int64_t)blockIdx.x * kEnvBlock + threadIdx.x;
    if (n >= a.N) return;
    EnvR<MC, MP> e;
    load_env<MC, MP>(a, c, n, e);
    if (obs.ptr) { ViewOut out{obs.ptr + n * obs.env_stride, obs.comp_stride}; write_obs<V, MC, MP>(c, e, true, out); }
    else { NullOut nul; write_obs<V, MC, MP>(c, e, true, nul); }
    store_env<MC, MP>(a, c, n, e);
}

// reset_pedestrian / reset_cars of the reference (SC:948-958) for the masked envs; params: one row per env
template <int V, int MC, int MP>
__global__ void __launch_bounds__(kEnvBlock) k_env_inject(const __grid_constant__ EnvArena a, const __grid_constant__ EnvConst c,
                                                          const __grid_constant__ RngKey key, int what, int slot, const float *params,
                                                          const uint8_t *mask) {
    const int64_t n = (int64_t)blockIdx.x * kEnvBlock + threadIdx.x;
    if (n >= a.N) return;
    if (mask && !mask[n]) return;
    EnvR<MC, MP> e;
    load_env<MC, MP>(a, c, n, e);
    const uint64_t gid = (uint64_t)(key.env_id0 + n);
    e.rng.env_lo = (uint32_t)gid; e.rng.env_hi = (uint32_t)(gid >> 32); e.rng.k0 = key.k0; e.rng.k1 = key.k1;
    if (what == 0) inject_pedestrian<V, MC, MP>(c, e, slot, params + n * 9);
    else inject_car(e.car[slot], params + n * 4);
    store_env<MC, MP>(a, c, n, e);
}

// ---- instantiation table -----------------------------------------------------------------------
struct EnvKernelEntry {
    int variant, mc, mp, step_smem;
    void (*step)(EnvArena, EnvConst, RngKey, StepIO);
    void (*reset)(EnvArena, EnvConst, RngKey, const uint8_t *, mhppo_view);
    void (*observe)(EnvArena, EnvConst, mhppo_view);
    void (*inject)(EnvArena, EnvConst, RngKey, int, int, const float *, const uint8_t *);
};

#define MHPPO_ENV_ENTRY(V, MC, MP) { V, MC, MP, (int)sizeof(StepShared<MC>), k_env_step<V, MC, MP>, k_env_reset<V, MC, MP>, k_env_observe<V, MC, MP>, k_env_inject<V, MC, MP> }

// each env_inst_*.cu defines one of these
const EnvKernelEntry *env_table_stop(int *n);
const EnvKernelEntry *env_table_naif(int *n);
const EnvKernelEntry *env_table_coop(int *n);
const EnvKernelEntry *env_table_4cars(int *n);
const EnvKernelEntry *env_table_4cars2(int *n);
const EnvKernelEntry *env_table_scalable(int *n);

}  // namespace mhppo
