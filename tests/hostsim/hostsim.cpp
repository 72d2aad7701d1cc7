// hostsim.cpp -- TEST-ONLY host build of the kernel's per-env logic (mh-ppo_b200/csrc/env_core.cuh,
// env_state.cuh) so the CUDA code path can be unit-tested against the golden fixtures and the oracle on a
// machine without a GPU.  It is built by tests/hostsim/build.py into tests/hostsim/, is never
// loaded by the product package, and is NOT a CPU fallback: mh-ppo_b200 fails loudly without CUDA.
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#define _GNU_SOURCE 1
#include "../../mh-ppo_b200/csrc/env_step.cuh"

using namespace mhppo;

typedef mhppo_view View;

struct Sim {
    int variant, mc, mp;
    EnvConst c; EnvArena a; uint32_t k0, k1; int64_t env_id0;
};

template <int V, int MC, int MP>
static void sim_reset(Sim *s, const uint8_t *mask, mhppo_view obs) {
    for (int64_t n = 0; n < s->a.N; ++n) {
        if (mask && !mask[n]) continue;
        Rng rng;
        const uint64_t gid = (uint64_t)(s->env_id0 + n);
        rng.env_lo = (uint32_t)gid; rng.env_hi = (uint32_t)(gid >> 32); rng.k0 = s->k0; rng.k1 = s->k1;
        rng.ctr = f2u(s->a.env_e[n].w);
        reset_and_store<V, MC, MP>(s->a, s->c, n, rng, obs);
    }
}

template <int V, int MC, int MP>
static void sim_step(Sim *s, mhppo_view actions, mhppo_view obs, mhppo_view rewards, mhppo_view reward_light, uint8_t *done,
                     int autoreset, mhppo_view term_obs) {
    StepIO io;
    io.actions = actions; io.obs = obs; io.rewards = rewards; io.reward_light = reward_light; io.term_obs = term_obs;
    io.done = done; io.autoreset = autoreset; io.n_begin = 0; io.n_end = s->a.N;
    RngKey key; key.k0 = s->k0; key.k1 = s->k1; key.env_id0 = s->env_id0;
    CarSlots<MC, 1> cars;
    for (int64_t n = 0; n < s->a.N; ++n) env_step_thread<V, MC, MP, 1>(s->a, s->c, key, io, n, cars, 0);   // the kernel's thread body
}

template <int V, int MC, int MP>
static void sim_inject(Sim *s, int what, int slot, const float *params, const uint8_t *mask) {
    for (int64_t n = 0; n < s->a.N; ++n) {                                    // body of k_env_inject
        if (mask && !mask[n]) continue;
        EnvR<MC, MP> e;
        load_env<MC, MP>(s->a, s->c, n, e);
        const uint64_t gid = (uint64_t)(s->env_id0 + n);
        e.rng.env_lo = (uint32_t)gid; e.rng.env_hi = (uint32_t)(gid >> 32); e.rng.k0 = s->k0; e.rng.k1 = s->k1;
        if (what == 0) inject_pedestrian<V, MC, MP>(s->c, e, slot, params + n * 9);
        else inject_car(e.car[slot], params + n * 4);
        store_env<MC, MP>(s->a, s->c, n, e);
    }
}
template <int V, int MC, int MP>
static void sim_observe(Sim *s, mhppo_view obs) {
    for (int64_t n = 0; n < s->a.N; ++n) {                                    // body of k_env_observe
        EnvR<MC, MP> e;
        load_env<MC, MP>(s->a, s->c, n, e);
        ViewOut out{obs.ptr + n * obs.env_stride, obs.comp_stride};
        write_obs<V, MC, MP>(s->c, e, true, out);
        store_env<MC, MP>(s->a, s->c, n, e);
    }
}

#define DISPATCH(FN, ...)                                                                          \
    do {                                                                                           \
        const int key = s->variant * 10000 + s->mc * 100 + s->mp;                                  \
        switch (key) {                                                                             \
            case 50403: FN<V_SCAL, 4, 3>(__VA_ARGS__); break;                                      \
            case 50201: FN<V_SCAL, 2, 1>(__VA_ARGS__); break;                                      \
            case 50804: FN<V_SCAL, 8, 4>(__VA_ARGS__); break;                                      \
            case 20804: FN<V_COOP, 8, 4>(__VA_ARGS__); break;                                      \
            case 20201: FN<V_COOP, 2, 1>(__VA_ARGS__); break;                                      \
            case 804: FN<V_STOP, 8, 4>(__VA_ARGS__); break;                                        \
            case 10804: FN<V_NAIF, 8, 4>(__VA_ARGS__); break;                                      \
            case 30202: FN<V_4CARS, 2, 2>(__VA_ARGS__); break;                                     \
            case 30402: FN<V_4CARS, 4, 2>(__VA_ARGS__); break;                                     \
            case 30604: FN<V_4CARS, 6, 4>(__VA_ARGS__); break;                                     \
            case 40202: FN<V_4CARS2, 2, 2>(__VA_ARGS__); break;                                    \
            case 40402: FN<V_4CARS2, 4, 2>(__VA_ARGS__); break;                                    \
            case 40604: FN<V_4CARS2, 6, 4>(__VA_ARGS__); break;                                    \
            default: return -5;                                                                    \
        }                                                                                          \
    } while (0)

extern "C" {

int hs_create(int variant, int mc, int mp, const EnvConst *c, int64_t N, uint64_t seed, int64_t env_id0, void **out) {
    Sim *s = (Sim *)calloc(1, sizeof(Sim));
    s->variant = variant; s->mc = mc; s->mp = mp; s->c = *c; env_const_finish(s->c); s->k0 = (uint32_t)seed; s->k1 = (uint32_t)(seed >> 32);
    s->env_id0 = env_id0; s->a.N = N;
    s->a.car_a = (float4 *)calloc((size_t)N * mc, 16); s->a.car_b = (float4 *)calloc((size_t)N * mc, 16);
    s->a.ped_a = (float4 *)calloc((size_t)N * mp, 16); s->a.ped_b = (float4 *)calloc((size_t)N * mp, 16);
    s->a.ped_c = (float4 *)calloc((size_t)N * mp, 16); s->a.env_e = (float4 *)calloc((size_t)N, 16);
    *out = s;
    return 0;
}
void hs_destroy(void *h) {
    Sim *s = (Sim *)h;
    free(s->a.car_a); free(s->a.car_b); free(s->a.ped_a); free(s->a.ped_b); free(s->a.ped_c); free(s->a.env_e); free(s);
}
int hs_reset(void *h, const uint8_t *mask, View obs) { Sim *s = (Sim *)h; DISPATCH(sim_reset, s, mask, obs); return 0; }
int hs_step(void *h, View actions, View obs, View rewards, View reward_light, uint8_t *done, int autoreset, View term_obs) {
    Sim *s = (Sim *)h; DISPATCH(sim_step, s, actions, obs, rewards, reward_light, done, autoreset, term_obs); return 0;
}
void hs_export(void *h, float *car_f, int32_t *car_i, float *ped_f, int32_t *ped_i, double *env_f, int64_t *env_i) {
    Sim *s = (Sim *)h; DumpPtrs d{car_f, car_i, ped_f, ped_i, env_f, env_i};
    for (int64_t n = 0; n < s->a.N; ++n) export_one(s->a, s->c, n, d);
}
void hs_import(void *h, float *car_f, int32_t *car_i, float *ped_f, int32_t *ped_i, double *env_f, int64_t *env_i) {
    Sim *s = (Sim *)h; DumpPtrs d{car_f, car_i, ped_f, ped_i, env_f, env_i};
    for (int64_t n = 0; n < s->a.N; ++n) import_one(s->a, s->c, n, d);
}
int hs_inject(void *h, int what, int slot, const float *params, const uint8_t *mask) {
    Sim *s = (Sim *)h; DISPATCH(sim_inject, s, what, slot, params, mask); return 0;
}
int hs_observe(void *h, View obs) { Sim *s = (Sim *)h; DISPATCH(sim_observe, s, obs); return 0; }
int hs_sizeof_envconst(void) { return (int)sizeof(EnvConst); }
}
