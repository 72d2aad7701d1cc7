// mhppo_api.cu -- host side of the C ABI declared in include/mhppo.h (env part).
//
// Owns the HBM state arena of a vectorised env, picks the kernel instantiation for the requested
// (variant, car slots, pedestrians) and enqueues the kernels on the caller's stream.  No torch
// types, no allocation after create, no hidden synchronisation (except the *_host entry points).
#include <atomic>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "env_kernels.cuh"

namespace mhppo {

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const std::string &msg) { g_err = msg; return code; }
int api_fail(int code, const std::string &msg) { return fail(code, msg); }   // shared with ppo_api.cu
void rollout_forget_env(void *env);                                          // ppo_api.cu: drops a CUDA graph captured on this env
void api_count_launch() { g_launches.fetch_add(1); }
static int cuda_fail(cudaError_t e, const char *what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return MHPPO_ECUDA;
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);        \
    } while (0)

constexpr int kHostStreams = 3;          // slices of the host-buffer step rotate over these streams
constexpr int64_t kHostSliceMin = 16384;  // envs per slice at least (a slice must still fill the GPU)
constexpr int kHostSlicesMax = 8;

struct EnvHandle {
    mhppo_env_cfg cfg;
    EnvConst c;
    EnvArena a;
    RngKey key;
    const EnvKernelEntry *k;
    void *arena_base;
    size_t arena_bytes;
    // device staging for the *_host entry points (row-major, reference layout) and the streams / events that
    // pipeline slices of the env range through H2D -> kernel -> D2H
    float *d_act, *d_obs, *d_rew, *d_rl; uint8_t *d_done;
    cudaStream_t hs[kHostStreams];
    cudaEvent_t ev_in, ev_out[kHostStreams];
};

__global__ void __launch_bounds__(kEnvBlock) k_env_export(EnvArena a, EnvConst c, DumpPtrs d) {
    const int64_t n = (int64_t)blockIdx.x * kEnvBlock + threadIdx.x;
    if (n < a.N) export_one(a, c, n, d);
}
__global__ void __launch_bounds__(kEnvBlock) k_env_import(EnvArena a, EnvConst c, DumpPtrs d) {
    const int64_t n = (int64_t)blockIdx.x * kEnvBlock + threadIdx.x;
    if (n < a.N) import_one(a, c, n, d);
}

static const EnvKernelEntry *find_kernel(int variant, int C, int P) {
    int n = 0;
    const EnvKernelEntry *t = nullptr;
    switch (variant) {
        case V_STOP: t = env_table_stop(&n); break;
        case V_NAIF: t = env_table_naif(&n); break;
        case V_COOP: t = env_table_coop(&n); break;
        case V_4CARS: t = env_table_4cars(&n); break;
        case V_4CARS2: t = env_table_4cars2(&n); break;
        case V_SCAL: t = env_table_scalable(&n); break;
        default: return nullptr;
    }
    const bool exact_c = (variant == V_4CARS || variant == V_4CARS2);  // follower registers sit at MC/2
    const EnvKernelEntry *best = nullptr;
    for (int i = 0; i < n; ++i) {
        const EnvKernelEntry &e = t[i];
        if (e.mp < P) continue;
        if (exact_c ? (e.mc != C) : (e.mc < C)) continue;
        if (!best || e.mc * 16 + e.mp < best->mc * 16 + best->mp) best = &e;
    }
    return best;
}

static unsigned grid_for(int64_t n) { return (unsigned)((n + kEnvBlock - 1) / kEnvBlock); }


// Self-test of the gap-acceptance predicate (env_step.cuh gap_eval): the filtered decision, the level that took it and the
// exact fp64 decision with its critical gap, one thread per sample.
__global__ void k_gap_selftest(int64_t n, const double *dx, const double *vden, const double *light, const double *size,
                               const double *v0y, const int32_t *gender, const int32_t *age, const uint32_t *ctr,
                               uint32_t env_lo, uint32_t env_hi, uint32_t k0, uint32_t k1,
                               uint8_t *fast, uint8_t *level, uint8_t *exact, double *cg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lv = 0;
    fast[i] = gap_eval(dx[i], vden[i], light[i], size[i], v0y[i], gender[i], age[i], ctr[i], env_lo, env_hi, k0, k1, &lv) ? 1 : 0;
    level[i] = (uint8_t)lv;
    const PhiloxBlock b = philox4x32_10(ctr[i], 0u, env_lo, env_hi, k0, k1);
    const double g = cg_exact(v0y[i], gender[i], age[i], size[i], u53(b.w0, b.w1), u53(b.w2, b.w3));
    exact[i] = ((fabs(dx[i] / vden[i]) + light[i]) < g) ? 1 : 0;
    if (cg) cg[i] = g;
}
}  // namespace mhppo

using namespace mhppo;

extern "C" {

int mhppo_abi_version(void) { return MHPPO_ABI_VERSION; }
const char *mhppo_last_error(void) { return g_err.c_str(); }
int64_t mhppo_launch_count(void) { return g_launches.load(); }

int mhppo_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int mhppo_env_create(const mhppo_env_cfg *cfg, void **handle) {
    if (!cfg || !handle) return fail(MHPPO_EINVAL, "null argument");
    *handle = nullptr;
    if (cfg->variant < 0 || cfg->variant > 5) return fail(MHPPO_EINVAL, "unknown env variant");
    if (cfg->nb_car < 1 || cfg->nb_ped < 1 || cfg->nb_lines < 1 || cfg->n_envs < 1)
        return fail(MHPPO_EINVAL, "nb_car, nb_ped, nb_lines and n_envs must be >= 1");
    if (cfg->nb_lines > 29) return fail(MHPPO_EINVAL, "nb_lines > 29 does not fit the packed line_pos field");
    if (cfg->max_episode < 1 || cfg->max_episode > 250)
        return fail(MHPPO_EINVAL, "max_episode must be in [1, 250] (8-bit dt counters)");
    if (!(cfg->dt > 0.0)) return fail(MHPPO_EINVAL, "dt must be > 0");
    const int v = cfg->variant;
    const bool four = (v == V_4CARS || v == V_4CARS2);
    const int C = (v == V_SCAL) ? 2 * cfg->nb_lines : (four ? 2 * cfg->nb_car : cfg->nb_car);
    if (v == V_SCAL && cfg->nb_car > C) return fail(MHPPO_EINVAL, "scalable env needs nb_car <= 2*nb_lines (random.sample, SC:903)");
    const EnvKernelEntry *k = find_kernel(v, C, cfg->nb_ped);
    if (!k) {
        char b[160];
        snprintf(b, sizeof b, "no sm_100a kernel instantiation for variant %d with %d car slots and %d pedestrians", v, C, cfg->nb_ped);
        return fail(MHPPO_EUNSUPPORTED, b);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MHPPO_ENODEV, "no CUDA device: mhppo_b200 has no CPU fallback");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(MHPPO_EINVAL, "device ordinal out of range");
    CK(cudaSetDevice(cfg->device));
    CK(cudaFuncSetAttribute((const void *)k->step, cudaFuncAttributeMaxDynamicSharedMemorySize, k->step_smem));
    // the step kernel streams its global data once (no L1 reuse) and keeps the cars in shared memory
    CK(cudaFuncSetAttribute((const void *)k->step, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));

    EnvHandle *h = new (std::nothrow) EnvHandle();
    if (!h) return fail(MHPPO_ENOMEM, "host allocation failed");
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg; h->k = k;
    EnvConst &c = h->c;
    c.nC = C; c.nP = cfg->nb_ped; c.L = cfg->nb_lines; c.nb_car = cfg->nb_car;
    c.nlead = four ? cfg->nb_car : C;
    c.nA = (v == V_SCAL) ? 4 * cfg->nb_lines : ((v == V_4CARS2) ? 4 * cfg->nb_car : 2 * cfg->nb_car);
    c.nobs = ((v == V_SCAL) ? 7 : 6) * C + ((v == V_SCAL) ? 4 : 3) + 9 * cfg->nb_ped;
    {   // done = time >= (max_episode-1)*dt with time the fp64 running sum of dt (SC:874-875, 945)
        double t = 0.0; const double lim = (cfg->max_episode - 1) * cfg->dt; int kk = 0;
        while (!(t >= lim) && kk < 100000) { t = t + cfg->dt; ++kk; }
        c.done_idx = kk;
    }
    c.sin_model = cfg->sin_model; c.dt = cfg->dt; c.acc_lo = cfg->car_b[0]; c.acc_hi = cfg->car_b[2];
    for (int i = 0; i < 8; ++i) c.pb[i] = cfg->ped_b[i];
    c.cross_lo = cfg->cross_b[0]; c.cross_hi = cfg->cross_b[1];
    env_const_finish(c);
    h->key.k0 = (uint32_t)cfg->seed; h->key.k1 = (uint32_t)(cfg->seed >> 32); h->key.env_id0 = cfg->env_id0;

    const int64_t N = cfg->n_envs;
    const size_t per_env = (size_t)(2 * C + 3 * cfg->nb_ped + 1) * sizeof(float4);
    h->arena_bytes = per_env * (size_t)N;
    cudaError_t e = cudaMalloc(&h->arena_base, h->arena_bytes);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaMalloc(state arena)"); }
    e = cudaMemset(h->arena_base, 0, h->arena_bytes);
    if (e != cudaSuccess) { cudaFree(h->arena_base); delete h; return cuda_fail(e, "cudaMemset(state arena)"); }
    float4 *p = (float4 *)h->arena_base;
    h->a.N = N;
    h->a.car_a = p; p += (size_t)C * N;
    h->a.car_b = p; p += (size_t)C * N;
    h->a.ped_a = p; p += (size_t)cfg->nb_ped * N;
    h->a.ped_b = p; p += (size_t)cfg->nb_ped * N;
    h->a.ped_c = p; p += (size_t)cfg->nb_ped * N;
    h->a.env_e = p;
    // device staging for the host-buffer entry points
    const size_t nb = (size_t)N * sizeof(float);
    e = cudaMalloc(&h->d_act, nb * c.nA);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_obs, nb * c.nobs);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_rew, nb * c.nlead);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_rl, nb * c.nlead);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_done, (size_t)N);
    if (e != cudaSuccess) { mhppo_env_destroy(h); return cuda_fail(e, "cudaMalloc(host-API staging)"); }
    for (int i = 0; i < kHostStreams && e == cudaSuccess; ++i) {
        e = cudaStreamCreateWithFlags(&h->hs[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming);
    if (e != cudaSuccess) { mhppo_env_destroy(h); return cuda_fail(e, "cudaStreamCreate(host-API pipeline)"); }
    *handle = h;
    return MHPPO_OK;
}

int mhppo_env_destroy(void *handle) {
    EnvHandle *h = (EnvHandle *)handle;
    if (!h) return MHPPO_OK;
    rollout_forget_env(h);
    cudaFree(h->arena_base); cudaFree(h->d_act); cudaFree(h->d_obs); cudaFree(h->d_rew); cudaFree(h->d_rl);
    cudaFree(h->d_done);
    for (int i = 0; i < kHostStreams; ++i) {
        if (h->hs[i]) cudaStreamDestroy(h->hs[i]);
        if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
    }
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    delete h;
    return MHPPO_OK;
}

int mhppo_env_get_dims(void *handle, mhppo_env_dims *d) {
    EnvHandle *h = (EnvHandle *)handle;
    if (!h || !d) return fail(MHPPO_EINVAL, "null argument");
    memset(d, 0, sizeof(*d));
    d->n_slots = h->c.nC; d->n_lead = h->c.nlead; d->n_action = h->c.nA; d->n_obs = h->c.nobs;
    d->n_ped = h->c.nP; d->done_step = h->c.done_idx;
    return MHPPO_OK;
}

int64_t mhppo_env_state_bytes_per_env(void *handle) {
    EnvHandle *h = (EnvHandle *)handle;
    return h ? (int64_t)(h->arena_bytes / (size_t)h->a.N) : 0;
}

int mhppo_env_reset(void *handle, const uint8_t *mask_dev, mhppo_view obs_dev, void *stream) {
    EnvHandle *h = (EnvHandle *)handle;
    if (!h) return fail(MHPPO_EINVAL, "null handle");
    auto fn = h->k->reset;
    fn<<<grid_for(h->a.N), kEnvBlock, 0, (cudaStream_t)stream>>>(h->a, h->c, h->key, mask_dev, obs_dev);
    g_launches.fetch_add(1);
    CK(cudaGetLastError());
    return MHPPO_OK;
}

static int inject(void *handle, int what, int slot, int n_slots, const float *params, const uint8_t *mask, void *stream) {
    EnvHandle *h = (EnvHandle *)handle;
    if (!h || !params) return fail(MHPPO_EINVAL, "null argument");
    if (slot < 0 || slot >= n_slots) return fail(MHPPO_EINVAL, "slot index out of range");
    auto fn = h->k->inject;
    fn<<<grid_for(h->a.N), kEnvBlock, 0, (cudaStream_t)stream>>>(h->a, h->c, h->key, what, slot, params, mask);
    g_launches.fetch_add(1);
    CK(cudaGetLastError());
    return MHPPO_OK;
}
int mhppo_env_reset_pedestrian(void *handle, int32_t num_ped, const float *params_dev, const uint8_t *mask_dev, void *stream) {
    return inject(handle, 0, num_ped, handle ? ((EnvHandle *)handle)->c.nP : 0, params_dev, mask_dev, stream);
}
int mhppo_env_reset_cars(void *handle, int32_t num_car, const float *params_dev, const uint8_t *mask_dev, void *stream) {
    return inject(handle, 1, num_car, handle ? ((EnvHandle *)handle)->c.nC : 0, params_dev, mask_dev, stream);
}

int mhppo_env_observe(void *handle, mhppo_view obs_dev, void *stream) {
    EnvHandle *h = (EnvHandle *)handle;
    if (!h) return fail(MHPPO_EINVAL, "null handle");
    auto fn = h->k->observe;
    fn<<<grid_for(h->a.N), kEnvBlock, 0, (cudaStream_t)stream>>>(h->a, h->c, obs_dev);
    g_launches.fetch_add(1);
    CK(cudaGetLastError());
    return MHPPO_OK;
}

int mhppo_env_step(void *handle, mhppo_view actions_dev, mhppo_view obs_dev, mhppo_view rewards_dev,
                   mhppo_view reward_light_dev, uint8_t *done_dev, int autoreset, mhppo_view term_obs_dev,
                   void *stream) {
    EnvHandle *h = (EnvHandle *)handle;
    if (!h) return fail(MHPPO_EINVAL, "null handle");
    if (!actions_dev.ptr) return fail(MHPPO_EINVAL, "actions are required");
    // the step kernel keeps the observation component stride in one 32-bit register (env_step.cuh)
    if ((obs_dev.ptr && (obs_dev.comp_stride < 0 || obs_dev.comp_stride > INT32_MAX)) ||
        (term_obs_dev.ptr && (term_obs_dev.comp_stride < 0 || term_obs_dev.comp_stride > INT32_MAX)))
        return fail(MHPPO_EINVAL, "observation comp_stride must be in [0, 2^31)");
    StepIO io;
    io.actions = actions_dev; io.obs = obs_dev; io.rewards = rewards_dev; io.reward_light = reward_light_dev;
    io.term_obs = term_obs_dev; io.done = done_dev; io.autoreset = autoreset; io.n_begin = 0; io.n_end = h->a.N;
    auto fn = h->k->step;
    fn<<<grid_for(h->a.N), kEnvBlock, h->k->step_smem, (cudaStream_t)stream>>>(h->a, h->c, h->key, io);
    g_launches.fetch_add(1);
    CK(cudaGetLastError());
    return MHPPO_OK;
}

int mhppo_env_reset_host(void *handle, float *obs_host, void *stream) {
    EnvHandle *h = (EnvHandle *)handle;
    if (!h) return fail(MHPPO_EINVAL, "null handle");
    cudaStream_t s = (cudaStream_t)stream;
    mhppo_view obs = { h->d_obs, h->c.nobs, 1 };
    int rc = mhppo_env_reset(handle, nullptr, obs, stream);
    if (rc) return rc;
    if (obs_host) CK(cudaMemcpyAsync(obs_host, h->d_obs, sizeof(float) * h->c.nobs * h->a.N, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return MHPPO_OK;
}

int mhppo_env_step_host(void *handle, const float *actions_host, float *obs_host, float *rewards_host,
                        float *reward_light_host, uint8_t *done_host, int autoreset, void *stream) {
    EnvHandle *h = (EnvHandle *)handle;
    if (!h || !actions_host) return fail(MHPPO_EINVAL, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t N = h->a.N; const EnvConst &c = h->c;
    // The env range is cut into slices; slice k runs H2D(actions) -> kernel -> D2H(results) on stream k % 3, so the
    // D2H of one slice (the long pole: n_obs + 2 n_lead floats per env over PCIe) overlaps the H2D and the kernel of the
    // next ones (the two copy engines and the SMs work at the same time).  Rows of a slice are contiguous in every
    // host buffer: one cudaMemcpyAsync per buffer and slice.
    int ns = (int)(N / kHostSliceMin);
    ns = ns < 1 ? 1 : (ns > kHostSlicesMax ? kHostSlicesMax : ns);
    const int64_t per = ((N + ns - 1) / ns + kEnvBlock - 1) / kEnvBlock * kEnvBlock;
    CK(cudaEventRecord(h->ev_in, s));                                // work already queued on the caller's stream comes first
    const int used = ns < kHostStreams ? ns : kHostStreams;
    for (int i = 0; i < used; ++i) CK(cudaStreamWaitEvent(h->hs[i], h->ev_in, 0));
    StepIO io;
    const mhppo_view none = { nullptr, 0, 0 };
    io.actions = mhppo_view{ h->d_act, c.nA, 1 };
    io.obs = obs_host ? mhppo_view{ h->d_obs, c.nobs, 1 } : none;
    io.rewards = rewards_host ? mhppo_view{ h->d_rew, c.nlead, 1 } : none;
    io.reward_light = reward_light_host ? mhppo_view{ h->d_rl, c.nlead, 1 } : none;
    io.term_obs = none; io.done = done_host ? h->d_done : nullptr; io.autoreset = autoreset;
    auto fn = h->k->step;
    for (int k = 0; k < ns; ++k) {
        const int64_t n0 = (int64_t)k * per, n1 = (n0 + per < N) ? n0 + per : N;
        if (n0 >= n1) break;
        cudaStream_t q = h->hs[k % kHostStreams];
        const size_t cnt = (size_t)(n1 - n0);
        CK(cudaMemcpyAsync(h->d_act + n0 * c.nA, actions_host + n0 * c.nA, sizeof(float) * c.nA * cnt, cudaMemcpyHostToDevice, q));
        io.n_begin = n0; io.n_end = n1;
        fn<<<grid_for(n1 - n0), kEnvBlock, h->k->step_smem, q>>>(h->a, h->c, h->key, io);
        g_launches.fetch_add(1);
        CK(cudaGetLastError());
        if (obs_host) CK(cudaMemcpyAsync(obs_host + n0 * c.nobs, h->d_obs + n0 * c.nobs, sizeof(float) * c.nobs * cnt, cudaMemcpyDeviceToHost, q));
        if (rewards_host) CK(cudaMemcpyAsync(rewards_host + n0 * c.nlead, h->d_rew + n0 * c.nlead, sizeof(float) * c.nlead * cnt, cudaMemcpyDeviceToHost, q));
        if (reward_light_host) CK(cudaMemcpyAsync(reward_light_host + n0 * c.nlead, h->d_rl + n0 * c.nlead, sizeof(float) * c.nlead * cnt, cudaMemcpyDeviceToHost, q));
        if (done_host) CK(cudaMemcpyAsync(done_host + n0, h->d_done + n0, cnt, cudaMemcpyDeviceToHost, q));
    }
    for (int i = 0; i < used; ++i) {                                  // the caller's stream continues after every slice
        CK(cudaEventRecord(h->ev_out[i], h->hs[i]));
        CK(cudaStreamWaitEvent(s, h->ev_out[i], 0));
    }
    CK(cudaStreamSynchronize(s));
    return MHPPO_OK;
}

int mhppo_env_export_state(void *handle, float *car_f, int32_t *car_i, float *ped_f, int32_t *ped_i, double *env_f,
                           int64_t *env_i, void *stream) {
    EnvHandle *h = (EnvHandle *)handle;
    if (!h || !car_f || !car_i || !ped_f || !ped_i || !env_f || !env_i) return fail(MHPPO_EINVAL, "null argument");
    DumpPtrs d{car_f, car_i, ped_f, ped_i, env_f, env_i};
    k_env_export<<<grid_for(h->a.N), kEnvBlock, 0, (cudaStream_t)stream>>>(h->a, h->c, d);
    g_launches.fetch_add(1);
    CK(cudaGetLastError());
    return MHPPO_OK;
}

int mhppo_env_import_state(void *handle, const float *car_f, const int32_t *car_i, const float *ped_f,
                           const int32_t *ped_i, const double *env_f, const int64_t *env_i, void *stream) {
    EnvHandle *h = (EnvHandle *)handle;
    if (!h || !car_f || !car_i || !ped_f || !ped_i || !env_f || !env_i) return fail(MHPPO_EINVAL, "null argument");
    DumpPtrs d{(float *)car_f, (int32_t *)car_i, (float *)ped_f, (int32_t *)ped_i, (double *)env_f, (int64_t *)env_i};
    k_env_import<<<grid_for(h->a.N), kEnvBlock, 0, (cudaStream_t)stream>>>(h->a, h->c, d);
    g_launches.fetch_add(1);
    CK(cudaGetLastError());
    return MHPPO_OK;
}

int mhppo_gap_selftest(int64_t n, const double *dx_dev, const double *vden_dev, const double *light_dev, const double *size_dev,
                       const double *v0y_dev, const int32_t *gender_dev, const int32_t *age_dev, const uint32_t *ctr_dev,
                       uint64_t env_id, uint32_t key0, uint32_t key1, uint8_t *fast_dev, uint8_t *level_dev, uint8_t *exact_dev,
                       double *cg_dev, void *stream) {
    if (n <= 0 || !dx_dev || !vden_dev || !light_dev || !size_dev || !v0y_dev || !gender_dev || !age_dev || !ctr_dev || !fast_dev ||
        !level_dev || !exact_dev)
        return fail(MHPPO_EINVAL, "null argument");
    k_gap_selftest<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, dx_dev, vden_dev, light_dev, size_dev, v0y_dev,
                                                                              gender_dev, age_dev, ctr_dev, (uint32_t)env_id,
                                                                              (uint32_t)(env_id >> 32), key0, key1, fast_dev,
                                                                              level_dev, exact_dev, cg_dev);
    CK(cudaGetLastError());
    return MHPPO_OK;
}

}  // extern "C"
