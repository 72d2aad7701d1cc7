// ppo_rollout.cuh -- rollout-side kernels of the PPO loop (Env_rollout.iterations_rand, PY:357-516):
// per-(car, pedestrian) feature builders fused with the policy MLP, min-over-pedestrians, sampling and
// log-prob, writing straight into the rollout buffers; and the reverse-scan reward-to-go.
//
// Data layout (all fp32, env index innermost so lane = env is coalesced):
//   obs        [n_obs][N]      the env's component-major observation buffer (scalable class)
//   action_d   [C*P][N] int8   +-1 per (car, ped): which continuous net drives that pair (PY:423)
//   light      [C][N]          +-1 light per car, fixed for the episode (PY:424)
//   sample s = (t*C + car)*N + env,  S = T*C*N
//   obs_c      [13][S]         features of the arg-min pedestrian (PY:448-450)
//   act, logp, rew, rl  [S]    rew / rl are written by the env-step kernel itself through its views
//   choice sample m = car*N + env:  obs_d [D][M], act_d, logp_d, rew_d [M]
#pragma once
#include "mlp.cuh"
#include "philox.cuh"

namespace mhppo {

// Two feature layouts (chosen per rollout config, mhppo_rollout_cfg.legacy_nb_car):
//   scalable  Coop-MH-PPO-scalable.py: C = 2*nb_lines car slots of 7 floats (with `exist`), env row of 4, pedestrians that do
//             not exist are skipped by the continuous head (PY:440), 6 columns per other car in the choice features (the 6th
//             is the OWN exist flag, PY:593), closest_ped_d filters on `exist` (PY:614-627);
//   legacy    the two older notebooks (Coop-MH-PPO.ipynb on `coop`, MH-PPO.ipynb on `naif`; their PPO cells are identical):
//             C = nb_car cars of 6 floats, env row of 3 (lines at index 2), every pedestrian slot is visited (NB2 cell 1,
//             `for p in range(self.env.nb_ped)`), 5 columns per other car (NB2 obs_car_ped_d), closest_ped_d is seeded with
//             pedestrian 0 and has no exist filter (NB2 closest_ped_d).
struct RolloutDims {
    int P, C, L, n_obs, D;      // D = 2 + 6*(C-1) + 8 + 2 (PY:111); legacy: 2 + 5*(C-1) + 8 + 2
    int car_w, env_w, legacy;
    int64_t N;
    uint32_t k0, k1;
    int64_t env_id0;
    HeadCfg head;
};

// flat scalable observation (gym Dict key order car, env, ped; PY:543-554), component-major
struct ObsView {
    const float *o; int64_t N, n; int ped0, env0, car_w, lines_k, legacy;
    __device__ __forceinline__ float car(int i, int k) const { return o[(int64_t)(car_w * i + k) * N + n]; }
    __device__ __forceinline__ float env(int k) const { return o[(int64_t)(env0 + k) * N + n]; }
    __device__ __forceinline__ float ped(int p, int k) const { return o[(int64_t)(ped0 + 9 * p + k) * N + n]; }
};

__device__ __forceinline__ ObsView obs_view(const RolloutDims &d, const float *obs, int64_t n) {
    return ObsView{obs, d.N, n, d.car_w * d.C + d.env_w, d.car_w * d.C, d.car_w, d.env_w - 1, d.legacy};
}

struct LaneGeom { float crossing, end_cross, dist_start, dist_end; };
// Env_rollout.is_in_cross + leave_cross, PY:526-539 (fp32, reference operation order, no FMA)
__device__ __forceinline__ LaneGeom lane_geom(float car_line, float py, float dir, float cl, float lines) {
    const float cross = (cl * 2.0f) / lines;
    const float ls = (-cl) + cross * car_line;
    const float le = (-cl) + cross * (car_line + 1.0f);
    LaneGeom g;
    g.dist_start = (py - ls) * ((dir > 0.f) ? 1.f : 0.f) + (le - py) * ((dir < 0.f) ? 1.f : 0.f);
    g.crossing = (py > ls && py < le) ? 1.f : 0.f;
    const bool neg = (dir == -1.0f);
    g.end_cross = (neg ? (py < ls) : (py > le)) ? 1.f : 0.f;
    g.dist_end = neg ? (ls - py) : (py - le);
    return g;
}

// Env_rollout.obs_car_ped, PY:541-572 -> 13 features into x[0..12]; returns ped exist
__device__ __forceinline__ bool feat_c(const ObsView &v, int i, int p, float *x) {
    const float Vc = v.car(i, 1), dVc = v.car(i, 2), cx = v.car(i, 3), line = v.car(i, 5);
    const float vpx = v.ped(p, 0), vpy = v.ped(p, 1), px = v.ped(p, 2), py = v.ped(p, 3), dl = v.ped(p, 4), ex = v.ped(p, 7),
                dir = v.ped(p, 8);
    const float e0 = v.env(0), e3 = v.env(v.lines_k);
    const LaneGeom g = lane_geom(line, py, dir, e0, e3);
    const bool front = px > cx;
    const float dx = px - cx;
    const float ttc = front ? fminf(10.0f, dx / fmaxf(Vc - vpx, 0.01f)) : 10.0f;
    x[0] = Vc; x[1] = dVc; x[2] = vpy; x[3] = front ? 1.f : 0.f; x[4] = dx; x[5] = dl; x[6] = g.crossing; x[7] = g.end_cross;
    x[8] = g.dist_start; x[9] = g.dist_end; x[10] = ttc; x[11] = e0; x[12] = e3;
    return ex != 0.f;
}

// Env_rollout.obs_car_ped_d, PY:574-611 -> D features (note car_data[6], the OWN exist flag, PY:593)
__device__ __forceinline__ void feat_d(const ObsView &v, int C, int i, int p, float *x) {
    const float Vc = v.car(i, 1), dVc = v.car(i, 2), cx = v.car(i, 3), line = v.car(i, 5), own_exist = v.legacy ? 0.f : v.car(i, 6);
    const float vpy = v.ped(p, 1), px = v.ped(p, 2), py = v.ped(p, 3), dl = v.ped(p, 4), dir = v.ped(p, 8);
    int o = 0;
    x[o++] = Vc; x[o++] = dVc;
    for (int k = 0; k < C; ++k) {
        if (k == i) continue;
        const float x2 = v.car(k, 3);
        x[o++] = v.car(k, 1); x[o++] = (px > x2) ? 1.f : 0.f; x[o++] = px - x2; x[o++] = v.car(k, 4);
        x[o++] = v.car(k, 5) - line;
        if (!v.legacy) x[o++] = own_exist;
    }
    const float e0 = v.env(0), e3 = v.env(v.lines_k);
    const LaneGeom g = lane_geom(line, py, dir, e0, e3);
    x[o++] = vpy; x[o++] = (px > cx) ? 1.f : 0.f; x[o++] = px - cx; x[o++] = dl; x[o++] = g.crossing; x[o++] = g.end_cross;
    x[o++] = g.dist_start; x[o++] = g.dist_end; x[o++] = e0; x[o++] = e3;
}

// policy-noise contract: counter = (index, stream | iteration << 8, env_lo, env_hi); the env's own stream uses
// word 0 (philox.cuh), the continuous head 1, the discrete head 2
__device__ __forceinline__ PhiloxBlock policy_block(const RolloutDims &d, int64_t n, uint32_t index, uint32_t stream_word) {
    const uint64_t gid = (uint64_t)(d.env_id0 + n);
    return philox4x32_10(index, stream_word, (uint32_t)gid, (uint32_t)(gid >> 32), d.k0, d.k1);
}

// ---- episode-start discrete decision (PY:400-428): thread = (env, car) -------------------------------
template <int KP>
__global__ void __launch_bounds__(kFwdBlock, 2) k_choice_act(RolloutDims d, const float *__restrict__ obs, const float *__restrict__ net,
                                                          uint32_t iteration, int8_t *__restrict__ action_d, float *__restrict__ light,
                                                          float *__restrict__ obs_d, float *__restrict__ act_d, float *__restrict__ logp_d) {
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;
    float *rows = smem + ((net_params(KP) + 3) & ~3);
    stage_net<KP>(sw, net);
    __syncthreads();
    const int64_t n = (int64_t)blockIdx.x * kFwdBlock + threadIdx.x;
    const int i = blockIdx.y;
    if (n >= d.N) return;
    float *x = rows + (size_t)threadIdx.x * kRowFwd;      // one in-place row per sample (odd stride: conflict-free)
    const ObsView v = obs_view(d, obs, n);
    const float cx = v.car(i, 3);
    int best = 0; float best_a = 0.f, best_lp = 0.f; double dmin = 1000000.0;    // closest_ped_d, PY:614-627
    const float eps = 1.1920928955078125e-07f;
    for (int p = 0; p < d.P; ++p) {
        for (int k = 0; k < KP; ++k) x[k] = 0.f;
        feat_d(v, d.C, i, p, x);
        const float4 o = mlp_fwd_inplace<KP>(sw, x);
        const float m = fmaxf(o.x, o.y);                                          // softmax over the pair, PY:81-83
        const float e0 = expf(o.x - m), e1 = expf(o.y - m);
        const float p0 = e0 / (e0 + e1), p1 = e1 / (e0 + e1);
        const PhiloxBlock b = policy_block(d, n, (uint32_t)(i * d.P + p), 2u | (iteration << 8));
        const float u = (float)u53(b.w0, b.w1);
        const int a = (u >= p0) ? 1 : 0;                                          // Categorical.sample()
        const float pa = fminf(fmaxf((a ? p1 : p0) / (p0 + p1), eps), 1.0f - eps);  // torch probs_to_logits clamp
        const float lp = logf(pa);
        action_d[(int64_t)(i * d.P + p) * d.N + n] = (int8_t)(2 * a - 1);
        const double dist = (double)(v.ped(p, 2) - cx);
        if (d.legacy ? (p == 0 || dist < dmin) : (dist < dmin && v.ped(p, 7) != 0.f)) { dmin = dist; best = p; }   // legacy: seeded with ped 0
        if (p == 0 || best == p) { best_a = (float)a; best_lp = lp; }
    }
    // the car's light and its choice sample come from the closest pedestrian (PY:416-422)
    for (int k = 0; k < KP; ++k) x[k] = 0.f;
    feat_d(v, d.C, i, best, x);
    const int64_t m_idx = (int64_t)i * d.N + n, M = (int64_t)d.C * d.N;
    for (int k = 0; k < d.D; ++k) obs_d[(int64_t)k * M + m_idx] = x[k];
    light[m_idx] = 2.f * best_a - 1.f;
    act_d[m_idx] = best_a; logp_d[m_idx] = best_lp;
}

// ---- per-step continuous action (PY:434-453): thread = (env, car) -----------------------------------
struct ActIO {
    const float *obs; const int8_t *action_d; const float *light;
    float *actions;                 // [2C][N]: acc rows then light rows, the env-step kernel's input view
    float *obs_c, *act, *logp;      // rollout buffers
    int t, T; uint32_t iteration;
    const uint32_t *iter_dev;       // if set, the iteration number is read from device memory (replayed CUDA graphs)
};

// The CTA owns kFwdBlock envs of one car.  Which net a (car, pedestrian) pair runs through differs per env (PY:441-446), and a
// warp whose lanes read two nets' weights pays two shared-memory wavefronts per load: the kernel was bound by exactly that
// (shared-memory wavefronts 85 % of peak, FMA pipe 33 %, profiles/round2_policy_act_ncu_metrics.json).  So the EXISTING pairs of
// the CTA are first compacted into two lists in shared memory (cross from the front, wait from the back; absent pedestrians
// never enter, PY:440), then thread w evaluates list entry w: every warp reads ONE net (pure broadcast loads) and carries no
// idle lanes; the means go back through shared memory and the owners finish with the min over pedestrians, the arg-min
// features, the sample and its log-prob.
constexpr int kActMaxP = 4;
__global__ void __launch_bounds__(kFwdBlock, 2) k_policy_act(RolloutDims d, const float *__restrict__ net_cross,
                                                          const float *__restrict__ net_wait, ActIO io) {
    constexpr int KP = 16;
    extern __shared__ __align__(16) float smem[];
    constexpr int NP = (net_params(KP) + 3) & ~3;
    float *sw = smem;
    float *rows = smem + 2 * NP;
    float *res = rows + (size_t)kFwdBlock * kRowFwd;                       // [kFwdBlock][kActMaxP] means per (env, pedestrian)
    uint16_t *list = reinterpret_cast<uint16_t *>(res + kFwdBlock * kActMaxP);   // entries: env_local * 4 + pedestrian
    __shared__ int s_cnt[2];
    __shared__ __align__(8) uint64_t wbar;
    const int t = (int)threadIdx.x, lane = t & 31;
    // the two nets (2 x 19 KB) travel as two bulk asynchronous copies (TMA) while the owners classify their pairs; the workers
    // wait on the mbarrier before their first weight load.  (Unaligned caller buffers take the plain staging loop.)
    constexpr uint32_t kNetBytes = (uint32_t)(net_params(KP) * sizeof(float));
    static_assert(kNetBytes % 16 == 0 && (NP * sizeof(float)) % 16 == 0, "bulk copy granularity");
    const bool bulk = (((reinterpret_cast<uintptr_t>(net_cross) | reinterpret_cast<uintptr_t>(net_wait)) & 15u) == 0);
    if (bulk) { if (t == 0) bulk_bar_init(&wbar); }
    else { stage_net<KP>(sw, net_cross); stage_net<KP>(sw + NP, net_wait); }
    if (t < 2) s_cnt[t] = 0;
    __syncthreads();
    if (bulk && t == 0) {
        bulk_expect(&wbar, 2 * kNetBytes);
        bulk_g2s(sw, net_cross, kNetBytes, &wbar);
        bulk_g2s(sw + NP, net_wait, kNetBytes, &wbar);
    }
    const int64_t n0 = (int64_t)blockIdx.x * kFwdBlock;
    const int64_t n = n0 + t;
    const int i = blockIdx.y;
    const bool live = n < d.N;
    const int cap = kFwdBlock * d.P;
    // ---- 1. owners: which pairs exist, which net they take
    uint32_t exm = 0;
    for (int p = 0; p < d.P; ++p) {
        bool ex = false; int sel = 0;
        if (live) {
            ex = d.legacy || io.obs[(int64_t)(d.car_w * d.C + d.env_w + 9 * p + 7) * d.N + n] != 0.f;      // PY:440 (the older notebooks visit every slot)
            sel = (io.action_d[(int64_t)(i * d.P + p) * d.N + n] <= 0) ? 0 : 1;                            // cross | wait, PY:441-446
        }
        exm |= (ex ? 1u : 0u) << p;
        const unsigned mA = __ballot_sync(0xffffffffu, ex && sel == 0), mB = __ballot_sync(0xffffffffu, ex && sel == 1);
        const unsigned lt = (1u << lane) - 1u;
        if (mA) {
            int base = 0;
            const int leader = __ffs((int)mA) - 1;
            if (lane == leader) base = atomicAdd(&s_cnt[0], __popc(mA));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (ex && sel == 0) list[base + __popc(mA & lt)] = (uint16_t)(t * 4 + p);
        }
        if (mB) {
            int base = 0;
            const int leader = __ffs((int)mB) - 1;
            if (lane == leader) base = atomicAdd(&s_cnt[1], __popc(mB));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (ex && sel == 1) list[cap - 1 - (base + __popc(mB & lt))] = (uint16_t)(t * 4 + p);
        }
    }
    __syncthreads();
    if (bulk) bulk_wait(&wbar, 0);
    // ---- 2. workers: one list entry per thread and pass, one net per warp
    {
        const int nA = s_cnt[0], nB = s_cnt[1];
        float *row = rows + (size_t)t * kRowFwd;         // one in-place row per thread (odd stride: conflict-free)
        for (int net = 0; net < 2; ++net) {
            const int cnt = net ? nB : nA;
            for (int w = t; w < cnt; w += kFwdBlock) {
                const int ent = net ? list[cap - 1 - w] : list[w];
                const int e = ent >> 2, p = ent & 3;
                const ObsView v = obs_view(d, io.obs, n0 + e);
                feat_c(v, i, p, row);
                row[13] = row[14] = row[15] = 0.f;
                const float4 o = mlp_fwd_inplace<KP>(sw + net * NP, row);
                res[e * kActMaxP + p] = tanhf(o.x) * d.head.std + d.head.mean;      // head type 1, PY:88-90
            }
        }
    }
    __syncthreads();
    if (!live) return;
    // ---- 3. owners: min over the existing pedestrians, the arg-min pedestrian's features, sample, log-prob
    float mean = d.head.acc_hi;                          // car_b[1,0], PY:436
    int best = 0;                                        // state_c_tensor starts as ped 0's features, PY:437
    for (int p = 0; p < d.P; ++p) {
        if (!((exm >> p) & 1u)) continue;
        const float m = res[t * kActMaxP + p];
        mean = fminf(mean, m);
        if (m == mean) best = p;                         // PY:449-450 (ties: the later pedestrian)
    }
    float st[13];
    const ObsView v = obs_view(d, io.obs, n);
    feat_c(v, i, best, st);
    const uint32_t iteration = io.iter_dev ? *io.iter_dev : io.iteration;
    const PhiloxBlock b = policy_block(d, n, (uint32_t)(io.t * d.C + i), 1u | (iteration << 8));
    const double z = sqrt(-2.0 * log(1.0 - u53(b.w0, b.w1))) * cos(2.0 * 3.141592653589793 * u53(b.w2, b.w3));
    const float a = mean + d.head.sigma * (float)z;               // MultivariateNormal(mean, variance I).sample()
    const float lp = -((a - mean) * (a - mean)) * d.head.inv_2var - d.head.logp_c;   // -(a-mu)^2 / (2 var) - ln(2 pi var)/2
    io.actions[(int64_t)i * d.N + n] = a;
    io.actions[(int64_t)(d.C + i) * d.N + n] = io.light[(int64_t)i * d.N + n];
    const int64_t S = (int64_t)io.T * d.C * d.N, s = ((int64_t)io.t * d.C + i) * d.N + n;
#pragma unroll
    for (int k = 0; k < 13; ++k) io.obs_c[(int64_t)k * S + s] = st[k];
    io.act[s] = a; io.logp[s] = lp;
}

// ---- deterministic evaluation rollout (Env_rollout.iterations, PY:152-252): thread = (env, car) ---------
// Decisions: argmax of the choice net per (car, pedestrian), taken at the first step of an episode (`force`) and again
// whenever the state says ped_traffic != nb_ped (PY:222-224); otherwise the stored decisions stay.
template <int KP>
__global__ void __launch_bounds__(kFwdBlock, 2) k_choice_eval(RolloutDims d, const float *__restrict__ obs, const float *__restrict__ net,
                                                           int force, int8_t *__restrict__ action_d) {
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;
    float *rows = smem + ((net_params(KP) + 3) & ~3);
    stage_net<KP>(sw, net);
    __syncthreads();
    const int64_t n = (int64_t)blockIdx.x * kFwdBlock + threadIdx.x;
    const int i = blockIdx.y;
    if (n >= d.N) return;
    float *x = rows + (size_t)threadIdx.x * kRowFwd;
    const ObsView v = obs_view(d, obs, n);
    if (!force && v.env(1) == (float)d.P) return;                                  // PY:222: state["env"][1] != nb_ped
    for (int p = 0; p < d.P; ++p) {
        for (int k = 0; k < KP; ++k) x[k] = 0.f;
        feat_d(v, d.C, i, p, x);
        const float4 o = mlp_fwd_inplace<KP>(sw, x);
        const float m = fmaxf(o.x, o.y);                                          // softmax over the pair, PY:81-83
        const float e0 = expf(o.x - m), e1 = expf(o.y - m);
        const float p0 = e0 / (e0 + e1), p1 = e1 / (e0 + e1);
        action_d[(int64_t)(i * d.P + p) * d.N + n] = (int8_t)((p1 > p0) ? 1 : -1);   // 2*argmax - 1, first index on a tie (PY:188-192)
    }
}

// Accelerations (PY:195-214): every pedestrian slot is visited, existing or not; a pedestrian that has left the car's lane
// (feature 7) asks for the speed-recovery acceleration clip((speed_limit - v)/dt, acc_lo, acc_hi) instead of a net output;
// the running min is capped by (10 - v)/dt.  The lights handed to env.step are the FIRST C entries of the per-(car,
// pedestrian) decision vector: the reference appends the whole vector and the env reads actions[C + i] (PY:192, 216).
struct EvalIO {
    const float *obs; const int8_t *action_d;
    float *actions;                 // [2C][N]: acc rows then light rows
    float *act;                     // [C][N] record of the accelerations of this step, or null
    float dt, speed_limit, acc_lo, acc_hi;
};
__global__ void __launch_bounds__(kFwdBlock, 2) k_policy_eval(RolloutDims d, const float *__restrict__ net_cross,
                                                           const float *__restrict__ net_wait, EvalIO io) {
    constexpr int KP = 16;
    extern __shared__ __align__(16) float smem[];
    constexpr int NP = (net_params(KP) + 3) & ~3;
    float *sw = smem;
    float *rows = smem + 2 * NP;
    stage_net<KP>(sw, net_cross);
    stage_net<KP>(sw + NP, net_wait);
    __syncthreads();
    const int64_t n = (int64_t)blockIdx.x * kFwdBlock + threadIdx.x;
    const int i = blockIdx.y;
    if (n >= d.N) return;
    float *row = rows + (size_t)threadIdx.x * kRowFwd;
    const ObsView v = obs_view(d, io.obs, n);
    float a = io.acc_hi;                                 // car_b[1,0], PY:197
    float x[13];
    for (int p = 0; p < d.P; ++p) {
        feat_c(v, i, p, x);
        float na;
        if (x[7] != 0.f) {
            na = fmaxf(fminf((io.speed_limit - x[0]) / io.dt, io.acc_hi), io.acc_lo);    // PY:200-201
        } else {
            const int sel = (io.action_d[(int64_t)(i * d.P + p) * d.N + n] <= 0) ? 0 : 1;   // cross | wait, PY:203-206
#pragma unroll
            for (int k = 0; k < 13; ++k) row[k] = x[k];
            row[13] = row[14] = row[15] = 0.f;
            const float4 o = mlp_fwd_inplace<KP>(sw + sel * NP, row);
            na = tanhf(o.x) * d.head.std + d.head.mean;
        }
        a = (na < a) ? na : a;                           // PY:209
        const float cap = (10.0f - x[0]) / io.dt;        // PY:210
        a = (cap < a) ? cap : a;
    }
    io.actions[(int64_t)i * d.N + n] = a;
    io.actions[(int64_t)(d.C + i) * d.N + n] = (float)io.action_d[(int64_t)i * d.N + n];   // flat index i, PY:192/216
    if (io.act) io.act[(int64_t)i * d.N + n] = a;
}

// ---- futur_rewards (PY:658-684) + episodic choice reward (PY:461): thread = (car, env) ---------------
__global__ void __launch_bounds__(256) k_returns(const float *__restrict__ rew, const float *__restrict__ rl, int T, int64_t CN,
                                                 double gamma, float *__restrict__ rtg, float *__restrict__ rew_d) {
    const int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (m >= CN) return;
    double acc = 0.0;
    float worst = 0.f;
    for (int t = T - 1; t >= 0; --t) {
        const int64_t s = (int64_t)t * CN + m;
        acc = (double)rew[s] + gamma * acc;
        rtg[s] = (float)acc;
        worst = fminf(worst, rl[s]);
    }
    rew_d[m] = worst;
}

}  // namespace mhppo
