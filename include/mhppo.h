/* mhppo.h -- C ABI of libmhppo_b200.so: the B200 (sm_100a) drop-in for MH-PPO's hot path.
 *
 * The reference (BrunoudA/MH-PPO) has no FFI: its seam is the gym-0.26 `Env` Python API of the six
 * `Environments/Env_hybrid_multi_*.py` classes plus the torch calls of `Env_rollout` / `Algo_PPO`
 * in `Coop-MH-PPO-scalable.py`.  Each entry point below names the reference interface it replaces
 * (file:line, abbreviations: SC = Environments/Env_hybrid_multi_coop_scalable.py, CO = .._coop.py,
 * ST = .._stop.py, NA = .._naif.py, C4 = .._coop_4cars.py, C42 = .._coop_4cars2.py,
 * PY = Coop-MH-PPO-scalable.py).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *  - plain C types only; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - every pointer named *_dev is device memory owned by the caller (e.g. torch tensor.data_ptr());
 *    the library never frees or retains it past the call; *_host pointers are host memory;
 *  - every call returns 0 on success or a negative MHPPO_E* code; mhppo_last_error() gives the text
 *    for the calling thread; nothing throws, nothing allocates after *_create;
 *  - all work is enqueued on `stream`, no hidden synchronisation (except the *_host entry points,
 *    which synchronise `stream` before returning because they fill host buffers);
 *  - one handle = one GPU = one host thread at a time.
 *  - there is NO CPU fallback: without a CUDA device every call fails with MHPPO_ENODEV.
 */
#ifndef MHPPO_H
#define MHPPO_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MHPPO_ABI_VERSION 1

enum { MHPPO_OK = 0, MHPPO_EINVAL = -1, MHPPO_ENODEV = -2, MHPPO_ECUDA = -3, MHPPO_ENOMEM = -4,
       MHPPO_EUNSUPPORTED = -5 };

/* env classes of the reference, Environments/__init__.py:3-41 */
enum { MHPPO_ENV_STOP = 0, MHPPO_ENV_NAIF = 1, MHPPO_ENV_COOP = 2, MHPPO_ENV_COOP_4CARS = 3,
       MHPPO_ENV_COOP_4CARS2 = 4, MHPPO_ENV_COOP_SCALABLE = 5 };

/* Constructor arguments of the env classes (SC:680: car_b, ped_b, cross_b, nb_car, nb_ped, nb_lines,
 * dt, max_episode, simulation) plus what a vectorised env needs (n_envs, RNG stream, device). */
typedef struct mhppo_env_cfg {
    int32_t variant;      /* MHPPO_ENV_* */
    int32_t nb_car, nb_ped, nb_lines;
    int32_t max_episode;  /* 80 in the drivers (PY:1026) */
    int32_t sin_model;    /* simulation == "sin" */
    int32_t device;       /* CUDA device ordinal */
    int32_t reserved;
    double dt;            /* 0.3 */
    double car_b[4];      /* row-major 2x2 */
    double ped_b[8];      /* row-major 2x4 */
    double cross_b[2];
    uint64_t seed;        /* Philox key (RNG contract: DESIGN.md "RNG") */
    int64_t n_envs;       /* envs held by this handle */
    int64_t env_id0;      /* global id of env 0: rank r of R passes r*n_envs so shards are disjoint streams */
} mhppo_env_cfg;

typedef struct mhppo_env_dims {
    int32_t n_slots;      /* car slots in state: nb_car | 2*nb_car (4cars*) | 2*nb_lines (scalable) */
    int32_t n_lead;       /* cars with a reward / reward_light entry */
    int32_t n_action;     /* action vector length, [acc.., light..] as in SC:798-802 / C42:809-811 */
    int32_t n_obs;        /* flat observation: keys in gym-0.26 Dict order car,(car_follow),env,ped */
    int32_t n_ped;
    int32_t done_step;    /* step index whose step() returns done (79 for dt=.3, max_episode=80) */
    int32_t reserved[2];
} mhppo_env_dims;

/* A 2-D fp32 view [n_envs, width] with element strides, so callers may keep either the gym layout
 * (env_stride = width, comp_stride = 1) or the coalesced device layout (env_stride = 1,
 * comp_stride = n_envs).  ptr == NULL means "not wanted" for outputs.  The observation views of
 * mhppo_env_step need 0 <= comp_stride < 2^31 (MHPPO_EINVAL otherwise). */
typedef struct mhppo_view {
    float *ptr;
    int64_t env_stride;
    int64_t comp_stride;
} mhppo_view;

int mhppo_abi_version(void);
const char *mhppo_last_error(void);
int mhppo_device_count(void);

/* gym.make(id, **cfg) / Crosswalk_hybrid_multi_*.__init__  (SC:680-736, REG:3-41) */
int mhppo_env_create(const mhppo_env_cfg *cfg, void **handle);
int mhppo_env_destroy(void *handle);
int mhppo_env_get_dims(void *handle, mhppo_env_dims *dims);

/* Crosswalk_hybrid_multi_*.reset  (SC:884-946, CO:838-892, ST:847, NA:829, C4:850-911, C42:866-927).
 * mask_dev: uint8[n_envs] or NULL (= all). obs receives the first observation of reset envs only. */
int mhppo_env_reset(void *handle, const uint8_t *mask_dev, mhppo_view obs_dev, void *stream);

/* State injection of the reference (fixed scenarios: choix_test PY:629-633, reset_distrib NA:903-913), for the envs with
 * mask_dev[n] != 0 (NULL = all).  One parameter row per env, device memory:
 *   reset_pedestrian (SC:948-955; CO:895, ST:904, C4:913, C42:929): params [N][9] = speed_x, speed_y, pos_x, pos_y, dl, leave, CZ,
 *     exist, direction.  Like the reference it first rebuilds EVERY pedestrian as a placeholder (their constructor draws are
 *     consumed from the env's stream), then applies pedestrian.reset_ped (SC:106-137) to slot num_ped; `leave` / `CZ` are stored
 *     as truth values.  naif (NA:897-898, 100-135): params = speed_x, speed_y, pos_x, pos_y, dl, direction, cross, -, - ; no
 *     rebuild; `cross` becomes the env's crossing width.
 *   reset_cars (SC:957-958 -> car.reset_car SC:583-587): params [N][4] = speed_x, pos_x, light, line (0 <= line < nb_lines: the
 *     kernels evaluate the lane predicates once per lane of the crossing).
 * Call mhppo_env_observe afterwards for the observation (the reference calls get_state(), PY:172-173). */
int mhppo_env_reset_pedestrian(void *handle, int32_t num_ped, const float *params_dev, const uint8_t *mask_dev, void *stream);
int mhppo_env_reset_cars(void *handle, int32_t num_car, const float *params_dev, const uint8_t *mask_dev, void *stream);

/* Crosswalk_hybrid_multi_*.get_state (SC:960-969): observation of the current state without stepping (after an
 * import_state / state injection).  As in the reference, get_data folds the current gap into every pedestrian's
 * running-min `delta` (SC:457), i.e. the call is not idempotent on that one field. */
int mhppo_env_observe(void *handle, mhppo_view obs_dev, void *stream);

/* Crosswalk_hybrid_multi_*.step  (SC:789-878, CO:745-832, ST:754, NA:736, C4:783-844, C42:799-860).
 * actions [n_envs,n_action]; obs [n_envs,n_obs]; rewards, reward_light (= env.reward_light, SC:846)
 * [n_envs,n_lead]; done uint8[n_envs].  autoreset != 0: a done env is re-initialised inside the same
 * kernel (gym vector convention): obs = first observation of the new episode, term_obs (optional) =
 * terminal observation. */
int mhppo_env_step(void *handle, mhppo_view actions_dev, mhppo_view obs_dev, mhppo_view rewards_dev,
                   mhppo_view reward_light_dev, uint8_t *done_dev, int autoreset, mhppo_view term_obs_dev,
                   void *stream);

/* Same call for HOST buffers in the reference's own layout (row-major [n_envs, width], what
 * env.step(np.ndarray) takes and returns): copies actions host->device, steps, copies the results
 * back and synchronises `stream`.  The env range is processed in slices on streams owned by the handle so that the
 * device->host copy of one slice overlaps the host->device copy and the kernel of the next; the device staging
 * buffers live in the handle, the host buffers are the caller's (pin them for full PCIe rate). */
int mhppo_env_step_host(void *handle, const float *actions_host, float *obs_host, float *rewards_host,
                        float *reward_light_host, uint8_t *done_host, int autoreset, void *stream);
int mhppo_env_reset_host(void *handle, float *obs_host, void *stream);

/* get_state()/reset_pedestrian()/reset_cars() of the reference read/poke Python attributes
 * (SC:948-969).  The vectorised equivalent moves the whole state as the canonical dump
 *   car_f [N,C,7] Ac,Vc,Sc,light,possible_accident,error_scenario,Ts      car_i [N,C,2] line,exist
 *   ped_f [N,P,9] Vp_x,Vp_y,Sp_x,Sp_y,v0x,v0y,cross_stop,delta,worst_dl
 *   ped_i [N,P,9] t0/dt,waiting_time/dt,crossing_time/dt,time_stop,line_pos,direction,gender,age,flags
 *   env_f [N,1]   cross (fp64)                env_i [N,4] step_idx,ped_traffic,car_traffic,rng_ctr
 * flags bits: 0 exist,1 is_crossing,2 decision,3 at_crossing,4 ped_left,5 ped_in_cross,
 * 6 ped_not_waiting,7 accident,8 worst_scenario_accident,9 follow_rule,10 stop,11 need_to_stop,
 * 12 ratio_eps (set by reset_ped: x follows y with ratio v0x / (v0y + 1e-3), SC:111, instead of v0x / v0y, SC:73).
 * All pointers are device memory. */
int mhppo_env_export_state(void *handle, float *car_f_dev, int32_t *car_i_dev, float *ped_f_dev,
                           int32_t *ped_i_dev, double *env_f_dev, int64_t *env_i_dev, void *stream);
int mhppo_env_import_state(void *handle, const float *car_f_dev, const int32_t *car_i_dev,
                           const float *ped_f_dev, const int32_t *ped_i_dev, const double *env_f_dev,
                           const int64_t *env_i_dev, void *stream);

/* bytes of HBM the handle holds per env (state arena), for the roofline accounting */
int64_t mhppo_env_state_bytes_per_env(void *handle);
/* number of kernel launches this library has issued in this process (bench.py "gpu_launches") */
int64_t mhppo_launch_count(void);

/* =====================================================================================================
 * PPO side: Model_PPO / Env_rollout / Algo_PPO of Coop-MH-PPO-scalable.py (scalable env class).
 *
 * Networks (Model_PPO, PY:42-93): in -> 32 -> 64 -> 32 -> out, ReLU; heads: 0 linear (critics),
 * 1 3*tanh(z)-1 (cross / wait actors, PY:1044-1045), 2 softmax over two logits (choice actor).
 * Flat parameter layout shared by every kernel (host packs/unpacks state_dicts):
 *   W1t[KP][32] b1[32] W2t[32][64] b2[64] W3t[64][32] b3[32] W4t[32][4] b4[4]   (weights transposed, zero padded;
 *   KP = mhppo_net_padded_in(n_in)).
 *
 * Buffers (device, fp32, env index innermost): obs [n_obs][N] (the env's component-major observation);
 * action_d int8 [C*P][N]; light [C][N]; sample s = (t*C + car)*N + env, S = T*C*N: obs_c [13][S], act, logp,
 * rew, rl, rtg, V [S]; choice sample m = car*N + env, M = C*N: obs_d [D][M], act_d, logp_d, rew_d [M];
 * route int8 [C][N] (0 cross, 1 wait, -1 car absent; PY:489-502).  C = 2*nb_lines, D = 2+6*(C-1)+10. */
typedef struct mhppo_rollout_cfg {
    int32_t nb_ped, nb_lines;
    int32_t T;            /* steps per episode (80) */
    int32_t legacy_nb_car; /* 0: Coop-MH-PPO-scalable.py layout (C = 2*nb_lines slots of 7 floats, env row of 4).  > 0: the older
                            * notebooks' layout on the coop / naif / stop env classes (Coop-MH-PPO.ipynb, MH-PPO.ipynb): C = this many
                            * cars of 6 floats, env row of 3, D = 2+5*(C-1)+10, every pedestrian slot visited, closest pedestrian
                            * seeded with slot 0 and not filtered on `exist` */
    int64_t n_envs;
    uint64_t seed;        /* same Philox key as the env */
    int64_t env_id0;
} mhppo_rollout_cfg;

/* Gaussian head of the cross / wait actors and its exploration noise, for the calling thread's later rollout / update calls:
 * mu = tanh(z) * std + mean (Model_PPO head type 1, PY:88-90; the driver derives mean = (car_b[1,0] + car_b[0,0]) / 2 and
 * std = (car_b[1,0] - car_b[0,0]) / 2, PY:1044-1045), a ~ N(mu, variance) (Algo_PPO.value_std, PY:726-729), acc_hi = car_b[1,0]
 * is the start value of the min over pedestrians (PY:436).  Defaults: -1, 3, 0.5, 2. */
int mhppo_set_gaussian_head(float mean, float std, float variance, float acc_hi);
int mhppo_net_padded_in(int32_t n_in);
int mhppo_net_param_count(int32_t n_in);

/* episode-start discrete decision of Env_rollout.iterations_rand (PY:400-428): obs_car_ped_d (PY:574-611) ->
 * choice net -> Categorical sample per (car, ped); closest_ped_d (PY:614-627) picks the car's light */
int mhppo_choice_act(const mhppo_rollout_cfg *cfg, const float *obs_dev, const float *net_choice_dev, uint32_t iteration,
                     int8_t *action_d_dev, float *light_dev, float *obs_d_dev, float *act_d_dev, float *logp_d_dev,
                     void *stream);
/* per-step continuous action (PY:434-453): obs_car_ped (PY:541-572) -> cross|wait net per pair -> min over existing
 * pedestrians -> N(mean, 0.5) sample + log-prob; writes the env's action buffer [2C][N] and the rollout buffers */
int mhppo_policy_act(const mhppo_rollout_cfg *cfg, const float *obs_dev, const float *net_cross_dev, const float *net_wait_dev,
                     const int8_t *action_d_dev, const float *light_dev, int32_t t, uint32_t iteration, float *actions_dev,
                     float *obs_c_dev, float *act_dev, float *logp_dev, void *stream);
/* The whole T-step inner loop of Env_rollout.iterations_rand (PY:386-476) in one call: per step policy_act, then env.step
 * with rewards / reward_light written into rew_dev / rl_dev [T][C][N] (no auto-reset: the episode ends at step T-1).
 * use_graph != 0: from the second call with the same arguments on, the 2T launches are replayed as one CUDA graph on a
 * stream owned by the library (launch-bound at small env counts); the caller's stream is ordered around it by events. */
int mhppo_rollout_steps(void *env_handle, const mhppo_rollout_cfg *cfg, float *obs_dev, const float *net_cross_dev,
                        const float *net_wait_dev, const int8_t *action_d_dev, const float *light_dev, uint32_t iteration,
                        float *actions_dev, float *obs_c_dev, float *act_dev, float *logp_dev, float *rew_dev, float *rl_dev,
                        uint8_t *done_dev, int32_t use_graph, void *stream);
/* deterministic evaluation rollout, Env_rollout.iterations (PY:152-252).
 * choice_eval: argmax of the choice net per (car, ped) (PY:177-192) for the envs that re-decide: all when force != 0 (first
 *   step of an episode), else those whose state has ped_traffic != nb_ped (PY:222-224); other envs keep action_d.
 * policy_eval: per car, min over ALL pedestrian slots of {net output | speed-recovery clip((speed_limit - v)/dt, acc_lo,
 *   acc_hi) when the pedestrian has left the lane}, capped by (10 - v)/dt (PY:195-214); writes the env's action buffer
 *   [2C][N] (lights = the first C entries of action_d, PY:192/216) and optionally the accelerations act_dev [C][N]. */
int mhppo_choice_eval(const mhppo_rollout_cfg *cfg, const float *obs_dev, const float *net_choice_dev, int32_t force,
                      int8_t *action_d_dev, void *stream);
int mhppo_policy_eval(const mhppo_rollout_cfg *cfg, const float *obs_dev, const float *net_cross_dev, const float *net_wait_dev,
                      const int8_t *action_d_dev, float dt, float speed_limit, float acc_lo, float acc_hi, float *actions_dev,
                      float *act_dev, void *stream);
/* Episode statistics of the evaluation records (the analysis cell get_average, PY:1550-1675, and the waiting-time histogram input
 * PY:1390-1403) on the device.  obs_rec: the dense record of Env_rollout.iterations, [E][T][n_obs][N] (state before each step).
 * ped_out [2][E][P][N]: waiting time on the kerb, crossing time, seconds (PY:1591-1601).  car_out [8][E][C][N]: time to 25 m past the
 * pedestrians (PY:1622-1631), free-flow time (25 - x0)/v0 (PY:1612), light at the last row (PY:1620), light after the first step
 * (PY:1613), could-stop flag (-1 if light1 >= 0, else 0 / 1, PY:1613-1617), number / sum / sum of squares of the 25 m passage times
 * (PY:1624-1625).  sums [E][N][8] = rows, sum v0, sum v0^2, sum |a0|, sum a0, sum a0^2, sum |vp0|, sum vp0^2 (PY:1552-1562);
 * yield_speed [E][N][3] = n, sum v, sum v^2 over the rows of cars with light1 == 1 (PY:1621).  The printed means / standard
 * deviations / scenario frequencies are reductions of these arrays (host mirror: Env_rollout.get_average). */
int mhppo_episode_stats(const float *obs_rec_dev, int32_t E, int32_t T, int32_t n_cars, int32_t n_ped, int32_t car_w, int32_t env_w,
                        int64_t N, float dt, float *ped_out_dev, float *car_out_dev, double *sums_dev, double *yield_speed_dev,
                        void *stream);
/* Env_rollout.futur_rewards (PY:658-684): reverse scan rtg_t = r_t + gamma*rtg_{t+1}, zero bootstrap; and the
 * episodic choice reward min(0, min_t reward_light) (PY:461).  CN = C*N */
int mhppo_returns(const float *rew_dev, const float *rl_dev, int32_t T, int64_t CN, double gamma, float *rtg_dev,
                  float *rew_d_dev, void *stream);

/* Algo_PPO.train_model_c / train_model_d (PY:778-851), one epoch = value_stats -> [all-reduce] -> ppo_grad (actor),
 * ppo_grad (critic) -> [all-reduce] -> adam.  x: features [D][S], sample slot s = t*CN + m.  idx: int32[K], the
 * compacted list of the (car, env) columns m routed to the net being trained (PY:489-502), NULL = all CN columns. */
int64_t mhppo_update_workspace_bytes(int32_t n_in);
int mhppo_value_stats(int32_t n_in, const float *x_dev, int32_t D, int64_t S, const int32_t *idx_dev, int64_t K, int64_t CN,
                      const float *critic_dev, const float *rtg_dev, float *V_dev, double *stats3_dev, void *workspace_dev,
                      void *stream);
/* head: 0 critic MSE, 1 Gaussian actor clipped surrogate, 2 categorical actor (with the (M,M) broadcast of PY:834-842,
 * f0/f1 = fraction of selected samples whose action is 0/1).  grad_dev: flat gradient, loss_dev: fp64 scalar. */
int mhppo_ppo_grad(int32_t n_in, int32_t head, const float *x_dev, int32_t D, int64_t S, const int32_t *idx_dev, int64_t K,
                   int64_t CN, const float *net_dev, const float *act_dev, const float *logp_old_dev, const float *rtg_dev,
                   const float *V_dev, float adv_mean, float adv_inv_std, float inv_n, float f0, float f1, float *grad_dev,
                   double *loss_dev, void *workspace_dev, void *stream);
/* mhppo_ppo_grad with the advantage statistics taken from DEVICE memory: adv_stats_dev = (sum A, sum A^2, n) over the whole batch
 * (the output of mhppo_critic_grad_stats, all-reduced over the ranks); mean and 1 / (unbiased std + 1e-10) (PY:787) are derived
 * inside the kernel, so an epoch needs no host round trip. */
int mhppo_ppo_grad_dev(int32_t n_in, int32_t head, const float *x_dev, int32_t D, int64_t S, const int32_t *idx_dev, int64_t K,
                       int64_t CN, const float *net_dev, const float *act_dev, const float *logp_old_dev, const float *rtg_dev,
                       const float *V_dev, const double *adv_stats_dev, float inv_n, float f0, float f1, float *grad_dev,
                       double *loss_dev, void *workspace_dev, void *stream);
/* critic pass of one epoch in a single kernel: ppo_grad with head 0 that also returns V = critic(s) (PY:785) and the
 * advantage statistics (sum A, sum A^2, n) of A = rtg - V, so an epoch is critic_grad_stats -> [all-reduce stats] ->
 * ppo_grad (actor) -> [all-reduce grads] -> adam, adam; V and the statistics are those of the critic before its step,
 * as in the reference (the actor and critic steps of PY:810-815 act on different networks).  inv_n = 1 / global count */
int mhppo_critic_grad_stats(int32_t n_in, const float *x_dev, int32_t D, int64_t S, const int32_t *idx_dev, int64_t K, int64_t CN,
                            const float *critic_dev, const float *rtg_dev, float inv_n, float *grad_dev, double *loss_dev,
                            float *V_dev, double *stats3_dev, void *workspace_dev, void *stream);
/* torch.optim.Adam defaults (PY:719-724); step counts from 1; grad_scale multiplies the gradient first */
int mhppo_adam(float *param_dev, const float *grad_dev, float *m_dev, float *v_dev, int32_t n, float lr, float beta1,
               float beta2, float eps, int32_t step, float grad_scale, void *stream);

/* the actor's and the critic's Adam step of one epoch in one launch (they act on different nets, PY:810-815) */
int mhppo_adam2(float *p0_dev, const float *g0_dev, float *m0_dev, float *v0_dev, int32_t n0, float lr0, int32_t step0, float *p1_dev,
                const float *g1_dev, float *m1_dev, float *v1_dev, int32_t n1, float lr1, int32_t step1, float beta1, float beta2,
                float eps, void *stream);

/* Self-test of the tcgen05 / TMEM building blocks behind the policy GEMMs: D[128,N] = A[128,K] * B[N,K]^T on one CTA,
 * mode 0 = one tf32 pass, mode 1 = 3xTF32 split accumulation.  (N,K) in {(64,32),(32,64),(64,16),(16,32)}. */
/* implementation of the 13-input nets' kernels: 0 auto (tcgen05 for the update: forward + backward-data of ppo_grad /
 * critic_grad_stats and the critic forward of value_stats; exact-fp32 FFMA for the rollout inference), 1 FFMA everywhere
 * (exact fp32 cross-check), 2 tcgen05 wherever a tensor-core kernel exists (also the rollout inference) */
int mhppo_set_mlp_mode(int32_t mode);
/* 1 if a tensor-core kernel ever timed out waiting for its MMAs (diagnostic; synchronises the device) */
int mhppo_tc_failures(void);
/* Self-test of the step kernel's gap-acceptance predicate (`car_time + light < CG`, SC:167-170, 419-428): for n samples
 * (device arrays) returns the filtered decision the kernel takes (fast), the level that took it (0 = no draw needed,
 * 1 = fp32 draw, 2 = exact fallback), the exact fp64 decision and the exact critical gap (cg_dev may be NULL).  The normal
 * draw of sample i is Philox block ctr[i] of stream (env_id, key0, key1).  fast must equal exact for every input. */
int mhppo_gap_selftest(int64_t n, const double *dx_dev, const double *vden_dev, const double *light_dev, const double *size_dev,
                       const double *v0y_dev, const int32_t *gender_dev, const int32_t *age_dev, const uint32_t *ctr_dev,
                       uint64_t env_id, uint32_t key0, uint32_t key1, uint8_t *fast_dev, uint8_t *level_dev, uint8_t *exact_dev,
                       double *cg_dev, void *stream);
int mhppo_tc_selftest(const float *A_dev, const float *B_dev, float *D_dev, int32_t N, int32_t K, int32_t mode, void *stream);

#ifdef __cplusplus
}
#endif
#endif
