"""CPU restatement of the PPO side of the MH-PPO hot path.  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import this module; the
product (mh-ppo_b200/) never does.  Citations: PY = /root/reference/Coop-MH-PPO-scalable.py.

Parity status: PINNED.  tests/test_ppo_oracle_vs_reference.py checks every function here against the
unmodified reference classes (Model_PPO, Env_rollout, Algo_PPO, loaded by oracle/refshim/refppo.py)
in the build container; tests/golden/ppo_*.npz are fixtures generated from the reference by
tools/gen_golden_ppo.py for machines without /root/reference.

What is restated
  feature builders   Env_rollout.obs_car_ped / obs_car_ped_d / closest_ped_d / is_in_cross / leave_cross
                     (PY:526-627), vectorised over envs, float32 like the reference under numpy 2
  policy networks    Model_PPO.forward (PY:70-93) on a state_dict in the reference's key layout
  rollout step       the per-step action selection of Env_rollout.iterations_rand (PY:434-453) and the
                     episode-start discrete decision (PY:400-428), with the sampling noise passed in
  returns            Env_rollout.futur_rewards (PY:658-684)
  update             Algo_PPO.train_model_c / train_model_d (PY:778-851) incl. the (M,M) categorical
                     broadcast of train_model_d, with torch autograd and torch.optim.Adam
"""
import math

import numpy as np
import torch

F32 = np.float32


def obs_layout(nb_ped, nb_lines, legacy_nb_car=0):
    """Slices of the flat observation (gym Dict key order car, env, ped; PY:543-554).  legacy_nb_car = 0: the scalable
    script (2*nb_lines car slots of 7 floats, env row of 4).  legacy_nb_car > 0: the two older notebooks on the coop / naif
    env classes (Coop-MH-PPO.ipynb / MH-PPO.ipynb, identical PPO cells): nb_car cars of 6 floats, env row of 3 (lines last)."""
    if legacy_nb_car:
        size_car = 6 * legacy_nb_car
        return dict(size_car=size_car, cw=6, size_env=3, lines_k=2, ped0=size_car + 3, C=legacy_nb_car, P=nb_ped, legacy=True)
    size_car = 7 * 2 * nb_lines
    return dict(size_car=size_car, cw=7, size_env=4, lines_k=3, ped0=size_car + 4, C=2 * nb_lines, P=nb_ped, legacy=False)


def _lane(car_line, ped_pos, ped_dir, cross_lines, lines):
    """is_in_cross + leave_cross, PY:526-539 (float32 arithmetic, reference operation order)."""
    cross = (cross_lines * F32(2)) / lines
    line_start = (-cross_lines) + cross * car_line
    line_end = (-cross_lines) + cross * (car_line + F32(1))
    dist_start = (ped_pos - line_start) * (ped_dir > 0) + (line_end - ped_pos) * (ped_dir < 0)
    crossing = (ped_pos > line_start) & (ped_pos < line_end)
    neg = ped_dir == -1
    end_cross = np.where(neg, ped_pos < line_start, ped_pos > line_end)
    dist_end = np.where(neg, line_start - ped_pos, ped_pos - line_end)
    return crossing, dist_start.astype(F32), end_cross, dist_end.astype(F32)


def obs_car_ped(obs, i, p, nb_ped, nb_lines, legacy_nb_car=0):
    """Env_rollout.obs_car_ped, PY:541-572 (same arithmetic in the older notebooks, on their 6 / 3-float rows).
    obs [N, n_obs] float32 -> (features [N,13] float32, exist [N])."""
    L = obs_layout(nb_ped, nb_lines, legacy_nb_car)
    obs = np.asarray(obs, F32)
    cw, lk = L["cw"], L["lines_k"]
    car = obs[:, i * cw:(i + 1) * cw]
    ped = obs[:, L["ped0"] + p * 9: L["ped0"] + (p + 1) * 9]
    env = obs[:, L["size_car"]: L["size_car"] + L["size_env"]]
    crossing, dist_start, end_cross, dist_end = _lane(car[:, 5], ped[:, 3], ped[:, 8], env[:, 0], env[:, lk])
    front = ped[:, 2] > car[:, 3]
    dx = ped[:, 2] - car[:, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        ttc = np.minimum(F32(10.0), dx / np.maximum(car[:, 1] - ped[:, 0], F32(0.01))) * front + F32(10.0) * (~front)
    f = np.stack([car[:, 1], car[:, 2], ped[:, 1], front.astype(F32), dx, ped[:, 4], crossing.astype(F32), end_cross.astype(F32),
                  dist_start, dist_end, ttc.astype(F32), env[:, 0], env[:, lk]], axis=1).astype(F32)
    return f, ped[:, 7]


def obs_car_ped_d(obs, i, p, nb_ped, nb_lines, legacy_nb_car=0):
    """Env_rollout.obs_car_ped_d, PY:574-611 -> [N, 2+6(2L-1)+8+2] (note car_data[6], the OWN exist flag, PY:593);
    the older notebooks have 5 columns per other car (no exist flag) -> [N, 2+5(nb_car-1)+8+2]."""
    L = obs_layout(nb_ped, nb_lines, legacy_nb_car)
    obs = np.asarray(obs, F32)
    cw, lk = L["cw"], L["lines_k"]
    car = obs[:, i * cw:(i + 1) * cw]
    ped = obs[:, L["ped0"] + p * 9: L["ped0"] + (p + 1) * 9]
    env = obs[:, L["size_car"]: L["size_car"] + L["size_env"]]
    cols = [car[:, 1], car[:, 2]]
    for k in range(L["C"]):
        if k == i:
            continue
        c2 = obs[:, k * cw:(k + 1) * cw]
        cols += [c2[:, 1], (ped[:, 2] > c2[:, 3]).astype(F32), ped[:, 2] - c2[:, 3], c2[:, 4], c2[:, 5] - car[:, 5]]
        if not L["legacy"]:
            cols.append(car[:, 6])
    crossing, dist_start, end_cross, dist_end = _lane(car[:, 5], ped[:, 3], ped[:, 8], env[:, 0], env[:, lk])
    cols += [ped[:, 1], (ped[:, 2] > car[:, 3]).astype(F32), ped[:, 2] - car[:, 3], ped[:, 4], crossing.astype(F32),
             end_cross.astype(F32), dist_start, dist_end, env[:, 0], env[:, lk]]
    return np.stack(cols, axis=1).astype(F32), ped[:, 7]


def closest_ped_d(obs, i, nb_ped, nb_lines, legacy_nb_car=0):
    """Env_rollout.closest_ped_d, PY:614-627: argmin of the SIGNED ped_x - car_x over existing peds.  The older notebooks
    seed the minimum with pedestrian 0 and do not look at `exist`."""
    L = obs_layout(nb_ped, nb_lines, legacy_nb_car)
    obs = np.asarray(obs, F32)
    best = np.zeros(obs.shape[0], np.int64)
    dmin = np.full(obs.shape[0], 1000000.0, np.float64)
    if L["legacy"]:
        dmin = (obs[:, L["ped0"] + 2] - obs[:, i * L["cw"] + 3]).astype(np.float64)
    for p in range(nb_ped):
        ped = obs[:, L["ped0"] + p * 9: L["ped0"] + (p + 1) * 9]
        d = (ped[:, 2] - obs[:, i * L["cw"] + 3]).astype(np.float64)
        take = (d < dmin) if L["legacy"] else ((d < dmin) & (ped[:, 7] != 0))
        dmin = np.where(take, d, dmin)
        best = np.where(take, p, best)
    return best


# --------------------------------------------------------------------------------------------------
def mlp_forward(sd, x, model_type, mean=-1.0, std=3.0):
    """Model_PPO.forward, PY:70-93.  sd: state_dict with keys layer{1..4}.{weight,bias}; x [B, in] float32."""
    x = torch.as_tensor(x, dtype=torch.float32)
    h = torch.relu(torch.nn.functional.linear(x, sd["layer1.weight"], sd["layer1.bias"]))
    h = torch.relu(torch.nn.functional.linear(h, sd["layer2.weight"], sd["layer2.bias"]))
    h = torch.relu(torch.nn.functional.linear(h, sd["layer3.weight"], sd["layer3.bias"]))
    o = torch.nn.functional.linear(h, sd["layer4.weight"], sd["layer4.bias"])
    if model_type == 2:
        return torch.softmax(o.reshape(-1, 2), dim=-1)
    if model_type == 1:
        return torch.tanh(o) * std + mean
    return o


LOG_PI_HALF = 0.5 * math.log(math.pi)


def continuous_step(obs, action_d, sd_cross, sd_wait, z, nb_ped, nb_lines, acc_hi=2.0, legacy_nb_car=0):
    """Per-step action selection of iterations_rand, PY:434-453, for every env and car slot.

    obs [N,n_obs] f32; action_d [N, C*P] in {-1,+1}; z [N,C] standard normal noise.
    Returns mean [N,C], action [N,C], logp [N,C], state_c [N,C,13] (features of the arg-min pedestrian).
    MultivariateNormal(mean, 0.5 I): a = mean + sqrt(.5) z, logp = -(a-mean)^2 - ln(pi)/2.
    """
    obs = np.asarray(obs, F32)
    N = obs.shape[0]
    C, P = (legacy_nb_car or 2 * nb_lines), nb_ped
    mean = np.full((N, C), acc_hi, F32)
    state = np.zeros((N, C, 13), F32)
    for i in range(C):
        state[:, i], _ = obs_car_ped(obs, i, 0, nb_ped, nb_lines, legacy_nb_car)
        for p in range(P):
            f, exist = obs_car_ped(obs, i, p, nb_ped, nb_lines, legacy_nb_car)
            if legacy_nb_car:
                exist = np.ones_like(exist)                             # the older notebooks visit every pedestrian slot
            with torch.no_grad():
                m_cross = mlp_forward(sd_cross, f, 1).numpy()[:, 0]
                m_wait = mlp_forward(sd_wait, f, 1).numpy()[:, 0]
            m = np.where(action_d[:, i * P + p] <= 0, m_cross, m_wait).astype(F32)
            ex = exist != 0
            newmin = np.where(ex, np.minimum(mean[:, i], m), mean[:, i])
            take = ex & (m == newmin)                                   # PY:448-450
            mean[:, i] = newmin
            state[:, i] = np.where(take[:, None], f, state[:, i])
    a = (mean + F32(math.sqrt(0.5)) * np.asarray(z, F32)).astype(F32)
    logp = (-(a - mean) ** 2 - F32(LOG_PI_HALF)).astype(F32)
    return mean, a, logp, state


def discrete_step(obs, sd_choice, u, nb_ped, nb_lines, legacy_nb_car=0):
    """Episode-start decision of iterations_rand, PY:400-428.  u [N, C*P] uniforms in [0,1).

    Categorical(probs).sample() is restated as `action = 1 if u >= p0 else 0`.
    Returns action_all_d [N,C*P] in {0,1}, logp_all [N,C*P], feat_d [N,C*P,D], and per car the
    light's (action, logp, feature, closest ped).
    """
    obs = np.asarray(obs, F32)
    N = obs.shape[0]
    C, P = (legacy_nb_car or 2 * nb_lines), nb_ped
    acts = np.zeros((N, C * P), np.int64)
    logps = np.zeros((N, C * P), F32)
    feats = []
    eps = float(np.finfo(np.float32).eps)
    for i in range(C):
        for p in range(P):
            f, _ = obs_car_ped_d(obs, i, p, nb_ped, nb_lines, legacy_nb_car)
            feats.append(f)
            with torch.no_grad():
                pr = mlp_forward(sd_choice, f, 2).numpy()
            a = (np.asarray(u)[:, i * P + p] >= pr[:, 0]).astype(np.int64)
            pa = np.clip(np.where(a == 0, pr[:, 0], pr[:, 1]) / (pr[:, 0] + pr[:, 1]), eps, 1 - eps)
            acts[:, i * P + p] = a
            logps[:, i * P + p] = np.log(pa).astype(F32)
    feats = np.stack(feats, axis=1)
    closest = np.stack([closest_ped_d(obs, i, nb_ped, nb_lines, legacy_nb_car) for i in range(C)], axis=1)
    idx = np.arange(C)[None, :] * P + closest
    rows = np.arange(N)[:, None]
    return acts, logps, feats, acts[rows, idx], logps[rows, idx], feats[rows, idx], closest


def eval_discrete(obs, sd_choice, nb_ped, nb_lines):
    """Deterministic decisions of Env_rollout.iterations, PY:177-190: argmax of the choice net per (car, pedestrian).
    Returns action_all_d [N, C*P] in {0, 1} (torch.argmax: the first index on a tie)."""
    obs = np.asarray(obs, F32)
    C, P = 2 * nb_lines, nb_ped
    acts = np.zeros((obs.shape[0], C * P), np.int64)
    for i in range(C):
        for p in range(P):
            f, _ = obs_car_ped_d(obs, i, p, nb_ped, nb_lines)
            with torch.no_grad():
                pr = mlp_forward(sd_choice, f, 2)
            acts[:, i * P + p] = torch.argmax(pr, dim=1).numpy()
    return acts


def eval_continuous(obs, action_d, sd_cross, sd_wait, nb_ped, nb_lines, dt=0.3, speed_limit=10.0, acc_lo=-4.0, acc_hi=2.0):
    """Deterministic accelerations of Env_rollout.iterations, PY:195-214, fp32 like the reference's numpy scalars
    (NumPy >= 2 promotion: float32 op python-float stays float32).  Every pedestrian slot is visited, existing or not;
    a pedestrian that has left the car's lane (feature 7, end_cross) asks for the speed-recovery acceleration
    clip((speed_limit - v) / dt, acc_lo, acc_hi) instead of a net output; the running min is also capped by (10 - v) / dt."""
    obs = np.asarray(obs, F32)
    N = obs.shape[0]
    C, P = 2 * nb_lines, nb_ped
    out = np.zeros((N, C), F32)
    for i in range(C):
        a = np.full(N, acc_hi, F32)
        for p in range(P):
            f, _ = obs_car_ped(obs, i, p, nb_ped, nb_lines)
            with torch.no_grad():
                m_cross = mlp_forward(sd_cross, f, 1).numpy()[:, 0]
                m_wait = mlp_forward(sd_wait, f, 1).numpy()[:, 0]
            net = np.where(action_d[:, i * P + p] <= 0, m_cross, m_wait).astype(F32)
            rec = np.maximum(np.minimum((F32(speed_limit) - f[:, 0]) / F32(dt), F32(acc_hi)), F32(acc_lo)).astype(F32)
            new = np.where(f[:, 7] != 0, rec, net)
            a = np.where(new < a, new, a)                               # python min(a, new)
            cap = ((F32(10.0) - f[:, 0]) / F32(dt)).astype(F32)
            a = np.where(cap < a, cap, a)
        out[:, i] = a
    return out


def eval_episode(env, sd_cross, sd_wait, sd_choice, nb_ped, nb_lines, T=80, dt=0.3):
    """One deterministic evaluation episode per env of Env_rollout.iterations (PY:152-252) on a vectorised env oracle.
    The decisions are taken at step 0 and again before every step whose state has ped_traffic != nb_ped (PY:222-224);
    the lights handed to env.step are the FIRST C entries of the per-(car, pedestrian) decision vector (PY:192, 216:
    `action_d_light = 2*action_all_d - 1` is appended whole and the env reads actions[C + i]).
    Returns obs [T,N,n_obs] (state before each step), acts [T,N,C], rew [T,N,C], rl [T,N,C], action_d [T,N,C*P],
    waiting [T,N,P] (pedestrian.waiting_time after each step), done [N]."""
    C, P = 2 * nb_lines, nb_ped
    obs = np.asarray(env.reset(), F32)
    N = obs.shape[0]
    need = np.ones(N, bool)
    acts_d = np.zeros((N, C * P), np.int64)
    out = dict(obs=[], acts=[], rew=[], rl=[], action_d=[], waiting=[])
    done = None
    for t in range(T):
        new_d = eval_discrete(obs, sd_choice, nb_ped, nb_lines)
        acts_d = np.where(need[:, None], new_d, acts_d)
        action_d = (2 * acts_d - 1).astype(np.float64)
        a = eval_continuous(obs, action_d, sd_cross, sd_wait, nb_ped, nb_lines, dt=dt)
        out["obs"].append(obs.copy()); out["acts"].append(a.copy()); out["action_d"].append(action_d.copy())
        obs, rew, rl, done = env.step(np.concatenate([a.astype(np.float64), action_d[:, :C]], axis=1), autoreset=False)
        obs = np.asarray(obs, F32)
        out["rew"].append(np.asarray(rew).copy()); out["rl"].append(np.asarray(rl).copy())
        out["waiting"].append(env.get_state()["ped_i"][:, :, 1].astype(np.float64) * dt)   # waiting_time is a multiple of dt
        need = obs[:, 7 * C + 1] != F32(P)
    res = {k: np.stack(v) for k, v in out.items()}
    res["done"] = np.asarray(done)
    return res


def reward_to_go(rews, gamma=0.99):
    """futur_rewards, PY:658-684: rews [T, ...] -> discounted reward-to-go with zero bootstrap (fp64 like the reference)."""
    rews = np.asarray(rews, np.float64)
    out = np.zeros_like(rews)
    acc = np.zeros(rews.shape[1:], np.float64)
    for t in range(rews.shape[0] - 1, -1, -1):
        acc = rews[t] + gamma * acc
        out[t] = acc
    return out


# --------------------------------------------------------------------------------------------------
class Net(torch.nn.Module):
    """Same parameters / key layout as Model_PPO (PY:42-68)."""

    def __init__(self, n_in, n_out, model_type, mean=-1.0, std=3.0):
        super().__init__()
        self.model_type, self.mean, self.std = model_type, mean, std
        self.layer1 = torch.nn.Linear(n_in, 32)
        self.layer2 = torch.nn.Linear(32, 64)
        self.layer3 = torch.nn.Linear(64, 32)
        self.layer4 = torch.nn.Linear(32, n_out)
        torch.nn.init.orthogonal_(self.layer4.weight)

    def forward(self, x):
        return mlp_forward(dict(self.named_parameters()), x, self.model_type, self.mean, self.std)


def train_step_c(actor, critic, opt_a, opt_c, states, actions, logp_old, rtgs):
    """Algo_PPO.train_model_c, PY:778-815 (one epoch).  Returns (actor_loss, critic_loss)."""
    s = torch.as_tensor(np.asarray(states), dtype=torch.float32)
    rtg = torch.as_tensor(np.asarray(rtgs), dtype=torch.float32).flatten()
    V = torch.squeeze(critic(s))
    adv = rtg - V
    adv = (adv - adv.mean()) / (adv.std() + 1e-10)
    mu = actor(s).reshape(-1)
    a = torch.as_tensor(np.asarray(actions)).reshape(-1)
    logp = -((a - mu) ** 2) - LOG_PI_HALF           # MultivariateNormal(mu, 0.5 I_1).log_prob
    ratio = torch.exp(logp - torch.as_tensor(np.asarray(logp_old)))
    actor_loss = (-torch.min(ratio * adv, torch.clamp(ratio, 0.8, 1.2) * adv)).mean()
    critic_loss = torch.nn.functional.mse_loss(V.float(), rtg.float())
    opt_a.zero_grad()
    actor_loss.backward(retain_graph=True)
    opt_a.step()
    opt_c.zero_grad()
    critic_loss.backward()
    opt_c.step()
    return float(actor_loss), float(critic_loss)


def train_step_d(actor, critic, opt_a, opt_c, states, actions, logp_old, rtgs):
    """Algo_PPO.train_model_d, PY:818-851 (one epoch), including the (M,1) action broadcast that makes the
    log-prob / ratio / loss an (M,M) matrix."""
    s = torch.as_tensor(np.asarray(states), dtype=torch.float32)
    rtg = torch.as_tensor(np.asarray(rtgs), dtype=torch.float32).flatten()
    V = torch.squeeze(critic(s))
    adv = rtg - V
    adv = (adv - adv.mean()) / (adv.std() + 1e-10)
    probs = actor(s).reshape(-1, 2)
    a = torch.as_tensor(np.asarray(actions)).reshape(-1, 1)
    logp = torch.distributions.Categorical(probs).log_prob(a)           # (M, M)
    ratio = torch.exp(logp - torch.as_tensor(np.asarray(logp_old)))
    actor_loss = (-torch.min(ratio * adv, torch.clamp(ratio, 0.8, 1.2) * adv)).mean()
    critic_loss = torch.nn.functional.mse_loss(V.float(), rtg.float())
    opt_a.zero_grad()
    actor_loss.backward(retain_graph=True)
    opt_a.step()
    opt_c.zero_grad()
    critic_loss.backward()
    opt_c.step()
    return float(actor_loss), float(critic_loss)


# --------------------------------------------------------------------------------------------------
# policy-noise contract (DESIGN.md "RNG"): Philox4x32-10 blocks keyed like the env stream,
# counter = (index, stream | iteration << 8, env_lo, env_hi), stream 1 = continuous head (index =
# t*C + car, standard normal by Box-Muller on the block), stream 2 = discrete head (index = car*P + ped,
# uniform from words 0,1).  Vectorised numpy implementation.
def philox_blocks(index, stream_word, env_id, seed):
    c0 = np.asarray(index, np.uint64) & np.uint64(0xffffffff)
    env_id = np.asarray(env_id, np.uint64)
    c0, env_id = np.broadcast_arrays(c0, env_id)
    c0 = c0.copy()
    c1 = np.full(c0.shape, stream_word, np.uint64)
    c2 = env_id & np.uint64(0xffffffff)
    c3 = env_id >> np.uint64(32)
    k0, k1 = np.uint64(seed & 0xffffffff), np.uint64((seed >> 32) & 0xffffffff)
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xffffffff)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ k0
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ k1
        c1, c3, c0, c2 = p1 & MASK, p0 & MASK, n0 & MASK, n2 & MASK
        k0 = (k0 + np.uint64(0x9E3779B9)) & MASK
        k1 = (k1 + np.uint64(0xBB67AE85)) & MASK
    return c0, c1, c2, c3


def _u53(a, b):
    return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0


def policy_normal(index, env_id, seed, iteration=0):
    w = philox_blocks(index, 1 | (iteration << 8), env_id, seed)
    return np.sqrt(-2.0 * np.log(1.0 - _u53(w[0], w[1]))) * np.cos(2.0 * np.pi * _u53(w[2], w[3]))


def policy_uniform(index, env_id, seed, iteration=0):
    w = philox_blocks(index, 2 | (iteration << 8), env_id, seed)
    return _u53(w[0], w[1])


def rollout_episode(env, sd_cross, sd_wait, sd_choice, seed, env_ids, nb_ped, nb_lines, T=80, iteration=0, legacy_nb_car=0):
    """One episode per env of Env_rollout.iterations_rand (PY:357-516) on a vectorised env oracle
    (oracle.OracleVecEnv, scalable class).  Returns the rollout buffers in the vectorised layout:
      obs_c [T,C,N,13], act [T,C,N], logp [T,C,N], rew [T,C,N], route [C,N] (0 cross / 1 wait / -1 car absent),
      obs_d [C,N,D], act_d [C,N], logp_d [C,N], rew_d [C,N] (min over the episode of reward_light, PY:461)."""
    C, P = (legacy_nb_car or 2 * nb_lines), nb_ped
    env_ids = np.asarray(env_ids, np.int64)
    N = env_ids.shape[0]
    obs = env.reset()
    exist = env.get_state()["car_i"][:, :C, 1].astype(bool)             # cars_exist, PY:380 (all cars in the older notebooks)
    if legacy_nb_car:
        exist = np.ones_like(exist)
    u = np.stack([policy_uniform(k, env_ids, seed, iteration) for k in range(C * P)], axis=1)
    acts_d, logps_d, feats_d, light_a, light_lp, light_f, _ = discrete_step(obs, sd_choice, u.astype(np.float32), nb_ped, nb_lines,
                                                                            legacy_nb_car)
    action_d = (2 * acts_d - 1).astype(np.float64)                      # PY:423
    light = (2 * light_a - 1).astype(np.float64)                        # PY:424
    buf = dict(obs_c=np.zeros((T, C, N, 13), F32), act=np.zeros((T, C, N), F32), logp=np.zeros((T, C, N), F32),
               rew=np.zeros((T, C, N), np.float64))
    ep_rew = np.zeros((N, C), np.float64)
    for t in range(T):
        z = np.stack([policy_normal(t * C + i, env_ids, seed, iteration) for i in range(C)], axis=1)
        mean, a, logp, state = continuous_step(obs, action_d, sd_cross, sd_wait, z, nb_ped, nb_lines, legacy_nb_car=legacy_nb_car)
        obs, rew, rl, done = env.step(np.concatenate([a.astype(np.float64), light], axis=1))
        ep_rew = np.minimum(ep_rew, rl)                                 # PY:461
        buf["obs_c"][t] = state.transpose(1, 0, 2)
        buf["act"][t], buf["logp"][t], buf["rew"][t] = a.T, logp.T, rew.T
    route = np.where(exist, (action_d[:, :C] > 0).astype(np.int64), -1).T    # PY:491: action_d[i], flat index i
    buf.update(route=route, obs_d=light_f.transpose(1, 0, 2), act_d=light_a.T, logp_d=light_lp.T, rew_d=ep_rew.T,
               exist=exist.T, done=done)
    return buf
