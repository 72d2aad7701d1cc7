"""Drive the UNMODIFIED reference rollout (Env_rollout.iterations_rand, PY:357-516) with injected
sampling noise, so the vectorised implementation can be compared sample for sample (build container
only; TEST INFRASTRUCTURE).

The reference draws its exploration noise from torch's global generator through
torch.distributions.MultivariateNormal / Categorical.  The names `MultivariateNormal` and
`Categorical` in the loaded script namespace are replaced by thin stand-ins that take their noise from
the project's policy-noise contract (DESIGN.md "RNG": Philox blocks keyed like the env stream, stream
id 1 = continuous head, 2 = discrete head) and delegate log_prob to the real torch classes.
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
from philox import philox4x32_10, u53  # noqa: E402


def policy_normal(seed, env_id, index, iteration=0):
    w = philox4x32_10((index & 0xffffffff, 1 | (iteration << 8), env_id & 0xffffffff, env_id >> 32), (seed & 0xffffffff, seed >> 32))
    return math.sqrt(-2.0 * math.log(1.0 - u53(w[0], w[1]))) * math.cos(2.0 * math.pi * u53(w[2], w[3]))


def policy_uniform(seed, env_id, index, iteration=0):
    w = philox4x32_10((index & 0xffffffff, 2 | (iteration << 8), env_id & 0xffffffff, env_id >> 32), (seed & 0xffffffff, seed >> 32))
    return u53(w[0], w[1])


class NoiseFeed:
    def __init__(self, seed, env_id, iteration=0):
        self.seed, self.env_id, self.iteration = seed, env_id, iteration
        self.kc = 0
        self.kd = 0

    def normal(self):
        z = policy_normal(self.seed, self.env_id, self.kc, self.iteration)
        self.kc += 1
        return z

    def uniform(self):
        u = policy_uniform(self.seed, self.env_id, self.kd, self.iteration)
        self.kd += 1
        return u


def patch_distributions(ns, feed):
    real_mvn, real_cat = torch.distributions.MultivariateNormal, torch.distributions.Categorical

    class MVN:
        def __init__(self, mean, cov):
            self.d = real_mvn(mean, cov)
            self.mean = mean

        def sample(self):
            return (self.mean + math.sqrt(0.5) * torch.tensor([feed.normal()], dtype=torch.float32)).float()

        def log_prob(self, a):
            return self.d.log_prob(a)

    class Cat:
        def __init__(self, probs):
            self.d = real_cat(probs)
            self.probs = probs

        def sample(self):
            return torch.tensor(1 if np.float32(feed.uniform()) >= self.probs[0].item() else 0)

        def log_prob(self, a):
            return self.d.log_prob(a)

    ns["MultivariateNormal"], ns["Categorical"] = MVN, Cat
    return real_mvn, real_cat


def reference_episode(ns, algo, env, seed, env_id, iteration=0):
    """One episode of the reference rollout on stream (seed, env_id); returns its buffers as arrays."""
    feed = NoiseFeed(seed, env_id, iteration)
    real = patch_distributions(ns, feed)
    try:
        r = algo.rollout
        r.reset()
        env._mh_rng.set_stream(seed, env_id, 0)
        r.iterations_rand(algo.actor_net_cross, algo.actor_net_wait, algo.actor_net_choice, algo.cov_mat, algo.cov_mat_d, 1, 0.0)
        rt = r.futur_rewards()
        out = {}
        for name in ("cross", "wait", "choice"):
            out["obs_" + name] = np.array(getattr(r, "batch_obs_" + name), np.float32).reshape(-1, 13 if name != "choice" else r.shape_env_d)
            out["acts_" + name] = np.array(getattr(r, "batch_acts_" + name), np.float64).reshape(-1)
            out["logp_" + name] = np.array(getattr(r, "batch_log_probs_" + name), np.float64).reshape(-1)
            out["rews_" + name] = np.array(getattr(r, "batch_rews_" + name), np.float64).reshape(-1)
        out["rtg_cross"], out["rtg_wait"], out["rtg_choice"] = [x.numpy().reshape(-1) for x in rt]
        out["car_exist"] = np.array([getattr(c, "exist", True) for c in env.cars])
        return out
    finally:
        ns["MultivariateNormal"], ns["Categorical"] = real


def reference_eval_episode(ns, algo, env, seed, env_id):
    """One episode of the reference's deterministic evaluation rollout (Env_rollout.iterations, PY:152-252) on stream
    (seed, env_id).  `iterations` keeps nothing per step but its return values, so env.step is wrapped to record the
    actions it receives, the rewards, reward_light and the pedestrians' waiting times."""
    r = algo.rollout
    env._mh_rng.set_stream(seed, env_id, 0)
    rec = dict(actions=[], rew=[], rl=[], waiting=[])
    real_step = env.step

    def step(a):
        out = real_step(a)
        rec["actions"].append(np.array(a, np.float64)); rec["rew"].append(np.array(out[1], np.float64))
        rec["rl"].append(np.array(env.reward_light, np.float64))
        rec["waiting"].append(np.array([p.waiting_time for p in env.pedestrian], np.float64))
        return out

    env.step = step
    try:
        obs, acts, rews_c, rews_d, wt = r.iterations(algo.actor_net_cross, algo.actor_net_wait, algo.actor_net_choice, 1)
    finally:
        del env.step
    out = {k: np.stack(v) for k, v in rec.items()}
    out.update(obs=obs.numpy(), acts=acts.numpy(), rews_c=rews_c.numpy().reshape(len(rec["rew"]), -1), rews_d=rews_d.numpy(),
               waiting_batch=wt.numpy())
    return out
