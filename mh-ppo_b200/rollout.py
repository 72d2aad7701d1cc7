"""Env_rollout -- vectorised mirror of the reference's rollout class (Coop-MH-PPO-scalable.py:96-684)
for the scalable env class: every env runs one 80-step episode per call, the per-(car, pedestrian)
feature builders, policy inference, sampling, log-probs and buffer writes run in the kernels of
libmhppo_b200.so, and the env-step kernel writes rewards / reward_light straight into the rollout
buffers (zero copy).

Buffers (CUDA, fp32, env index innermost; include/mhppo.h "Buffers"):
  obs_c [13][S], act/logp/rew/rl/rtg [S] with S = T*C*N and sample s = (t*C + car)*N + env
  obs_d [D][M], act_d/logp_d/rew_d [M] with M = C*N;   route int8 [C][N]: 0 cross, 1 wait, -1 car absent
"""
import ctypes as C

import torch

from . import _lib
from ._lib import RolloutCfg, View, check


class Env_rollout:
    def __init__(self, env, nb_cars, max_steps, dt, seed=None):
        if env.variant != "coop_scalable":
            raise NotImplementedError("the PPO rollout kernels implement Coop-MH-PPO-scalable.py (scalable env class)")
        self.env, self.nb_cars, self.max_steps, self.dt = env, nb_cars, max_steps, dt
        self._L = _lib.lib()
        N, Cn, P, T = env.n_envs, env.n_slots, env.nb_ped, max_steps
        self.N, self.C, self.P, self.T = N, Cn, P, T
        self.shape_env = 2 + 9 + 2                       # PY:110
        self.shape_env_d = 2 + 6 * (Cn - 1) + 8 + 2      # PY:111
        dev = env.device
        S, M = T * Cn * N, Cn * N
        self.S, self.M = S, M
        f = lambda *s: torch.zeros(*s, device=dev)
        self.obs_c, self.act, self.logp, self.rew, self.rl, self.rtg, self.V = f(13, S), f(S), f(S), f(S), f(S), f(S), f(S)
        self.obs_d, self.act_d, self.logp_d, self.rew_d, self.V_d = f(self.shape_env_d, M), f(M), f(M), f(M), f(M)
        self.action_d = torch.zeros(Cn * P, N, dtype=torch.int8, device=dev)
        self.light = f(Cn, N)
        self.actions = f(2 * Cn, N)                      # env action buffer, component-major
        self.route = torch.zeros(Cn, N, dtype=torch.int8, device=dev)
        self.exist = torch.zeros(Cn, N, dtype=torch.int8, device=dev)
        self._cfg = RolloutCfg(nb_ped=P, nb_lines=env.nb_lines, T=T, n_envs=N, seed=env._seed, env_id0=env._env_id0)
        self.iteration = 0

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.env.device).cuda_stream)

    def reset(self):
        """Reset (PY:125-150): a fresh episode in every env."""
        self.env.reset()

    def iterations_rand(self, actor_net_cross, actor_net_wait, actor_net_choice, cov_mat=None, cov_mat_d=None, batch_size=None,
                        random_rate=0.0):
        """One episode per env of the reference's rollout (PY:357-516).  Resets the envs first (PY:377)."""
        env, L, cfg, st = self.env, self._L, self._cfg, self._stream()
        N, Cn, T = self.N, self.C, self.T
        env.reset()
        self.exist.copy_(env.get_state()["car_i"][:, :, 1].t().to(torch.int8))               # cars_exist, PY:380
        check(L.mhppo_choice_act(C.byref(cfg), env._obs.data_ptr(), actor_net_choice.flat.data_ptr(), self.iteration,
                                 self.action_d.data_ptr(), self.light.data_ptr(), self.obs_d.data_ptr(), self.act_d.data_ptr(),
                                 self.logp_d.data_ptr(), st))
        was_auto = env.autoreset
        env.autoreset = False                           # the episode ends exactly at step T-1; the next call resets
        none = View(None, 0, 0)
        for t in range(T):
            check(L.mhppo_policy_act(C.byref(cfg), env._obs.data_ptr(), actor_net_cross.flat.data_ptr(),
                                     actor_net_wait.flat.data_ptr(), self.action_d.data_ptr(), self.light.data_ptr(), t,
                                     self.iteration, self.actions.data_ptr(), self.obs_c.data_ptr(), self.act.data_ptr(),
                                     self.logp.data_ptr(), st))
            off = t * Cn * N * 4
            check(L.mhppo_env_step(env._h, View(self.actions.data_ptr(), 1, N), View(env._obs.data_ptr(), 1, N),
                                   View(self.rew.data_ptr() + off, 1, N), View(self.rl.data_ptr() + off, 1, N),
                                   env._done.data_ptr(), 0, none, st))
        env.autoreset = was_auto
        # trajectory routing (PY:489-502): existing car i goes to the cross buffer if action_d[i] <= 0, where i indexes the
        # flat (car x ped) decision array -- the reference's own quirk (SURVEY.md hard part 6) -- else to the wait buffer
        flat_d = self.action_d[:Cn]                      # rows 0..C-1 of the (car*P + ped) array
        self.route.copy_(torch.where(self.exist != 0, (flat_d > 0).to(torch.int8), torch.full_like(flat_d, -1)))
        self.iteration += 1

    def futur_rewards(self):
        """Expected future rewards (PY:658-684) for every trajectory + the episodic choice reward (PY:461, 504)."""
        check(self._L.mhppo_returns(self.rew.data_ptr(), self.rl.data_ptr(), self.T, self.M, 0.99, self.rtg.data_ptr(),
                                    self.rew_d.data_ptr(), self._stream()))
        return self.rtg, self.rew_d

    def counts(self):
        """(cross, wait, choice) sample counts of the last rollout."""
        ncross = int((self.route == 0).sum().item()) * self.T
        nwait = int((self.route == 1).sum().item()) * self.T
        return ncross, nwait, int((self.exist != 0).sum().item())

    def immediate_rewards(self):
        """Mean immediate rewards per buffer (PY:635-656, reduced on device)."""
        r = self.rew.view(self.T, self.C, self.N)
        mc, mw = (self.route == 0), (self.route == 1)
        cross = r[:, mc].mean() if mc.any() else None
        wait = r[:, mw].mean() if mw.any() else None
        choice = self.rew_d.view(self.C, self.N)[self.exist != 0].mean()
        return cross, wait, choice
