"""Algo_PPO -- vectorised mirror of the reference's PPO class (Coop-MH-PPO-scalable.py:698-1001) for the
scalable env class: same constructor (`Algo_PPO(policy_class, env, **hyperparameters)`), same six networks
and six Adam optimisers, same update rule (clipped surrogate 0.8/1.2, no entropy term, MSE critic, 10 full-batch
epochs per net per iteration, advantages = reward-to-go - V normalised with the unbiased std).

One iteration = one 80-step episode in every env of the vectorised env.  Multi-GPU: one process per GPU, envs
sharded per rank; the only collectives are the all-reduce of the three advantage statistics and of the flat
gradient buffers (NCCL through torch.distributed), so an R-rank run performs the same update as one rank
holding all envs.
"""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import check
from .policy import Adam, Model_PPO
from .rollout import Env_rollout


_USE_DIST = True


def set_distributed(on):
    """Tests only: run a single-rank reference pass inside an initialised process group."""
    global _USE_DIST
    _USE_DIST = bool(on)


def _dist():
    import torch.distributed as dist
    return dist if (_USE_DIST and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None


def combine_stats(stats3):
    """(sum A, sum A^2, n) -> (mean, 1/(unbiased std + 1e-10), n): advantage normalisation of PY:787 from partial sums."""
    sA, sAA, n = float(stats3[0]), float(stats3[1]), float(stats3[2])
    if n < 1:
        return 0.0, 0.0, 0.0
    mean = sA / n
    var = max(sAA - n * mean * mean, 0.0) / (n - 1) if n > 1 else float("nan")
    std = math.sqrt(var) if n > 1 else float("nan")
    return mean, 1.0 / (std + 1e-10), n


def allreduce_sum_(t):
    d = _dist()
    if d is not None:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    return t


class Algo_PPO:
    def __init__(self, policy_class, env, **hyperparameters):
        self._init_hyperparameters(hyperparameters)
        self.env = env
        self.max_steps = env.max_episode
        dev = env.device
        nc = 2 * env.nb_lines
        mk = lambda *a, **k: policy_class(*a, device=dev, **k)
        self.actor_net_cross = mk(self.num_states_c, self.num_actions, 1, nb_car=nc, mean=self.mean, std=self.std)      # PY:711-718
        self.actor_net_wait = mk(self.num_states_c, self.num_actions, 1, nb_car=nc, mean=self.mean, std=self.std)
        self.actor_net_choice = mk(self.num_states_d, 2, 2)
        self.critic_net_cross = mk(self.num_states_c, 1, 0)
        self.critic_net_wait = mk(self.num_states_c, 1, 0)
        self.critic_net_choice = mk(self.num_states_d, 1, 0)
        self.optimizer_critic_cross = Adam(self.critic_net_cross, self.critic_lr)                                        # PY:719-724
        self.optimizer_critic_wait = Adam(self.critic_net_wait, self.critic_lr)
        self.optimizer_critic_choice = Adam(self.critic_net_choice, self.critic_d_lr)
        self.optimizer_actor_cross = Adam(self.actor_net_cross, self.actor_lr)
        self.optimizer_actor_wait = Adam(self.actor_net_wait, self.actor_lr)
        self.optimizer_actor_choice = Adam(self.actor_net_choice, self.actor_d_lr)
        if not hasattr(self, "value_std"):
            self.value_std = 0.5                                                                                         # PY:726
        self.value_std_d = 0.1                                                                                           # PY:727 (unused by the reference's Categorical)
        self.rollout = Env_rollout(env, env.n_slots, self.max_steps, self.dt)
        if self.num_states_d != self.rollout.shape_env_d:
            raise ValueError("num_states_d = %s, but the choice features of this env have %d columns" % (self.num_states_d, self.rollout.shape_env_d))
        self.rollout.value_std = self.value_std
        self.ep_reward_cross, self.ep_reward_wait, self.ep_reward_choice, self.ep_scenario_balance = [], [], [], []
        L = _lib.lib()
        nbytes = max(L.mhppo_update_workspace_bytes(self.num_states_c), L.mhppo_update_workspace_bytes(self.num_states_d))
        self._ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        self._stats = torch.zeros(3, dtype=torch.float64, device=dev)
        self._loss = torch.zeros(2, dtype=torch.float64, device=dev)
        self.last_losses = {}
        self._sel_cache = {}
        # one contiguous gradient buffer per (actor, critic) pair: a single all-reduce per epoch carries both
        self._gbuf = {}
        for name, actor, critic in (("cross", self.actor_net_cross, self.critic_net_cross), ("wait", self.actor_net_wait, self.critic_net_wait),
                                    ("choice", self.actor_net_choice, self.critic_net_choice)):
            g = torch.zeros(actor.n_param + critic.n_param, device=dev)
            actor.grad, critic.grad = g[:actor.n_param], g[actor.n_param:]
            self._gbuf[id(actor)] = g
        self.sync_parameters()

    def sync_parameters(self):
        """Multi-GPU: every rank must hold rank 0's parameters and optimiser state (only gradients and advantage statistics
        are all-reduced afterwards, so the replicas then stay identical).  Called after construction and loading()."""
        d = _dist()
        if d is None:
            return
        for _, _, net in self._nets():
            for t in (net.flat, net.m, net.v):
                d.broadcast(t, src=0)
            step = torch.tensor([net.step], dtype=torch.int64, device=net.flat.device)
            d.broadcast(step, src=0)
            net.step = int(step.item())

    def _init_hyperparameters(self, hyperparameters):
        """PY:919-933."""
        self.num_algo, self.total_loop, self.batch_size, self.gamma = 1, 0, 2048, 0.99
        self.critic_lr, self.actor_lr, self.critic_d_lr, self.actor_d_lr = 1e-3, 3e-4, 1e-3, 3e-4
        self.num_states_c, self.num_states_d, self.num_actions, self.mean, self.std, self.dt = 13, None, 1, -1.0, 3.0, 0.3
        for k, v in hyperparameters.items():
            setattr(self, k, v)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.env.device).cuda_stream)

    # -- one epoch of train_model_c (PY:778-815) on the cross (want=0) or wait (want=1) buffer ------------------
    def _selection(self, want):
        """Compacted (car, env) columns routed to the cross (0) / wait (1) buffer, or the existing cars (want=None)."""
        r = self.rollout
        key = (r.iteration, r.route.data_ptr(), int(r.route._version), int(r.exist._version))
        if self._sel_cache.get("key") != key:
            self._sel_cache = {"key": key}
        if want not in self._sel_cache:
            mask = (r.exist.view(-1) != 0) if want is None else (r.route.view(-1) == want)
            sel = torch.nonzero(mask).view(-1).to(torch.int32)
            self._sel_cache[want] = (sel if sel.numel() else torch.zeros(1, dtype=torch.int32, device=mask.device), sel.numel())
        return self._sel_cache[want]

    def _global_count(self, want, local):
        """Samples of this buffer over all ranks (the denominators of the losses, PY:806-809); the selection is fixed for
        the 10 epochs of an update, so the all-reduce happens once per buffer."""
        k = ("n", want)
        if k not in self._sel_cache:
            t = torch.tensor([float(local)], dtype=torch.float64, device=self.rollout.route.device)
            self._sel_cache[k] = int(allreduce_sum_(t).item())
        return self._sel_cache[k]

    def train_model_c(self, actor, critic, opti_actor, opti_critic, want):
        r, L, st = self.rollout, _lib.lib(), self._stream()
        ws = self._ws.data_ptr()
        idx, K = self._selection(want)                                      # K == 0 still runs (zero partials): other ranks may have samples
        n = self._global_count(want, K * (r.S // r.M))             # sample slots: S = T * M, M = C * N columns
        if n < 1:
            return False                                     # PY:869/874: the net is skipped when its buffer is empty
        # critic pass first: its forward is V = critic(s) of PY:785, so the same kernel returns V and the advantage statistics
        check(L.mhppo_critic_grad_stats(13, r.obs_c.data_ptr(), 13, r.S, idx.data_ptr(), K, r.M, critic.flat.data_ptr(),
                                        r.rtg.data_ptr(), 1.0 / n, critic.grad.data_ptr(), self._loss.data_ptr() + 8,
                                        r.V.data_ptr(), self._stats.data_ptr(), ws, st))
        allreduce_sum_(self._stats)                          # (sum A, sum A^2, n) over all ranks; normalised inside the actor kernel
        check(L.mhppo_ppo_grad_dev(13, 1, r.obs_c.data_ptr(), 13, r.S, idx.data_ptr(), K, r.M, actor.flat.data_ptr(),
                                   r.act.data_ptr(), r.logp.data_ptr(), r.rtg.data_ptr(), r.V.data_ptr(), self._stats.data_ptr(), 1.0 / n,
                                   0.0, 0.0, actor.grad.data_ptr(), self._loss.data_ptr(), ws, st))
        self._step_pair(actor, critic, opti_actor, opti_critic)
        return True

    def _step_pair(self, actor, critic, opti_actor, opti_critic):
        """All-reduce the pair's packed gradients (one collective) and take both Adam steps in one launch (PY:810-815)."""
        allreduce_sum_(self._gbuf[id(actor)])
        actor.step += 1; critic.step += 1
        check(_lib.lib().mhppo_adam2(actor.flat.data_ptr(), actor.grad.data_ptr(), actor.m.data_ptr(), actor.v.data_ptr(), actor.n_param,
                                     opti_actor.lr, actor.step, critic.flat.data_ptr(), critic.grad.data_ptr(), critic.m.data_ptr(),
                                     critic.v.data_ptr(), critic.n_param, opti_critic.lr, critic.step, 0.9, 0.999, 1e-8, self._stream()))

    # -- one epoch of train_model_d (PY:818-851) on the choice buffer -------------------------------------------
    def train_model_d(self, actor, critic, opti_actor, opti_critic):
        r, L, st = self.rollout, _lib.lib(), self._stream()
        ws, D = self._ws.data_ptr(), r.shape_env_d
        idx, K = self._selection(None)
        n = self._global_count(None, K)
        if n < 1:
            return False
        check(L.mhppo_critic_grad_stats(D, r.obs_d.data_ptr(), D, r.M, idx.data_ptr(), K, r.M, critic.flat.data_ptr(),
                                        r.rew_d.data_ptr(), 1.0 / n, critic.grad.data_ptr(), self._loss.data_ptr() + 8,
                                        r.V_d.data_ptr(), self._stats.data_ptr(), ws, st))
        allreduce_sum_(self._stats)
        f0, f1 = self._action_fractions(n)
        check(L.mhppo_ppo_grad_dev(D, 2, r.obs_d.data_ptr(), D, r.M, idx.data_ptr(), K, r.M, actor.flat.data_ptr(),
                                   r.act_d.data_ptr(), r.logp_d.data_ptr(), r.rew_d.data_ptr(), r.V_d.data_ptr(), self._stats.data_ptr(),
                                   1.0 / n, f0, f1, actor.grad.data_ptr(), self._loss.data_ptr(), ws, st))
        self._step_pair(actor, critic, opti_actor, opti_critic)
        return True

    def _action_fractions(self, n):
        """Fractions of the choice samples (all ranks) whose action is 0 / 1: the (M, M) broadcast of PY:834-842 collapses to
        them.  Fixed for the epochs of an update: counted once per rollout."""
        if "f01" not in self._sel_cache:
            r = self.rollout
            ex = r.exist.view(-1) != 0
            cnt = torch.stack([(ex & (r.act_d == 0)).sum(), (ex & (r.act_d == 1)).sum()]).double()
            cnt = allreduce_sum_(cnt).cpu()
            self._sel_cache["f01"] = (float(cnt[0]) / n, float(cnt[1]) / n)
        return self._sel_cache["f01"]

    def update(self, epochs=10):
        """The update half of train() (PY:866-882): 10 x (cross, wait) then 10 x choice."""
        self.rollout.value_std = self.value_std
        self.rollout._set_head(self.actor_net_cross)
        self.rollout.futur_rewards()
        for _ in range(epochs):
            self.train_model_c(self.actor_net_cross, self.critic_net_cross, self.optimizer_actor_cross, self.optimizer_critic_cross, 0)
            self.train_model_c(self.actor_net_wait, self.critic_net_wait, self.optimizer_actor_wait, self.optimizer_critic_wait, 1)
        for _ in range(epochs):
            self.train_model_d(self.actor_net_choice, self.critic_net_choice, self.optimizer_actor_choice, self.optimizer_critic_choice)
        if _lib.lib().mhppo_tc_failures():              # a tensor-core kernel gave up waiting for its MMAs: its gradients are garbage
            raise RuntimeError("a tcgen05 kernel of the update timed out on its mbarrier; the parameters of this update are not "
                               "trustworthy (reload a checkpoint; MHPPO_MLP=ffma selects the CUDA-core kernels)")

    # file-name stems of the three drivers: PY:908 / NB2 (coop) / NB1 (naif, "acc6")
    _STEM = {"coop_scalable": "pappo-scalable-coop", "coop": "pappo-coop", "naif": "pappo-acc6", "stop": "pappo-acc6"}
    _TRACE = "load_model/parameters/{stem}-{num_algo:02d}-{name}-step-{epoch:03d}000.npy"

    def train(self, nb_loop, verbose=False, root=".", save_traces=True):
        """Training loop (PY:854-917); ends by writing the four learning-curve traces with the reference's file names
        (PY:908-916) under `root`, so a run can be laid next to the reference's load_model/parameters/*.npy."""
        for ep in range(nb_loop):
            self.rollout.iterations_rand(self.actor_net_cross, self.actor_net_wait, self.actor_net_choice)
            self.update()
            cross, wait, choice, (nc, nw, nd) = self._global_rewards()
            if cross is not None:
                self.ep_reward_cross.append(float(cross))
            if wait is not None:
                self.ep_reward_wait.append(float(wait))
            self.ep_reward_choice.append(float(choice))
            self.ep_scenario_balance.append([nc, nw])
            self.total_loop += 1
            if verbose:
                print("Episode * {} * cross {:.4f} wait {:.4f} choice {:.4f}  (#cross {} #wait {})".format(
                    ep, np.mean(self.ep_reward_cross[-10:] or [0]), np.mean(self.ep_reward_wait[-10:] or [0]),
                    np.mean(self.ep_reward_choice[-10:]), nc, nw))
        d = _dist()
        if save_traces and (d is None or d.get_rank() == 0):
            for name, arr in (("reward_cross", np.array(self.ep_reward_cross)), ("reward_wait", np.array(self.ep_reward_wait)),
                              ("reward_choice", np.array(self.ep_reward_choice)),
                              ("scenario_balance", np.array(self.ep_scenario_balance).reshape((-1, 2)))):
                path = os.path.join(root, self._TRACE.format(stem=self._STEM[self.env.variant], num_algo=self.num_algo,
                                                             epoch=int(self.total_loop / 1000), name=name))
                os.makedirs(os.path.dirname(path), exist_ok=True)
                np.save(path, arr)

    def _global_rewards(self):
        """Mean immediate rewards per buffer and the sample counts (PY:885-906), over the envs of ALL ranks."""
        r = self.rollout
        rew = r.rew.view(r.T, r.C, r.N)
        mc, mw, ex = (r.route == 0), (r.route == 1), (r.exist != 0)
        t = torch.stack([rew[:, mc].sum().double(), mc.sum().double() * r.T, rew[:, mw].sum().double(), mw.sum().double() * r.T,
                         r.rew_d.view(r.C, r.N)[ex].sum().double(), ex.sum().double()])
        t = allreduce_sum_(t).cpu()
        cross = float(t[0] / t[1]) if t[1] > 0 else None
        wait = float(t[2] / t[3]) if t[3] > 0 else None
        choice = float(t[4] / t[5]) if t[5] > 0 else float("nan")
        return cross, wait, choice, (int(t[1]), int(t[3]), int(t[5]))

    # -- checkpoints: same file names and state_dict layout as the reference (PY:935-1001) ------------------------
    _PATH = "load_model/weights/{stem}-{name}-{num_algo:02d}-{kind}-step-{epoch:03d}0.pth"

    def evaluate(self, nbr_episodes, choix=False, **kw):
        """Testing (PY:738-747): the deterministic rollout of every env for `nbr_episodes` episodes.  Returns the dense
        per-step records of Env_rollout.iterations; `self.rollout.reference_batches(out, n)` gives env n's
        (state_batch, action_batch, rew_c_batch, rew_d_batch, time_stop_batch) in the reference's shapes."""
        self.rollout.reset()
        return self.rollout.iterations(self.actor_net_cross, self.actor_net_wait, self.actor_net_choice, nbr_episodes, choix=choix, **kw)

    def evaluate_dataset(self, snapshot, **kw):
        """Testing on a stored scenario dataset (PY:749-758): one deterministic episode from every scenario of the snapshot file
        (the reference unpickles `test_data_23.pickle`, a list of deep-copied envs; here a VecCrosswalkEnv.save_state file)."""
        return self.rollout.iterations_dataset(self.actor_net_cross, self.actor_net_wait, self.actor_net_choice, snapshot, **kw)

    def _nets(self):
        return [("cross", "actor", self.actor_net_cross), ("wait", "actor", self.actor_net_wait), ("choice", "actor", self.actor_net_choice),
                ("cross", "critic", self.critic_net_cross), ("wait", "critic", self.critic_net_wait), ("choice", "critic", self.critic_net_choice)]

    def saving(self, root="."):
        for name, kind, net in self._nets():
            path = os.path.join(root, self._PATH.format(stem=self._STEM[self.env.variant], name=name, kind=kind, num_algo=self.num_algo, epoch=int(self.total_loop / 10)))
            os.makedirs(os.path.dirname(path), exist_ok=True)
            torch.save(net.state_dict(), path)

    def loading_curriculum(self, num_actor, num_algo, total_loop, root="."):
        """Load ONE (actor, critic) pair from another configuration's checkpoint (PY:957-984): num_actor 0 cross, 1 wait,
        2 choice; `num_algo` names the configuration the files were trained on (curriculum over nb_ped / nb_car / nb_lines)."""
        self.total_loop = total_loop
        name = {0: "cross", 1: "wait", 2: "choice"}[int(num_actor)]
        for n, kind, net in self._nets():
            if n == name:
                path = os.path.join(root, self._PATH.format(stem=self._STEM[self.env.variant], name=n, kind=kind, num_algo=num_algo,
                                                            epoch=int(total_loop / 10)))
                net.load_state_dict(torch.load(path, map_location="cpu"))
        self.sync_parameters()

    def loading(self, num_algo, total_loop, root="."):
        self.num_algo, self.total_loop = num_algo, total_loop
        for name, kind, net in self._nets():
            path = os.path.join(root, self._PATH.format(stem=self._STEM[self.env.variant], name=name, kind=kind, num_algo=num_algo, epoch=int(total_loop / 10)))
            net.load_state_dict(torch.load(path, map_location="cpu"))
        self.sync_parameters()
