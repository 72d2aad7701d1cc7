"""VecCrosswalkEnv -- vectorised drop-in for the reference's gym env classes.

Mirrors the constructor and the reset/step/observation/reward API of
`Crosswalk_hybrid_multi_{stop,naif,coop,coop_4cars,coop_4cars2,coop_scalable}`
(reference Environments/Env_hybrid_multi_*.py; ids registered in Environments/__init__.py:3-41),
batched on a leading env axis, with the state resident in HBM and every step executed by the
sm_100a kernels of libmhppo_b200.so.

  reference                                            here
  gym.make(id, car_b=.., ped_b=.., cross_b=.., nb_car=.., nb_ped=.., nb_lines=.., dt=.., max_episode=..,
           simulation=..)                              make(id, n_envs, same kwargs, seed=, device=)
  env.reset() -> (dict{car,(car_follow),env,ped}, {})  reset() -> (dict of f32 CUDA tensors [N,.], {})
  env.step(a[2C]) -> (dict, rew[C], done, False, {})   step(a f32[N,A]) -> (dict, rew[N,C], done[N], trunc[N], {})
  env.reward_light (SC:846)                            .reward_light  f32[N,C]
  env.cars[i].exist (PY:380)                           .car_exist     bool[N,C]   (from get_state())
  env.pedestrian[i].waiting_time (PY:235)              .ped_waiting_time f32[N,P]

The observation dict holds views into one flat [n_obs, N] device buffer (coalesced kernel stores);
`flat_obs` is the [N, n_obs] view in gym-0.26 Dict key order car,(car_follow),env,ped, which is the
order the reference rollout slices (PY:549-554).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import EnvCfg, EnvDims, View, check

ENV_IDS = {
    "Crosswalk_hybrid_multi_stop-v0": "stop", "Crosswalk_hybrid_multi_naif-v0": "naif",
    "Crosswalk_hybrid_multi_coop-v0": "coop", "Crosswalk_hybrid_multi_coop_4cars-v0": "coop_4cars",
    "Crosswalk_hybrid_multi_coop_4cars2-v0": "coop_4cars2",
    "Crosswalk_hybrid_multi_coop_scalable-v0": "coop_scalable",
}
VARIANT_ID = {"stop": 0, "naif": 1, "coop": 2, "coop_4cars": 3, "coop_4cars2": 4, "coop_scalable": 5}

# constants of the reference drivers (PY:1007-1028)
CAR_B = np.array([[-4.0, 10.0], [2.0, 10.0]])
PED_B = np.array([[-0.05, 0.75, 0.0, -3.0], [0.05, 1.75, 4.0, -0.5]])
CROSS_B = np.array([2.5, 3.0])


def _view(t, soa):
    """mhppo_view of a 2-D fp32 CUDA tensor; soa: tensor is [width, N] (component-major)."""
    if t is None:
        return View(None, 0, 0)
    assert t.dtype == torch.float32 and t.is_cuda and t.dim() == 2
    if soa:
        return View(t.data_ptr(), t.stride(1), t.stride(0))
    return View(t.data_ptr(), t.stride(0), t.stride(1))


class VecCrosswalkEnv:
    def __init__(self, variant, n_envs, car_b=CAR_B, ped_b=PED_B, cross_b=CROSS_B, nb_car=1, nb_ped=1, nb_lines=1,
                 dt=0.3, max_episode=80, simulation="sin", seed=0, env_id0=0, device=None, autoreset=True):
        variant = ENV_IDS.get(variant, variant)
        if variant not in VARIANT_ID:
            raise ValueError("unknown env %r" % (variant,))
        self._L = _lib.lib()                      # raises LibraryMissing if the CUDA library is not built
        if not torch.cuda.is_available():
            raise _lib.MhppoError("no CUDA device: mhppo_b200 has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else torch.device(device).index or 0)
        self.variant, self.n_envs = variant, int(n_envs)
        # attributes the reference rollout reads (PY:165-205, 708)
        self.nb_car, self.nb_ped, self.nb_lines, self.dt, self.max_episode = nb_car, nb_ped, nb_lines, dt, max_episode
        self.car_b, self.ped_b, self.cross_b = np.asarray(car_b, float), np.asarray(ped_b, float), np.asarray(cross_b, float)
        self.speed_limit = 10                      # SC:892
        self.simulation, self.autoreset = simulation, bool(autoreset)
        self._seed, self._env_id0 = int(seed), int(env_id0)
        cfg = EnvCfg(variant=VARIANT_ID[variant], nb_car=nb_car, nb_ped=nb_ped, nb_lines=nb_lines, max_episode=max_episode,
                     sin_model=int(simulation == "sin"), device=self.device.index, dt=dt, seed=seed, n_envs=self.n_envs,
                     env_id0=env_id0)
        cfg.car_b[:] = list(self.car_b.ravel()); cfg.ped_b[:] = list(self.ped_b.ravel()); cfg.cross_b[:] = list(self.cross_b.ravel())
        self._h = C.c_void_p()
        check(self._L.mhppo_env_create(C.byref(cfg), C.byref(self._h)))
        d = EnvDims()
        check(self._L.mhppo_env_get_dims(self._h, C.byref(d)))
        self.n_slots, self.n_lead, self.n_action, self.n_obs, self.done_step = d.n_slots, d.n_lead, d.n_action, d.n_obs, d.done_step
        N, dev = self.n_envs, self.device
        # component-major device buffers: coalesced stores from the kernel, [N, .] views for the user
        self._obs = torch.zeros(self.n_obs, N, device=dev)
        self._term = torch.zeros(self.n_obs, N, device=dev)
        self._rew = torch.zeros(self.n_lead, N, device=dev)
        self._rl = torch.zeros(self.n_lead, N, device=dev)
        self._done = torch.zeros(N, dtype=torch.uint8, device=dev)
        self._trunc = torch.zeros(N, dtype=torch.bool, device=dev)
        car_w = 7 if variant == "coop_scalable" else 6
        env_w = 4 if variant == "coop_scalable" else 3
        n_follow = nb_car if "4cars" in variant else 0
        n_car_rows = self.n_slots - n_follow
        o, self._slices = 0, {}
        self._slices["car"] = (o, o + car_w * n_car_rows); o += car_w * n_car_rows
        if n_follow:
            self._slices["car_follow"] = (o, o + car_w * n_follow); o += car_w * n_follow
        self._slices["env"] = (o, o + env_w); o += env_w
        self._slices["ped"] = (o, o + 9 * nb_ped); o += 9 * nb_ped
        assert o == self.n_obs
        self.observation_shapes = {k: (b - a,) for k, (a, b) in self._slices.items()}

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.mhppo_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _obs_dict(self, buf):
        return {k: buf[a:b].t() for k, (a, b) in self._slices.items()}

    # -- gym API ----------------------------------------------------------------------------------
    @property
    def flat_obs(self):
        return self._obs.t()

    @property
    def reward_light(self):
        return self._rl.t()

    def reset(self, seed=None, options=None, mask=None):
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        check(self._L.mhppo_env_reset(self._h, None if m is None else m.data_ptr(), _view(self._obs, True), self._stream()))
        return self._obs_dict(self._obs), {}

    def step(self, actions):
        a = actions
        if not torch.is_tensor(a):
            a = torch.as_tensor(np.asarray(a), dtype=torch.float32, device=self.device)
        if a.dtype != torch.float32 or a.device != self.device:
            a = a.to(device=self.device, dtype=torch.float32)
        a = a.reshape(self.n_envs, self.n_action)
        check(self._L.mhppo_env_step(self._h, _view(a, False), _view(self._obs, True), _view(self._rew, True),
                                     _view(self._rl, True), self._done.data_ptr(), int(self.autoreset),
                                     _view(self._term, True) if self.autoreset else View(None, 0, 0), self._stream()))
        return self._obs_dict(self._obs), self._rew.t(), self._done.bool(), self._trunc, {}

    def observe(self):
        """env.get_state() of the reference (SC:960-969): the observation of the current state, e.g. after set_state /
        load_state / a state injection; refreshes the buffer the rollout kernels read."""
        check(self._L.mhppo_env_observe(self._h, _view(self._obs, True), self._stream()))
        return self._obs_dict(self._obs)

    @property
    def terminal_obs(self):
        return self._term.t()

    # -- host-buffer call (the reference's own calling convention: numpy in, numpy out) -------------
    def step_host(self, actions_host, obs_host, rewards_host, reward_light_host, done_host):
        """Row-major pinned host tensors [N, width]; H2D + kernel + D2H + stream sync inside."""
        check(self._L.mhppo_env_step_host(self._h, actions_host.data_ptr(), obs_host.data_ptr(), rewards_host.data_ptr(),
                                          reward_light_host.data_ptr(), done_host.data_ptr(), int(self.autoreset),
                                          self._stream()))

    def reset_host(self, obs_host):
        check(self._L.mhppo_env_reset_host(self._h, obs_host.data_ptr(), self._stream()))

    # -- state access (get_state / reset_pedestrian / reset_cars of the reference, SC:948-969) -------
    def get_state(self):
        N, Cn, P, dev = self.n_envs, self.n_slots, self.nb_ped, self.device
        s = dict(car_f=torch.empty(N, Cn, 7, device=dev), car_i=torch.empty(N, Cn, 2, dtype=torch.int32, device=dev),
                 ped_f=torch.empty(N, P, 9, device=dev), ped_i=torch.empty(N, P, 9, dtype=torch.int32, device=dev),
                 env_f=torch.empty(N, 1, dtype=torch.float64, device=dev), env_i=torch.empty(N, 4, dtype=torch.int64, device=dev))
        check(self._L.mhppo_env_export_state(self._h, *[s[k].data_ptr() for k in ("car_f", "car_i", "ped_f", "ped_i", "env_f", "env_i")],
                                             self._stream()))
        return s

    def set_state(self, s):
        dev = self.device
        want = dict(car_f=torch.float32, car_i=torch.int32, ped_f=torch.float32, ped_i=torch.int32, env_f=torch.float64,
                    env_i=torch.int64)
        t = {k: torch.as_tensor(np.asarray(s[k]) if not torch.is_tensor(s[k]) else s[k]).to(device=dev, dtype=dt).contiguous()
             for k, dt in want.items()}
        check(self._L.mhppo_env_import_state(self._h, *[t[k].data_ptr() for k in ("car_f", "car_i", "ped_f", "ped_i", "env_f", "env_i")],
                                             self._stream()))
        torch.cuda.current_stream(dev).synchronize()   # `t` must outlive the kernel

    # -- state injection (fixed scenarios: choix_test PY:629-633, reset_distrib NA:903-913) ----------------------------
    def _params(self, vals, width):
        a = torch.zeros(self.n_envs, width, device=self.device)
        for k, v in enumerate(vals):
            a[:, k] = torch.as_tensor(v, dtype=torch.float32, device=self.device) if not isinstance(v, (int, float, bool)) else float(v)
        return a

    def _mask(self, mask):
        return None if mask is None else torch.as_tensor(mask).to(device=self.device, dtype=torch.uint8).contiguous()

    def reset_pedestrian(self, num_ped, *vals, mask=None):
        """env.reset_pedestrian(num_ped, speed_x, speed_y, pos_x, pos_y, dl, leave, CZ, exist, direction) (SC:948-955) in the
        masked envs (default: all); every value a scalar or an [N] array.  naif: (num_ped, speed_x, speed_y, pos_x, pos_y, dl,
        direction, cross) (NA:897).  Like the reference it rebuilds every pedestrian as a placeholder first (not in naif)."""
        want = 7 if self.variant == "naif" else 9
        if len(vals) != want:
            raise TypeError("reset_pedestrian of %s takes %d values after num_ped" % (self.variant, want))
        prm, m = self._params(vals, 9), self._mask(mask)
        check(self._L.mhppo_env_reset_pedestrian(self._h, int(num_ped), prm.data_ptr(), None if m is None else m.data_ptr(), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()      # prm / m must outlive the kernel

    def reset_cars(self, num_car, speed_x, pos_x, light, line, mask=None):
        """env.reset_cars(num_car, speed_x, pos_x, light, line) (SC:957-958)."""
        prm, m = self._params((speed_x, pos_x, light, line), 4), self._mask(mask)
        check(self._L.mhppo_env_reset_cars(self._h, int(num_car), prm.data_ptr(), None if m is None else m.data_ptr(), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()

    def reset_distrib(self, state_distrib, mask=None):
        """naif's reset_distrib (NA:903-913): state_distrib [N, 5*nb_ped + nb_car + 1] = per pedestrian (speed_x, speed_y, pos_x,
        pos_y, direction sign), per car its position, then the crossing width; cars restart at 10 m/s, light 0, line = index."""
        if self.variant != "naif":
            raise NotImplementedError("reset_distrib exists on the naif env class only (NA:903)")
        sd = torch.as_tensor(state_distrib, dtype=torch.float32, device=self.device).reshape(self.n_envs, -1)
        P, Cn = self.nb_ped, self.nb_car
        cross = sd[:, Cn + 5 * P]
        for i in range(P):
            d = torch.where(sd[:, 5 * i + 4] >= 0, 1.0, -1.0)
            self.reset_pedestrian(i, sd[:, 5 * i], sd[:, 5 * i + 1], sd[:, 5 * i + 2], sd[:, 5 * i + 3], 0.0, d, cross, mask=mask)
        for i in range(Cn):
            self.reset_cars(i, 10.0, sd[:, i + 5 * P], 0.0, float(i), mask=mask)

    # -- scenario snapshots (the reference pickles copy.deepcopy(env) per scenario, PY:255-277, 1873-1883) ----------
    def save_state(self, path):
        """Write the canonical state dump of every env (include/mhppo.h "state dump") plus the constructor arguments to
        one .npz file: the flat struct-of-arrays counterpart of the reference's pickled env lists."""
        s = {k: v.cpu().numpy() for k, v in self.get_state().items()}
        np.savez_compressed(path, variant=self.variant, nb_car=self.nb_car, nb_ped=self.nb_ped, nb_lines=self.nb_lines,
                            n_envs=self.n_envs, dt=self.dt, max_episode=self.max_episode, seed=self._seed, env_id0=self._env_id0,
                            car_b=self.car_b, ped_b=self.ped_b, cross_b=self.cross_b, **s)

    def load_state(self, path):
        """Restore a snapshot written by save_state into this (same-shaped) env set; returns the current observation."""
        z = np.load(path)
        for k in ("nb_car", "nb_ped", "nb_lines", "n_envs"):
            if int(z[k]) != getattr(self, k):
                raise ValueError("snapshot %s=%d does not match this env (%d)" % (k, int(z[k]), getattr(self, k)))
        if str(z["variant"]) != self.variant:
            raise ValueError("snapshot is of env class %s, this is %s" % (z["variant"], self.variant))
        self.set_state({k: z[k] for k in ("car_f", "car_i", "ped_f", "ped_i", "env_f", "env_i")})
        return self.observe()

    @property
    def car_exist(self):
        return self.get_state()["car_i"][:, :, 1].bool()

    @property
    def ped_waiting_time(self):
        return self.get_state()["ped_i"][:, :, 1].float() * self.dt

    @property
    def state_bytes_per_env(self):
        return int(self._L.mhppo_env_state_bytes_per_env(self._h))


def make(id, n_envs, **kwargs):
    """gym.make(id, **kwargs) of the reference, vectorised over n_envs."""
    return VecCrosswalkEnv(id, n_envs, **kwargs)
