"""ctypes wrapper of the C oracle (oracle/mhppo_oracle.c).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product package (mh-ppo_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmhppo_oracle.so")

VARIANTS = ("stop", "naif", "coop", "coop_4cars", "coop_4cars2", "coop_scalable")
VARIANT_ID = {n: i for i, n in enumerate(VARIANTS)}

CAR_B = ((-4.0, 10.0), (2.0, 10.0))
PED_B = ((-0.05, 0.75, 0.0, -3.0), (0.05, 1.75, 4.0, -0.5))
CROSS_B = (2.5, 3.0)


class Cfg(C.Structure):
    _fields_ = [("variant", C.c_int32), ("nb_car", C.c_int32), ("nb_ped", C.c_int32), ("nb_lines", C.c_int32),
                ("max_episode", C.c_int32), ("sin_model", C.c_int32), ("store_f32", C.c_int32),
                ("n_threads", C.c_int32), ("dt", C.c_double), ("car_b", C.c_double * 4),
                ("ped_b", C.c_double * 8), ("cross_b", C.c_double * 2), ("seed", C.c_uint64)]


def build(force=False):
    """Compile the oracle with the Makefile next to this file (gcc, a few seconds)."""
    src = os.path.join(_HERE, "mhppo_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(src[:-1] + "h"))):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "libmhppo_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i64 = C.c_void_p, C.c_int64
        L.mho_create.argtypes = [C.POINTER(Cfg), i64, i64, C.POINTER(vp)]
        L.mho_create.restype = C.c_int
        L.mho_destroy.argtypes = [vp]
        L.mho_reset.argtypes = [vp, vp, vp]
        L.mho_step.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, vp]
        L.mho_get_state.argtypes = [vp] * 7
        L.mho_set_state.argtypes = [vp] * 7
        L.mho_reset_pedestrian.argtypes = [vp, C.c_int, vp, vp]
        L.mho_reset_cars.argtypes = [vp, C.c_int, vp, vp]
        L.mho_observe.argtypes = [vp, vp]
        L.mho_philox.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32 * 4)]
        for f in ("mho_n_slots", "mho_n_lead", "mho_n_action", "mho_n_obs"):
            getattr(L, f).argtypes = [C.POINTER(Cfg)]
            getattr(L, f).restype = C.c_int
        _lib = L
    return _lib


def philox(ctr, env_id, seed):
    out = (C.c_uint32 * 4)()
    lib().mho_philox(ctr, env_id, seed, C.byref(out))
    return tuple(out)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleVecEnv:
    """N independent reference-semantics envs on the host (batched, multi-threaded)."""

    def __init__(self, variant, n_envs, nb_car, nb_ped, nb_lines, seed=0, env_id0=0, dt=0.3, max_episode=80,
                 simulation="sin", store_f32=False, n_threads=0, car_b=CAR_B, ped_b=PED_B, cross_b=CROSS_B):
        cfg = Cfg()
        cfg.variant = VARIANT_ID[variant]
        cfg.nb_car, cfg.nb_ped, cfg.nb_lines = nb_car, nb_ped, nb_lines
        cfg.max_episode, cfg.sin_model = max_episode, int(simulation == "sin")
        cfg.store_f32, cfg.n_threads, cfg.dt, cfg.seed = int(store_f32), n_threads, dt, seed
        cfg.car_b[:] = [v for r in car_b for v in r]
        cfg.ped_b[:] = [v for r in ped_b for v in r]
        cfg.cross_b[:] = list(cross_b)
        self.cfg, self.N, self.variant = cfg, int(n_envs), variant
        L = lib()
        self.C, self.n_lead = L.mho_n_slots(C.byref(cfg)), L.mho_n_lead(C.byref(cfg))
        self.n_action, self.n_obs = L.mho_n_action(C.byref(cfg)), L.mho_n_obs(C.byref(cfg))
        self.P = nb_ped
        self._h = C.c_void_p()
        rc = L.mho_create(C.byref(cfg), self.N, env_id0, C.byref(self._h))
        if rc != 0:
            raise ValueError("mho_create failed: %d" % rc)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().mho_destroy(self._h)
            self._h = None

    def reset(self, mask=None):
        obs = np.empty((self.N, self.n_obs), np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().mho_reset(self._h, _ptr(m), _ptr(obs))
        return obs

    def step_into(self, actions, obs, rew, rl, done, autoreset=True):
        """env.step into caller-owned buffers (float64 [N, A] actions, float32 obs, float64 rewards / reward_light,
        uint8 done): the timed loop of bench.py's CPU legs, no allocation per step."""
        lib().mho_step(self._h, _ptr(actions), _ptr(obs), _ptr(rew), _ptr(rl), _ptr(done), int(autoreset), None)

    def step(self, actions, autoreset=False, want_term_obs=False):
        a = np.ascontiguousarray(actions, np.float64).reshape(self.N, self.n_action)
        obs = np.empty((self.N, self.n_obs), np.float32)
        rew = np.empty((self.N, self.n_lead), np.float64)
        rl = np.empty((self.N, self.n_lead), np.float64)
        done = np.empty((self.N,), np.uint8)
        term = np.empty((self.N, self.n_obs), np.float32) if want_term_obs else None
        lib().mho_step(self._h, _ptr(a), _ptr(obs), _ptr(rew), _ptr(rl), _ptr(done), int(autoreset), _ptr(term))
        if want_term_obs:
            return obs, rew, rl, done.astype(bool), term
        return obs, rew, rl, done.astype(bool)

    def _params(self, vals, width):
        a = np.zeros((self.N, width), np.float64)
        for k, v in enumerate(vals):
            a[:, k] = np.asarray(v, np.float64)
        return a

    def reset_pedestrian(self, num_ped, *vals, mask=None):
        """env.reset_pedestrian(num_ped, speed_x, speed_y, pos_x, pos_y, dl, leave, CZ, exist, direction) (SC:948-955; naif:
        (.., dl, direction, cross), NA:897) in the masked envs; every value a scalar or an [N] array."""
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().mho_reset_pedestrian(self._h, int(num_ped), _ptr(self._params(vals, 9)), _ptr(m))

    def reset_cars(self, num_car, *vals, mask=None):
        """env.reset_cars(num_car, speed_x, pos_x, light, line) (SC:957-958)."""
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().mho_reset_cars(self._h, int(num_car), _ptr(self._params(vals, 4)), _ptr(m))

    def observe(self):
        """env.get_state() (SC:960-969)."""
        obs = np.empty((self.N, self.n_obs), np.float32)
        lib().mho_observe(self._h, _ptr(obs))
        return obs

    def get_state(self):
        N, Cn, P = self.N, self.C, self.P
        s = dict(car_f=np.empty((N, Cn, 7)), car_i=np.empty((N, Cn, 2), np.int32), ped_f=np.empty((N, P, 9)),
                 ped_i=np.empty((N, P, 9), np.int32), env_f=np.empty((N, 1)), env_i=np.empty((N, 4), np.int64))
        lib().mho_get_state(self._h, *[_ptr(s[k]) for k in ("car_f", "car_i", "ped_f", "ped_i", "env_f", "env_i")])
        return s

    def set_state(self, s):
        a = [np.ascontiguousarray(s["car_f"], np.float64), np.ascontiguousarray(s["car_i"], np.int32),
             np.ascontiguousarray(s["ped_f"], np.float64), np.ascontiguousarray(s["ped_i"], np.int32),
             np.ascontiguousarray(s["env_f"], np.float64), np.ascontiguousarray(s["env_i"], np.int64)]
        lib().mho_set_state(self._h, *[_ptr(v) for v in a])
