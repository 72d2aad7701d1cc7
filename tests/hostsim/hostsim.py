"""ctypes wrapper of the TEST-ONLY host build of the kernel logic (tests/hostsim/hostsim.cpp)."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build as _build  # noqa: E402

VARIANT_ID = {"stop": 0, "naif": 1, "coop": 2, "coop_4cars": 3, "coop_4cars2": 4, "coop_scalable": 5}
CAR_B = ((-4.0, 10.0), (2.0, 10.0))
PED_B = ((-0.05, 0.75, 0.0, -3.0), (0.05, 1.75, 4.0, -0.5))
CROSS_B = (2.5, 3.0)


class EnvConst(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nC", "nP", "L", "nb_car", "nlead", "nA", "nobs", "done_idx", "sin_model")] + \
               [("dt", C.c_double), ("acc_lo", C.c_double), ("acc_hi", C.c_double), ("pb", C.c_double * 8),
                ("cross_lo", C.c_double), ("cross_hi", C.c_double),
                ("brake_den", C.c_double), ("idm_den", C.c_double), ("time_braking", C.c_double), ("brake_inv", C.c_double),
                ("idm_rden", C.c_double), ("rdt", C.c_double)]  # filled by hs_create


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("env_stride", C.c_int64), ("comp_stride", C.c_int64)]


def _view(a, soa=False):
    if a is None:
        return View(None, 0, 0)
    if soa:  # array is [width, N]
        return View(a.ctypes.data, 1, a.shape[1])
    return View(a.ctypes.data, a.shape[1], 1)


def done_index(dt, max_episode):
    t, lim, k = 0.0, (max_episode - 1) * dt, 0
    while not (t >= lim):
        t = t + dt
        k += 1
    return k


def pick_template(variant, C_slots, P):
    table = {"coop_scalable": [(2, 1), (4, 3), (8, 4)], "coop": [(2, 1), (8, 4)], "stop": [(8, 4)], "naif": [(8, 4)],
             "coop_4cars": [(2, 2), (4, 2), (6, 4)], "coop_4cars2": [(2, 2), (4, 2), (6, 4)]}[variant]
    exact4 = variant in ("coop_4cars", "coop_4cars2")
    for mc, mp in table:
        if (mc == C_slots if exact4 else mc >= C_slots) and mp >= P:
            return mc, mp
    raise ValueError("no hostsim instantiation for %s C=%d P=%d" % (variant, C_slots, P))


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build_sanitized() if os.environ.get("HOSTSIM_SANITIZE") else _build.build())   # HOSTSIM_SANITIZE: tests/test_hostsim_sanitizers.py
        L.hs_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(EnvConst), C.c_int64, C.c_uint64, C.c_int64,
                                C.POINTER(C.c_void_p)]
        L.hs_destroy.argtypes = [C.c_void_p]
        L.hs_reset.argtypes = [C.c_void_p, C.c_void_p, View]
        L.hs_step.argtypes = [C.c_void_p, View, View, View, View, C.c_void_p, C.c_int, View]
        L.hs_inject.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.hs_observe.argtypes = [C.c_void_p, View]
        L.hs_export.argtypes = [C.c_void_p] * 7
        L.hs_import.argtypes = [C.c_void_p] * 7
        assert L.hs_sizeof_envconst() == C.sizeof(EnvConst)
        _lib = L
    return _lib


class HostSimEnv:
    def __init__(self, variant, n_envs, nb_car, nb_ped, nb_lines, seed=0, env_id0=0, dt=0.3, max_episode=80,
                 simulation="sin", soa=False):
        v = VARIANT_ID[variant]
        Cs = 2 * nb_lines if variant == "coop_scalable" else (2 * nb_car if "4cars" in variant else nb_car)
        nlead = nb_car if "4cars" in variant else Cs
        nA = 4 * nb_lines if variant == "coop_scalable" else (4 * nb_car if variant == "coop_4cars2" else 2 * nb_car)
        nobs = (7 if variant == "coop_scalable" else 6) * Cs + (4 if variant == "coop_scalable" else 3) + 9 * nb_ped
        c = EnvConst(nC=Cs, nP=nb_ped, L=nb_lines, nb_car=nb_car, nlead=nlead, nA=nA, nobs=nobs,
                     done_idx=done_index(dt, max_episode), sin_model=int(simulation == "sin"), dt=dt,
                     acc_lo=CAR_B[0][0], acc_hi=CAR_B[1][0], cross_lo=CROSS_B[0], cross_hi=CROSS_B[1])
        c.pb[:] = [x for r in PED_B for x in r]
        self.c, self.N, self.C, self.P, self.n_lead, self.n_action, self.n_obs = c, n_envs, Cs, nb_ped, nlead, nA, nobs
        self.soa = soa
        mc, mp = pick_template(variant, Cs, nb_ped)
        self._h = C.c_void_p()
        assert lib().hs_create(v, mc, mp, C.byref(c), n_envs, seed, env_id0, C.byref(self._h)) == 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hs_destroy(self._h)
            self._h = None

    def _buf(self, w):
        return np.zeros((w, self.N), np.float32) if self.soa else np.zeros((self.N, w), np.float32)

    def _ret(self, a):
        return a.T if self.soa else a

    def reset(self, mask=None):
        obs = self._buf(self.n_obs)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        assert lib().hs_reset(self._h, None if m is None else m.ctypes.data, _view(obs, self.soa)) == 0
        return self._ret(obs)

    def step(self, actions, autoreset=False, want_term_obs=False):
        a = np.ascontiguousarray(actions, np.float32).reshape(self.N, self.n_action)
        obs, rew, rl = self._buf(self.n_obs), self._buf(self.n_lead), self._buf(self.n_lead)
        term = self._buf(self.n_obs) if want_term_obs else None
        done = np.zeros(self.N, np.uint8)
        assert lib().hs_step(self._h, _view(a), _view(obs, self.soa), _view(rew, self.soa), _view(rl, self.soa),
                             done.ctypes.data, int(autoreset), _view(term, self.soa)) == 0
        out = (self._ret(obs), self._ret(rew), self._ret(rl), done.astype(bool))
        return out + (self._ret(term),) if want_term_obs else out

    def _inject(self, what, slot, vals, width, mask):
        a = np.zeros((self.N, width), np.float32)
        for k, v in enumerate(vals):
            a[:, k] = np.asarray(v, np.float32)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        assert lib().hs_inject(self._h, what, int(slot), a.ctypes.data, None if m is None else m.ctypes.data) == 0

    def reset_pedestrian(self, num_ped, *vals, mask=None):
        self._inject(0, num_ped, vals, 9, mask)

    def reset_cars(self, num_car, *vals, mask=None):
        self._inject(1, num_car, vals, 4, mask)

    def observe(self):
        obs = self._buf(self.n_obs)
        assert lib().hs_observe(self._h, _view(obs, self.soa)) == 0
        return self._ret(obs)

    def get_state(self):
        N, Cn, P = self.N, self.C, self.P
        s = dict(car_f=np.zeros((N, Cn, 7), np.float32), car_i=np.zeros((N, Cn, 2), np.int32),
                 ped_f=np.zeros((N, P, 9), np.float32), ped_i=np.zeros((N, P, 9), np.int32),
                 env_f=np.zeros((N, 1), np.float64), env_i=np.zeros((N, 4), np.int64))
        lib().hs_export(self._h, *[s[k].ctypes.data for k in ("car_f", "car_i", "ped_f", "ped_i", "env_f", "env_i")])
        return s

    def set_state(self, s):
        a = [np.ascontiguousarray(s["car_f"], np.float32), np.ascontiguousarray(s["car_i"], np.int32),
             np.ascontiguousarray(s["ped_f"], np.float32), np.ascontiguousarray(s["ped_i"], np.int32),
             np.ascontiguousarray(s["env_f"], np.float64), np.ascontiguousarray(s["env_i"], np.int64)]
        lib().hs_import(self._h, *[v.ctypes.data for v in a])
