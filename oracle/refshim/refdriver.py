"""Drive the UNMODIFIED reference envs as the oracle of record (TEST INFRASTRUCTURE).

Only usable where `/root/reference` (or $MHPPO_REFERENCE) exists, i.e. in the
build container.  Used to (1) pin the C restatement `oracle/mhppo_oracle.c`
against the real reference and (2) generate the committed golden fixtures under
`tests/golden/` (`tools/gen_golden.py`).  Nothing here is imported by the product.

The reference modules are imported as-is; the only interventions are
  * `gym` resolves to the shim next to this file (gym is not installed),
  * each env module's `random` attribute is replaced by a `PhiloxRandom`
    (SURVEY.md 8c "RNG control"), and
  * `print` inside the env modules is silenced (detection() prints event strings,
    Env_hybrid_multi_coop_scalable.py:186,200,222,227,237).
"""
import importlib
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
from philox import PhiloxRandom  # noqa: E402

REF_ROOT = os.environ.get("MHPPO_REFERENCE", "/root/reference")

# variant id -> (module, class, gym id)
VARIANTS = {
    "stop": ("Env_hybrid_multi_stop", "Crosswalk_hybrid_multi_stop"),
    "naif": ("Env_hybrid_multi_naif", "Crosswalk_hybrid_multi_naif"),
    "coop": ("Env_hybrid_multi_coop", "Crosswalk_hybrid_multi_coop"),
    "coop_4cars": ("Env_hybrid_multi_coop_4cars", "Crosswalk_hybrid_multi_coop_4cars"),
    "coop_4cars2": ("Env_hybrid_multi_coop_4cars2", "Crosswalk_hybrid_multi_coop_4cars2"),
    "coop_scalable": ("Env_hybrid_multi_coop_scalable", "Crosswalk_hybrid_multi_coop_scalable"),
}
VARIANT_IDS = {n: i for i, n in enumerate(VARIANTS)}

CAR_B = np.array([[-4.0, 10.0], [2.0, 10.0]])
PED_B = np.array([[-0.05, 0.75, 0.0, -3.0], [0.05, 1.75, 4.0, -0.5]])
CROSS_B = np.array([2.5, 3.0])
DT = 0.3
MAX_EPISODE = 80

# flag bits of ped_i[..., 8] in the canonical state dump
F_EXIST, F_IS_CROSSING, F_DECISION, F_AT_CROSSING, F_PED_LEFT, F_PED_IN_CROSS = 0, 1, 2, 3, 4, 5
F_NOT_WAITING, F_ACCIDENT, F_WORST_ACC, F_FOLLOW_RULE, F_STOP, F_NEED_TO_STOP = 6, 7, 8, 9, 10, 11


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "Environments"))


_loaded = {}


def load_module(variant):
    """Import the reference module for `variant` and give it a PhiloxRandom."""
    if variant in _loaded:
        return _loaded[variant]
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    mod = importlib.import_module("Environments." + VARIANTS[variant][0])
    mod.random = PhiloxRandom()
    mod.print = lambda *a, **k: None  # module-level name shadows the builtin
    _loaded[variant] = mod
    return mod


def make_env(variant, nb_car, nb_ped, nb_lines, dt=DT, max_episode=MAX_EPISODE, simulation="sin",
             car_b=CAR_B, ped_b=PED_B, cross_b=CROSS_B):
    mod = load_module(variant)
    cls = getattr(mod, VARIANTS[variant][1])
    env = cls(car_b=car_b, ped_b=ped_b, cross_b=cross_b, nb_car=nb_car, nb_ped=nb_ped,
              nb_lines=nb_lines, dt=dt, max_episode=max_episode, simulation=simulation)
    env._mh_variant = variant
    env._mh_rng = mod.random
    return env


def all_cars(env):
    """Car slots in the project's canonical order: leaders, then followers."""
    cars = list(env.cars)
    if hasattr(env, "cars_follow"):
        cars += list(env.cars_follow)
    return cars


def extract_state(env):
    """Canonical state dump (see include/mhppo.h `mhppo_state_dump`) of one reference env."""
    cars = all_cars(env)
    peds = env.pedestrian
    C, P = len(cars), len(peds)
    car_f = np.zeros((C, 7), np.float64)
    car_i = np.zeros((C, 2), np.int32)
    for i, c in enumerate(cars):
        car_f[i] = [c.Ac, c.Vc, c.Sc, c.light, c.possible_accident, c.error_scenario, c.Ts]
        car_i[i] = [c.line, int(getattr(c, "exist", True))]
    ped_f = np.zeros((P, 9), np.float64)
    ped_i = np.zeros((P, 9), np.int32)
    dt = env.dt
    for j, p in enumerate(peds):
        ped_f[j] = [p.Vp_x, p.Vp_y, p.Sp_x, p.Sp_y, p.initial_speed[0], p.initial_speed[1],
                    getattr(p, "cross_stop", 0.0), p.delta, p.worst_dl]
        flags = 0
        for bit, name in ((F_EXIST, "exist"), (F_IS_CROSSING, "is_crossing"), (F_DECISION, "decision"),
                          (F_AT_CROSSING, "at_crossing"), (F_PED_LEFT, "ped_left"),
                          (F_PED_IN_CROSS, "ped_in_cross"), (F_NOT_WAITING, "ped_not_waiting"),
                          (F_ACCIDENT, "accident"), (F_WORST_ACC, "worst_scenario_accident"),
                          (F_FOLLOW_RULE, "follow_rule"), (F_STOP, "stop"), (F_NEED_TO_STOP, "need_to_stop")):
            if bool(getattr(p, name, False)):
                flags |= 1 << bit
        # reset_ped (SC:111) replaces the constructor's ratio v0x / v0y (SC:73) by v0x / (v0y + 1e-3): canonical flag bit 12
        if p.initial_speed[1] != 0 and p.ratio != p.initial_speed[0] / p.initial_speed[1]:
            flags |= 1 << 12
        ped_i[j] = [int(round(p.t0 / dt)), int(round(p.waiting_time / dt)), int(round(p.crossing_time / dt)),
                    int(p.time_stop), int(p.line_pos), int(p.direction), int(p.gender), int(p.age), flags]
    env_f = np.array([env.cross], np.float64)
    env_i = np.array([int(round(env.time / dt)), env.ped_traffic, getattr(env, "car_traffic", len(env.cars)),
                      env._mh_rng.ctr], np.int64)
    return dict(car_f=car_f, car_i=car_i, ped_f=ped_f, ped_i=ped_i, env_f=env_f, env_i=env_i)


def flat_obs(state_dict):
    """Flatten the obs dict in gym-0.26 `Dict` key order (sorted): car,(car_follow),env,ped."""
    return np.concatenate([np.asarray(state_dict[k], np.float32).ravel() for k in sorted(state_dict.keys())])


def run_episode(env, seed, env_id, actions, n_steps=None, ctr=0, dump_state=True):
    """reset + step the reference with stream (seed, env_id); actions: [T, A] float64.

    Returns a dict of stacked per-step arrays.  Index 0 of obs/state is the post-reset
    value; index t+1 is after step t.
    """
    env._mh_rng.set_stream(seed, env_id, ctr)
    obs, _ = env.reset()
    T = len(actions) if n_steps is None else n_steps
    out = dict(obs=[flat_obs(obs)], rewards=[], reward_light=[], done=[], state=[])
    if dump_state:
        out["state"].append(extract_state(env))
    for t in range(T):
        obs, rew, done, trunc, _ = env.step(np.asarray(actions[t], np.float64))
        out["obs"].append(flat_obs(obs))
        out["rewards"].append(np.asarray(rew, np.float64))
        out["reward_light"].append(np.asarray(env.reward_light, np.float64).copy())
        out["done"].append(bool(done))
        if dump_state:
            out["state"].append(extract_state(env))
    res = dict(obs=np.stack(out["obs"]), rewards=np.stack(out["rewards"]),
               reward_light=np.stack(out["reward_light"]), done=np.array(out["done"]))
    if dump_state:
        for k in out["state"][0]:
            res["st_" + k] = np.stack([s[k] for s in out["state"]])
    return res


def n_action(variant, nb_car, nb_lines):
    """Length of the action vector each variant's step() indexes."""
    if variant == "coop_scalable":
        return 4 * nb_lines          # [acc x 2L, light x 2L]      (SC:798-802)
    if variant == "coop_4cars2":
        return 4 * nb_car            # [acc_l, acc_f, light_l, light_f] (C42:809-811)
    return 2 * nb_car                # [acc x C, light x C]        (CO:754-755)


def random_actions(rng, variant, nb_car, nb_lines, T, light_mode="episode"):
    """Synthetic action stream used by goldens and the bench: acc~U(-4,2), light in {-1,+1}
    held for the whole episode (the rollout holds the light, PY:469-476)."""
    A = n_action(variant, nb_car, nb_lines)
    half = A // 2
    acts = np.zeros((T, A))
    acts[:, :half] = rng.uniform(-4.0, 2.0, size=(T, half))
    if light_mode == "episode":
        acts[:, half:] = rng.choice([-1.0, 1.0], size=(1, half))
    elif light_mode == "step":
        acts[:, half:] = rng.choice([-1.0, 0.0, 1.0], size=(T, half))
    return acts
