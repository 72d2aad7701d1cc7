"""Pin oracle/mhppo_oracle.c against the UNMODIFIED reference envs (build container only).

Runs free-running 80-step episodes of the reference (Philox RNG injected) and of the oracle in
reference semantics (store_f32=0) on the same (seed, env_id, actions) and compares, per step, the
observation, rewards, reward_light, done and the full internal state dump.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
import refdriver as rd  # noqa: E402
from oracle import oracle as O  # noqa: E402

CONFIGS = [
    ("coop_scalable", 4, 3, 2), ("coop_scalable", 1, 1, 1), ("coop_scalable", 8, 4, 4), ("coop_scalable", 3, 2, 3),
    ("coop", 2, 1, 2), ("coop", 2, 2, 2), ("coop", 1, 1, 1), ("coop", 3, 3, 3),
    ("stop", 1, 2, 1), ("stop", 2, 3, 2), ("naif", 1, 2, 1), ("naif", 2, 2, 2), ("naif", 3, 3, 2),
    ("coop_4cars", 2, 2, 2), ("coop_4cars", 1, 1, 1), ("coop_4cars2", 2, 2, 2), ("coop_4cars2", 3, 2, 3),
]


def compare_episode(variant, nb_car, nb_ped, nb_lines, seed, env_id, rng, light_mode, rtol=1e-9, atol=1e-9):
    env = rd.make_env(variant, nb_car, nb_ped, nb_lines)
    acts = rd.random_actions(rng, variant, nb_car, nb_lines, 80, light_mode)
    ref = rd.run_episode(env, seed, env_id, acts)
    ora = O.OracleVecEnv(variant, 1, nb_car, nb_ped, nb_lines, seed=seed, env_id0=env_id, store_f32=False, n_threads=1)
    obs0 = ora.reset()
    errs = []

    def chk(name, a, b, t):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        if a.shape != b.shape:
            errs.append((t, name, "shape", a.shape, b.shape)); return
        bad = ~np.isclose(a, b, rtol=rtol, atol=atol, equal_nan=True)
        if bad.any():
            idx = np.argwhere(bad)[0]
            errs.append((t, name, tuple(idx), a[tuple(idx)], b[tuple(idx)]))

    def chk_state(t):
        s = ora.get_state()
        for k in ("car_f", "car_i", "ped_f", "ped_i", "env_f", "env_i"):
            r = ref["st_" + k][t]
            o = s[k][0]
            if k == "ped_f" and variant in ("naif", "coop_4cars"):
                o = o.copy(); o[:, 6] = r[:, 6]  # cross_stop does not exist in these variants
            chk("st_" + k, r, o, t)

    chk("obs", ref["obs"][0], obs0[0], -1); chk_state(0)
    for t in range(80):
        obs, rew, rl, done = ora.step(acts[t][None])
        chk("obs", ref["obs"][t + 1], obs[0], t)
        chk("rewards", ref["rewards"][t], rew[0], t)
        chk("reward_light", ref["reward_light"][t], rl[0], t)
        chk("done", ref["done"][t], done[0], t)
        chk_state(t + 1)
        if errs:
            break
    return errs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", type=int, default=20)
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    rng = np.random.default_rng(2024)
    total_bad = 0
    for (v, c, p, l) in CONFIGS:
        if a.only and v != a.only:
            continue
        bad = 0
        first = None
        for ep in range(a.episodes):
            lm = ["episode", "step"][ep % 2]
            e = compare_episode(v, c, p, l, seed=1234 + ep, env_id=ep * 7 + 3, rng=rng, light_mode=lm)
            if e:
                bad += 1
                first = first or (ep, e[:3])
        print("%-14s car=%d ped=%d lines=%d : %d/%d episodes mismatch %s" % (v, c, p, l, bad, a.episodes, first or ""))
        total_bad += bad
    sys.exit(1 if total_bad else 0)


if __name__ == "__main__":
    main()
