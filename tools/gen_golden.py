"""Generate tests/golden/*.npz from the UNMODIFIED reference envs (build container only).

For each (variant, nb_car, nb_ped, nb_lines) a few free-running 80-step episodes of the reference
with the project's Philox stream injected (oracle/refshim).  Each fixture stores the stream
coordinates (seed, env_id), the fp32-representable action sequence, and per step the flat
observation (fp32), rewards / reward_light (fp64), done, and the canonical state dump.
Re-run:  python tools/gen_golden.py        (needs /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
import refdriver as rd  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CONFIGS = [
    ("coop_scalable", 4, 3, 2), ("coop_scalable", 1, 1, 1), ("coop_scalable", 8, 4, 4),
    ("coop", 2, 1, 2), ("coop", 3, 3, 3), ("stop", 1, 2, 1), ("stop", 2, 3, 2),
    ("naif", 1, 2, 1), ("naif", 3, 3, 2), ("coop_4cars", 2, 2, 2), ("coop_4cars2", 2, 2, 2), ("coop_4cars2", 3, 2, 3),
]
EPISODES = 3
SEED = 20261018


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(7)
    for (v, c, p, l) in CONFIGS:
        env = rd.make_env(v, c, p, l)
        eps = []
        for ep in range(EPISODES):
            acts = rd.random_actions(rng, v, c, l, 80, ["episode", "step", "episode"][ep]).astype(np.float32).astype(np.float64)
            env_id = 1000 * ep + 17
            r = rd.run_episode(env, SEED, env_id, acts)
            r["actions"] = acts.astype(np.float32)
            r["env_id"] = np.int64(env_id)
            eps.append(r)
        out = {"seed": np.int64(SEED), "variant": np.array(v), "cfg": np.array([c, p, l], np.int32)}
        for k in eps[0]:
            out[k] = np.stack([e[k] for e in eps])
        path = os.path.join(OUT, "%s_%d%d%d.npz" % (v, c, p, l))
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
