"""Generate tests/golden/ppo_rollout_{coop_212,naif_121}.npz (build container only): whole rollout episodes of the two older
drivers, loaded UNMODIFIED from their notebooks' first code cell (Coop-MH-PPO.ipynb on `coop` = BASELINE configs[1],
MH-PPO.ipynb on `naif` = configs[0]), with the sampling noise injected through the policy-noise contract
(tools/ref_rollout.py), plus the nets used.      Re-run:  python tools/gen_golden_legacy.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle", "refshim"), os.path.join(ROOT, "tools")]
import ref_rollout  # noqa: E402
import refppo  # noqa: E402


def main():
    for nb, variant, (c, p, l) in ((refppo.NB_COOP, "coop", (2, 1, 2)), (refppo.NB_NAIF, "naif", (1, 2, 1))):
        ns = refppo.load_namespace(nb)
        algo, env = refppo.make_algo(ns, variant, c, p, l, seed=1)
        out = {"cfg": np.array([c, p, l]), "variant": variant}
        for name, net in (("cross", algo.actor_net_cross), ("wait", algo.actor_net_wait), ("choice", algo.actor_net_choice)):
            out.update({name + "." + k: v.detach().numpy().copy() for k, v in net.state_dict().items()})
        streams = [(777, 5), (778, 9), (901, 123), (5, 1), (2024, 77), (31, 4)]
        out["streams"] = np.array(streams, np.int64)
        for e, (seed, env_id) in enumerate(streams):
            for k, v in ref_rollout.reference_episode(ns, algo, env, seed, env_id).items():
                out["ep%d.%s" % (e, k)] = v
        path = os.path.join(ROOT, "tests", "golden", "ppo_rollout_%s_%d%d%d.npz" % (variant, c, p, l))
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
