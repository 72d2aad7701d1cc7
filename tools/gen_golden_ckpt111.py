"""Generate tests/golden/ppo_eval_ckpt111.npz (build container only): whole episodes of the UNMODIFIED reference's
deterministic evaluation rollout (Env_rollout.iterations, PY:152-252) on the scalable env at nb_car = nb_ped = nb_lines = 1
with the SHIPPED trained checkpoints load_model/weights/pappo-scalable-coop-*-111-*-step-1000.pth (Algo_PPO.loading(111,
1000), PY:985-1001), plus those weights (the only published trained nets of the scalable script; 18-input choice net).
Re-run:  python tools/gen_golden_ckpt111.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle", "refshim"), os.path.join(ROOT, "tools")]
import ref_rollout  # noqa: E402
import refppo  # noqa: E402
import refdriver as rd  # noqa: E402


def main():
    ns = refppo.load_namespace()
    algo, env = refppo.make_algo(ns, "coop_scalable", 1, 1, 1, seed=1)
    cwd = os.getcwd()
    os.chdir(rd.REF_ROOT)
    try:
        algo.loading(111, 1000)
    finally:
        os.chdir(cwd)
    out = {}
    for name, net in (("cross", algo.actor_net_cross), ("wait", algo.actor_net_wait), ("choice", algo.actor_net_choice),
                      ("critic_cross", algo.critic_net_cross), ("critic_wait", algo.critic_net_wait), ("critic_choice", algo.critic_net_choice)):
        out.update({name + "." + k: v.detach().numpy().copy() for k, v in net.state_dict().items()})
    streams = [(777, 5), (901, 123), (7, 3), (8, 4), (2024, 11), (31, 1000)]
    out["streams"] = np.array(streams, np.int64)
    for e, (seed, env_id) in enumerate(streams):
        w = ref_rollout.reference_eval_episode(ns, algo, env, seed, env_id)
        for k, v in w.items():
            out["ep%d.%s" % (e, k)] = v
    p = os.path.join(ROOT, "tests", "golden", "ppo_eval_ckpt111.npz")
    np.savez_compressed(p, **out)
    print(p, os.path.getsize(p) // 1024, "KiB")


if __name__ == "__main__":
    main()
