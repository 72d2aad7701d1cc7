"""Learning-curve reproduction: Algo_PPO.train on the scalable 1/1/1 config -- the only configuration the
reference publishes curves for (load_model/parameters/pappo-scalable-coop-111-*-step-001000.npy, written by PY:908-916).

  python tools/train_curves.py --envs 26 --iters 1000 --out profiles/round2_learning/n26      # 26 envs = the reference's 2080 samples / iteration
  python tools/train_curves.py --envs 4096 --iters 300 --out profiles/round2_learning/n4096   # one large-N run

Writes the four traces with the reference's file names under <out>/load_model/parameters/, the checkpoints under
<out>/load_model/weights/, and <out>/summary.json (means of the first / last 10 iterations, wall time)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mhppo_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=26)
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=10)
    ap.add_argument("--out", default="profiles/round2_learning/n26")
    ap.add_argument("--car", type=int, default=1); ap.add_argument("--ped", type=int, default=1); ap.add_argument("--lines", type=int, default=1)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    env = mhppo_b200.VecCrosswalkEnv("coop_scalable", a.envs, nb_car=a.car, nb_ped=a.ped, nb_lines=a.lines, seed=a.seed)
    torch.manual_seed(a.seed)
    D = 2 + 6 * (2 * a.lines - 1) + 10
    algo = mhppo_b200.Algo_PPO(mhppo_b200.Model_PPO, env, num_algo=100 * a.ped + 10 * a.car + a.lines, num_states_c=13, num_states_d=D,
                               num_actions=1, mean=-1.0, std=3.0, nb_cars=a.car, dt=0.3)
    t0 = time.perf_counter()
    algo.train(a.iters, root=a.out)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    algo.saving(root=a.out)
    m = lambda x, s: float(np.mean(x[s])) if len(x) else None
    summ = {"envs": a.envs, "iterations": a.iters, "config": "%d/%d/%d" % (a.ped, a.car, a.lines), "wall_s": wall,
            "samples_per_iteration_mean": float(np.mean(np.sum(algo.ep_scenario_balance, axis=1))),
            "reward_cross_first10": m(algo.ep_reward_cross, slice(0, 10)), "reward_cross_last10": m(algo.ep_reward_cross, slice(-10, None)),
            "reward_wait_first10": m(algo.ep_reward_wait, slice(0, 10)), "reward_wait_last10": m(algo.ep_reward_wait, slice(-10, None)),
            "reward_choice_first10": m(algo.ep_reward_choice, slice(0, 10)), "reward_choice_last10": m(algo.ep_reward_choice, slice(-10, None)),
            "mlp_mode": os.environ.get("MHPPO_MLP", "auto")}
    json.dump(summ, open(os.path.join(a.out, "summary.json"), "w"), indent=1)
    print(json.dumps(summ))


if __name__ == "__main__":
    main()
