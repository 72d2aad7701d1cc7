"""Multi-GPU parity script (run under torchrun on a box with >= 2 GPUs, e.g.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_ppo_check.py
): R ranks, each holding N/R envs with disjoint global env ids, must perform the same PPO iteration as ONE rank holding
all N envs: identical env outputs (bit-exact rollout buffers on the shard) and the same updated
weights (fp32 summation order differs between 1 and R partial sums: 1e-5)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mhppo_b200  # noqa: E402
from mhppo_b200 import ppo  # noqa: E402


def run(N, env_id0, dev, use_dist, epochs=3):
    ppo.set_distributed(use_dist)
    env = mhppo_b200.VecCrosswalkEnv("coop_scalable", N, nb_car=4, nb_ped=3, nb_lines=2, seed=77, env_id0=env_id0, device=dev)
    # the sharded run seeds every rank differently: Algo_PPO must broadcast rank 0's initial parameters (sync_parameters)
    torch.manual_seed(0 if not use_dist else 1000 * int(os.environ.get("RANK", "0")))
    algo = mhppo_b200.Algo_PPO(mhppo_b200.Model_PPO, env, num_states_c=13, num_states_d=30, num_actions=1, mean=-1.0, std=3.0, nb_cars=4, dt=0.3)
    algo.rollout.iterations_rand(algo.actor_net_cross, algo.actor_net_wait, algo.actor_net_choice)
    algo.update(epochs=epochs)
    torch.cuda.synchronize()
    return algo


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N = 4096
    n = N // world
    sharded = run(n, rank * n, dev, True)
    whole = run(N, 0, dev, False)                     # every rank repeats the single-rank run locally
    r_s, r_w = sharded.rollout, whole.rollout
    T, C = r_s.T, r_s.C
    for name in ("act", "logp", "rew", "rtg"):
        a = getattr(r_s, name).view(T, C, n)
        b = getattr(r_w, name).view(T, C, N)[:, :, rank * n:(rank + 1) * n]
        assert torch.equal(a, b), "rollout buffer %s differs on rank %d" % (name, rank)
    worst = 0.0
    for (_, _, ns), (_, _, nw) in zip(sharded._nets(), whole._nets()):
        d = (ns.flat - nw.flat).abs().max().item() / max(nw.flat.abs().max().item(), 1e-12)
        worst = max(worst, d)
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("dist_ppo_check world=%d: rollout shards bit-identical, max relative weight difference %.3e" % (world, t.item()))
    assert t.item() < 1e-4, t.item()
    if rank == 0:
        print("OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
