"""Host-side multi-rank logic on CPU (gloo, world_size 2): the only cross-rank steps of the hot path are
(1) summing the advantage statistics (sum A, sum A^2, n) and (2) summing flat gradient buffers.
An R-rank run must normalise advantages and average gradients exactly like one rank holding every sample."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mhppo_b200.ppo import allreduce_sum_, combine_stats
    g = torch.Generator().manual_seed(7)
    A = torch.randn(1001, generator=g, dtype=torch.float64) * 3 + 1          # the global advantage vector
    grads = torch.randn(world, 4868, generator=g)                            # per-rank partial gradient sums
    mine = A[rank::world]
    stats = torch.tensor([mine.sum(), (mine * mine).sum(), float(mine.numel())], dtype=torch.float64)
    mean, inv_std, n = combine_stats(allreduce_sum_(stats))
    gsum = allreduce_sum_(grads[rank].clone())
    q.put((rank, mean, inv_std, n, gsum.numpy()))
    dist.destroy_process_group()


def test_two_rank_statistics_and_gradient_sum_equal_single_rank():
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(60) for p in ps]
    g = torch.Generator().manual_seed(7)
    A = torch.randn(1001, generator=g, dtype=torch.float64) * 3 + 1
    grads = torch.randn(world, 4868, generator=g)
    for rank, mean, inv_std, n, gsum in res:
        assert n == 1001
        assert abs(mean - float(A.mean())) < 1e-12
        assert abs(inv_std - 1.0 / (float(A.std()) + 1e-10)) < 1e-10     # torch .std() is the unbiased estimator (PY:787)
        np.testing.assert_allclose(gsum, grads.sum(0).numpy(), rtol=1e-6, atol=1e-6)


def test_env_shards_use_disjoint_global_ids():
    """Rank r of R owns envs [r*n, (r+1)*n): the Philox stream is keyed by the global id (oracle = same contract)."""
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    whole = O.OracleVecEnv("coop_scalable", 64, 4, 3, 2, seed=9, env_id0=0)
    parts = [O.OracleVecEnv("coop_scalable", 32, 4, 3, 2, seed=9, env_id0=32 * r) for r in range(2)]
    np.testing.assert_array_equal(whole.reset(), np.concatenate([p.reset() for p in parts]))
