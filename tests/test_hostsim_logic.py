"""The kernel's per-env logic (mh-ppo_b200/csrc/env_core.cuh + env_state.cuh) compiled for the host
by tests/hostsim (TEST-ONLY build, never loaded by the product) against the oracle in HBM semantics
and against the golden fixtures.  This is the CPU-side guard for the CUDA code path; the real
parity tests (-m gpu) run the same comparisons through the C ABI on the device."""
import numpy as np
import pytest

import hostsim as H
from common import ALL_CONFIGS, assert_close, compare_vec_envs, inject_and_compare
from conftest import golden_files, load_golden


@pytest.mark.parametrize("cfg", ALL_CONFIGS, ids=lambda c: "%s_%d%d%d" % c)
def test_hostsim_matches_oracle(oracle_mod, cfg):
    v, c, p, l = cfg
    N = 2048  # same scenario as tests/test_env_gpu.py::test_cuda_matches_oracle
    ref = oracle_mod.OracleVecEnv(v, N, c, p, l, seed=42, env_id0=5000, store_f32=True)
    got = H.HostSimEnv(v, N, c, p, l, seed=42, env_id0=5000, soa=(c % 2 == 0))
    compare_vec_envs(ref, got, 165, np.random.default_rng(3), rtol=1e-5, check_state_every=4)


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_hostsim_tracks_reference_golden(path):
    """fp32-stored state free-running for 80 steps next to the fp64 reference: bounded drift."""
    g = load_golden(path)
    c, p, l = [int(x) for x in g["cfg"]]
    for ep in range(g["actions"].shape[0]):
        env = H.HostSimEnv(g["variant"], 1, c, p, l, seed=int(g["seed"]), env_id0=int(g["env_id"][ep]))
        obs = env.reset()
        assert_close("obs", g["obs"][ep, 0], obs[0], 1e-5, 1e-5)
        for t in range(80):
            obs, rew, rl, done = env.step(g["actions"][ep, t][None])
            ctx = "(episode %d step %d)" % (ep, t)
            assert_close("obs", g["obs"][ep, t + 1], obs[0], 1e-4, 1e-4, ctx)
            assert_close("rewards", g["rewards"][ep, t], rew[0], 1e-4, 1e-4, ctx)
            assert_close("reward_light", g["reward_light"][ep, t], rl[0], 1e-4, 1e-4, ctx)
            assert bool(done[0]) == bool(g["done"][ep, t])


@pytest.mark.parametrize("cfg", [("coop_scalable", 4, 3, 2), ("coop", 2, 2, 2), ("stop", 2, 3, 2), ("naif", 2, 2, 2), ("coop_4cars", 2, 2, 2),
                                 ("coop_4cars2", 2, 2, 2)], ids=lambda c: "%s_%d%d%d" % c)
def test_hostsim_state_injection_matches_oracle(oracle_mod, cfg):
    """reset_pedestrian / reset_cars / get_state of the kernel logic (env_core.cuh inject_*, write_obs) vs the oracle."""
    v, c, p, l = cfg
    N = 512
    ref = oracle_mod.OracleVecEnv(v, N, c, p, l, seed=7, env_id0=900, store_f32=True)
    got = H.HostSimEnv(v, N, c, p, l, seed=7, env_id0=900)
    inject_and_compare(ref, got, v, np.random.default_rng(1))
