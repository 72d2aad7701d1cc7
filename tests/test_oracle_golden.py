"""Pins oracle/mhppo_oracle.c (reference semantics, store_f32=0) against the golden fixtures that
tools/gen_golden.py produced from the UNMODIFIED reference envs: every step of every episode, the
observation, rewards, reward_light, done flag and the whole internal state."""
import numpy as np
import pytest

from common import FLT_KEYS, INT_KEYS, assert_close, assert_equal
from conftest import golden_files, load_golden

NO_CROSS_STOP = ("naif", "coop_4cars")  # these classes have no cross_stop attribute (NA:87, C4:209)


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_reproduces_reference_episode(oracle_mod, path):
    g = load_golden(path)
    c, p, l = [int(x) for x in g["cfg"]]
    for ep in range(g["actions"].shape[0]):
        env = oracle_mod.OracleVecEnv(g["variant"], 1, c, p, l, seed=int(g["seed"]), env_id0=int(g["env_id"][ep]),
                                      store_f32=False, n_threads=1)
        obs = env.reset()
        ctx = "(episode %d reset)" % ep
        assert_close("obs", g["obs"][ep, 0], obs[0], 1e-6, 1e-9, ctx)

        def check_state(t):
            s = env.get_state()
            for k in INT_KEYS:
                assert_equal("state." + k, g["st_" + k][ep, t], s[k][0], ctx)
            for k in FLT_KEYS:
                got = s[k][0].copy()
                if k == "ped_f" and g["variant"] in NO_CROSS_STOP:
                    got[:, 6] = g["st_ped_f"][ep, t][:, 6]
                assert_close("state." + k, g["st_" + k][ep, t], got, 1e-9, 1e-9, ctx)

        check_state(0)
        for t in range(80):
            ctx = "(episode %d step %d)" % (ep, t)
            obs, rew, rl, done = env.step(g["actions"][ep, t][None].astype(np.float64))
            assert_close("obs", g["obs"][ep, t + 1], obs[0], 1e-6, 1e-9, ctx)
            assert_close("rewards", g["rewards"][ep, t], rew[0], 1e-9, 1e-9, ctx)
            assert_close("reward_light", g["reward_light"][ep, t], rl[0], 1e-9, 1e-9, ctx)
            assert bool(done[0]) == bool(g["done"][ep, t]), ctx
            check_state(t + 1)
        assert done[0] and int(g["done"][ep].sum()) == 1  # every episode is exactly 80 steps


def test_done_index_is_79(oracle_mod):
    env = oracle_mod.OracleVecEnv("coop", 4, 2, 1, 2, seed=1)
    env.reset()
    a = np.zeros((4, env.n_action))
    dones = [env.step(a)[3].copy() for _ in range(80)]
    assert not np.any(dones[:79]) and np.all(dones[79])


def test_autoreset_continues_stream(oracle_mod):
    """auto-reset == explicit reset of the done envs (same stream cursor)."""
    a_env = oracle_mod.OracleVecEnv("coop_scalable", 8, 4, 3, 2, seed=3)
    b_env = oracle_mod.OracleVecEnv("coop_scalable", 8, 4, 3, 2, seed=3)
    a_env.reset(); b_env.reset()
    rng = np.random.default_rng(1)
    for t in range(85):
        a = rng.uniform(-4, 2, (8, a_env.n_action))
        oa, _, _, da = a_env.step(a, autoreset=True)
        ob, _, _, db = b_env.step(a, autoreset=False)
        if db.any():
            ob = b_env.reset(mask=db)
        np.testing.assert_array_equal(oa, ob)
