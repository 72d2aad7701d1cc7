"""Pins oracle/ppo_oracle.py against the golden fixtures generated from the UNMODIFIED reference PPO classes
(tools/gen_golden_ppo.py), on machines without /root/reference: whole-episode rollout buffers and train steps."""
import os

import numpy as np
import torch

from conftest import GOLDEN_DIR


def _sd(z, prefix):
    return {k[len(prefix) + 1:]: torch.as_tensor(z[k]) for k in z.files if k.startswith(prefix + ".")}


def test_rollout_oracle_reproduces_reference_episodes(oracle_mod):
    from oracle import ppo_oracle as PO
    z = np.load(os.path.join(GOLDEN_DIR, "ppo_rollout_432.npz"))
    sds = [_sd(z, n) for n in ("cross", "wait", "choice")]
    for e, (seed, env_id) in enumerate(z["streams"]):
        seed, env_id = int(seed), int(env_id)
        venv = oracle_mod.OracleVecEnv("coop_scalable", 1, 4, 3, 2, seed=seed, env_id0=env_id, store_f32=False, n_threads=1)
        b = PO.rollout_episode(venv, *sds, seed, [env_id], 3, 2)
        want = {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith("ep%d." % e)}
        np.testing.assert_array_equal(b["exist"][:, 0], want["car_exist"])
        for name, r in (("cross", 0), ("wait", 1)):
            cars = [i for i in range(4) if b["route"][i, 0] == r]
            for key, mine, width in (("obs", "obs_c", 13), ("acts", "act", 0), ("logp", "logp", 0), ("rews", "rew", 0)):
                got = np.concatenate([b[mine][:, i, 0] for i in cars]) if cars else np.zeros((0, width) if width else 0)
                np.testing.assert_allclose(got, want[key + "_" + name], rtol=2e-5, atol=2e-5)
            rtg = np.concatenate([PO.reward_to_go(b["rew"][:, i, 0]) for i in cars]) if cars else np.zeros(0)
            np.testing.assert_allclose(rtg, want["rtg_" + name], rtol=2e-5, atol=2e-5)
        cars = [i for i in range(4) if b["exist"][i, 0]]
        np.testing.assert_allclose(np.stack([b["obs_d"][i, 0] for i in cars]), want["obs_choice"], rtol=1e-5, atol=1e-5)
        np.testing.assert_array_equal(np.array([b["act_d"][i, 0] for i in cars]), want["acts_choice"])
        np.testing.assert_allclose(np.array([b["rew_d"][i, 0] for i in cars]), want["rews_choice"], rtol=1e-6, atol=1e-9)


def test_eval_rollout_oracle_reproduces_reference_episodes(oracle_mod):
    from oracle import ppo_oracle as PO
    z = np.load(os.path.join(GOLDEN_DIR, "ppo_eval_432.npz"))
    sds = [_sd(z, n) for n in ("cross", "wait", "choice")]
    for e, (seed, env_id) in enumerate(z["streams"]):
        seed, env_id = int(seed), int(env_id)
        venv = oracle_mod.OracleVecEnv("coop_scalable", 1, 4, 3, 2, seed=seed, env_id0=env_id, store_f32=False, n_threads=1)
        b = PO.eval_episode(venv, *sds, 3, 2)
        want = {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith("ep%d." % e)}
        np.testing.assert_allclose(b["obs"][:, 0], want["obs"], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(b["acts"][:, 0], want["acts"], rtol=2e-5, atol=2e-5)
        np.testing.assert_array_equal(b["action_d"][:, 0], want["actions"][:, 4:])
        np.testing.assert_allclose(b["rew"][:, 0], want["rew"], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(b["rl"][:, 0], want["rl"], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(b["waiting"][:, 0], want["waiting"], rtol=0, atol=1e-9)


def test_train_step_oracle_reproduces_reference():
    from oracle import ppo_oracle as PO
    for kind, n_in, n_out, mt in (("c", 13, 1, 1), ("d", 30, 2, 2)):
        z = np.load(os.path.join(GOLDEN_DIR, "ppo_train_%s.npz" % kind))
        actor, critic = PO.Net(n_in, n_out, mt), PO.Net(n_in, 1, 0)
        actor.load_state_dict(_sd(z, "actor0")); critic.load_state_dict(_sd(z, "critic0"))
        oa, oc = torch.optim.Adam(actor.parameters(), 3e-4), torch.optim.Adam(critic.parameters(), 1e-3)
        step = PO.train_step_c if kind == "c" else PO.train_step_d
        for ep in range(1, 5):
            step(actor, critic, oa, oc, z["states"], z["actions"].astype(np.float64) if kind == "d" else z["actions"],
                 z["logp_old"].astype(np.float64) if kind == "c" else z["logp_old"], z["rtgs"])
            for net, pre in ((actor, "actor%d" % ep), (critic, "critic%d" % ep)):
                for k, v in net.state_dict().items():
                    np.testing.assert_allclose(v.numpy(), z[pre + "." + k], rtol=1e-5, atol=1e-7, err_msg="%s %s" % (pre, k))


def test_policy_noise_contract_matches_python_shim():
    from oracle import ppo_oracle as PO
    from philox import philox4x32_10, u53
    import math
    for idx, env, seed, it in ((0, 5, 777, 0), (17, 123456789012, 3, 2), (319, 9, 2**40 + 5, 1)):
        w = philox4x32_10((idx, 1 | (it << 8), env & 0xffffffff, env >> 32), (seed & 0xffffffff, seed >> 32))
        z = math.sqrt(-2.0 * math.log(1.0 - u53(w[0], w[1]))) * math.cos(2.0 * math.pi * u53(w[2], w[3]))
        assert abs(float(PO.policy_normal(idx, np.array([env]), seed, it)[0]) - z) < 1e-15
        w = philox4x32_10((idx, 2 | (it << 8), env & 0xffffffff, env >> 32), (seed & 0xffffffff, seed >> 32))
        assert float(PO.policy_uniform(idx, np.array([env]), seed, it)[0]) == u53(w[0], w[1])


def test_state_dict_pack_roundtrip():
    """Model_PPO keeps the reference's state_dict key layout (PY:54-68) around its flat kernel layout."""
    import mhppo_b200
    z = np.load(os.path.join(GOLDEN_DIR, "ppo_rollout_432.npz"))
    for name, n_in, n_out, mt in (("cross", 13, 1, 1), ("choice", 30, 2, 2)):
        sd = _sd(z, name)
        net = mhppo_b200.Model_PPO(n_in, n_out, mt, mean=-1.0, std=3.0, device="cpu")
        net.load_state_dict(sd)
        back = net.state_dict()
        assert sorted(back) == sorted(sd)
        for k in sd:
            assert back[k].shape == sd[k].shape
            torch.testing.assert_close(back[k], sd[k], rtol=0, atol=0)
        x = torch.randn(5, n_in)
        from oracle import ppo_oracle as PO
        want = PO.mlp_forward(sd, x, mt)
        torch.testing.assert_close(net(x).reshape(want.shape), want, rtol=1e-6, atol=1e-6)


def test_eval_oracle_reproduces_shipped_checkpoint_111_episodes(oracle_mod):
    from oracle import ppo_oracle as PO
    """The oracle's deterministic evaluation rollout with the SHIPPED trained nets (1/1/1 scalable env, 18-input choice
    net) against episodes recorded from the unmodified reference (tools/gen_golden_ckpt111.py)."""
    z = np.load(os.path.join(GOLDEN_DIR, "ppo_eval_ckpt111.npz"))
    sd = lambda p: {k[len(p) + 1:]: torch.as_tensor(z[k]) for k in z.files if k.startswith(p + ".")}
    sds = [sd("cross"), sd("wait"), sd("choice")]
    assert sds[2]["layer1.weight"].shape == (32, 18)
    for e, (seed, env_id) in enumerate(z["streams"]):
        venv = oracle_mod.OracleVecEnv("coop_scalable", 1, 1, 1, 1, seed=int(seed), env_id0=int(env_id), store_f32=False)
        b = PO.eval_episode(venv, *sds, 1, 1)
        want = {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith("ep%d." % e)}
        np.testing.assert_array_equal(b["action_d"][:, 0], want["actions"][:, 2:])
        np.testing.assert_allclose(b["acts"][:, 0], want["acts"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(b["rew"][:, 0], want["rew"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(b["rl"][:, 0], want["rl"], rtol=1e-5, atol=1e-5)


import pytest


@pytest.mark.parametrize("name", ["coop_212", "naif_121"])
def test_legacy_rollout_oracle_reproduces_notebook_episodes(oracle_mod, name):
    """The older drivers' feature layout (Coop-MH-PPO.ipynb on coop, MH-PPO.ipynb on naif): oracle vs episodes recorded from
    the notebooks' unmodified first code cell (tools/gen_golden_legacy.py)."""
    from oracle import ppo_oracle as PO
    z = np.load(os.path.join(GOLDEN_DIR, "ppo_rollout_%s.npz" % name))
    c, p, l = [int(x) for x in z["cfg"]]
    variant = str(z["variant"])
    sds = [_sd(z, n) for n in ("cross", "wait", "choice")]
    for e, (seed, env_id) in enumerate(z["streams"]):
        seed, env_id = int(seed), int(env_id)
        venv = oracle_mod.OracleVecEnv(variant, 1, c, p, l, seed=seed, env_id0=env_id, store_f32=False, n_threads=1)
        b = PO.rollout_episode(venv, *sds, seed, [env_id], p, l, legacy_nb_car=c)
        want = {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith("ep%d." % e)}
        for nm, r in (("cross", 0), ("wait", 1)):
            cars = [i for i in range(c) if b["route"][i, 0] == r]
            obs = np.concatenate([b["obs_c"][:, i, 0] for i in cars]) if cars else np.zeros((0, 13), np.float32)
            np.testing.assert_allclose(obs, want["obs_" + nm], rtol=2e-5, atol=2e-5)
            got = np.concatenate([b["act"][:, i, 0] for i in cars]) if cars else np.zeros(0)
            np.testing.assert_allclose(got, want["acts_" + nm], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(np.stack([b["obs_d"][i, 0] for i in range(c)]), want["obs_choice"], rtol=1e-5, atol=1e-5)
        np.testing.assert_array_equal(np.array([b["act_d"][i, 0] for i in range(c)]), want["acts_choice"])
