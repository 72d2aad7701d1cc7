"""Pins oracle/ppo_oracle.py against the UNMODIFIED reference PPO classes (build container only)."""
import os
import sys

import numpy as np
import pytest
import torch

import refdriver as rd

pytestmark = pytest.mark.skipif(not rd.available(), reason="reference sources not present")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def ref():
    import refppo
    ns = refppo.load_namespace()
    algo, env = refppo.make_algo(ns, "coop_scalable", 4, 3, 2, seed=1)
    return ns, algo, env


def _sd(net):
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def test_feature_builders_match_reference(ref):
    from oracle import ppo_oracle as PO
    ns, algo, env = ref
    r = algo.rollout
    rng = np.random.default_rng(0)
    for ep in range(6):
        env._mh_rng.set_stream(50 + ep, ep, 0)
        state, _ = env.reset()
        for t in range(80):
            obs = np.concatenate([state[k].ravel() for k in sorted(state)])
            if t % 7 == 0:
                for i in range(4):
                    assert r.closest_ped_d(obs, i) == PO.closest_ped_d(obs[None], i, 3, 2)[0]
                    for p in range(3):
                        f_ref, ex_ref = r.obs_car_ped(obs, i, p)
                        f, ex = PO.obs_car_ped(obs[None], i, p, 3, 2)
                        np.testing.assert_allclose(f[0], np.asarray(f_ref, np.float32), rtol=1e-6, atol=1e-6)
                        assert ex[0] == ex_ref
                        d_ref, _ = r.obs_car_ped_d(obs, i, p)
                        d, _ = PO.obs_car_ped_d(obs[None], i, p, 3, 2)
                        np.testing.assert_allclose(d[0], np.asarray(d_ref, np.float32), rtol=1e-6, atol=1e-6)
            a = rd.random_actions(rng, "coop_scalable", 4, 2, 1)[0]
            state, *_ = env.step(a)


def test_mlp_forward_matches_reference_on_shipped_checkpoints(ref):
    from oracle import ppo_oracle as PO
    ns, algo, env = ref
    wdir = os.path.join(rd.REF_ROOT, "load_model", "weights")
    x13 = torch.randn(64, 13, generator=torch.Generator().manual_seed(0))
    for name, mt in (("cross", 1), ("wait", 1)):
        sd = torch.load(os.path.join(wdir, "pappo-scalable-coop-%s-111-actor-step-1000.pth" % name))
        net = ns["Model_PPO"](13, 1, mt, mean=-1.0, std=3.0)
        net.load_state_dict(sd)
        torch.testing.assert_close(net(x13), PO.mlp_forward(sd, x13, mt), rtol=1e-6, atol=1e-6)
    sd = torch.load(os.path.join(wdir, "pappo-scalable-coop-choice-111-actor-step-1000.pth"))
    net = ns["Model_PPO"](18, 2, 2)
    net.load_state_dict(sd)
    x18 = torch.randn(64, 18, generator=torch.Generator().manual_seed(1))
    got = PO.mlp_forward(sd, x18, 2)
    torch.testing.assert_close(torch.stack([net(x18[i:i + 1]) for i in range(64)]), got, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("kind", ["c", "d"])
def test_train_step_matches_reference(ref, kind):
    from oracle import ppo_oracle as PO
    ns, algo, env = ref
    g = torch.Generator().manual_seed(3)
    M, n_in = (257, 13) if kind == "c" else (61, 30)
    states = torch.randn(M, n_in, generator=g).numpy()
    rtgs = torch.randn(M, generator=g) * 5 - 3
    if kind == "c":
        actor_r, critic_r = algo.actor_net_cross, algo.critic_net_cross
        opt_a, opt_c = algo.optimizer_actor_cross, algo.optimizer_critic_cross
        actions = (torch.randn(M, 1, generator=g) * 2 - 1).numpy()
        logp_old = (-torch.rand(M, generator=g) * 3).double().numpy()
        actor_o, critic_o = PO.Net(13, 1, 1), PO.Net(13, 1, 0)
        step_ref, step_o = algo.train_model_c, PO.train_step_c
        cov = algo.cov_mat
    else:
        actor_r, critic_r = algo.actor_net_choice, algo.critic_net_choice
        opt_a, opt_c = algo.optimizer_actor_choice, algo.optimizer_critic_choice
        actions = torch.randint(0, 2, (M, 1), generator=g).double().numpy()
        logp_old = (-torch.rand(M, generator=g) * 1.5).float().numpy()
        actor_o, critic_o = PO.Net(30, 2, 2), PO.Net(30, 1, 0)
        step_ref, step_o = algo.train_model_d, PO.train_step_d
        cov = algo.cov_mat_d
    actor_o.load_state_dict(_sd(actor_r)); critic_o.load_state_dict(_sd(critic_r))
    oa = torch.optim.Adam(actor_o.parameters(), 3e-4); oc = torch.optim.Adam(critic_o.parameters(), 1e-3)
    for epoch in range(4):
        step_ref(actor_r, critic_r, opt_a, opt_c, states, actions, logp_old, rtgs, cov)
        step_o(actor_o, critic_o, oa, oc, states, actions, logp_old, rtgs)
    for (k, a), (_, b) in zip(list(actor_r.state_dict().items()) + list(critic_r.state_dict().items()),
                              list(actor_o.state_dict().items()) + list(critic_o.state_dict().items())):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-7, msg=k)


def test_rollout_episode_matches_reference(ref):
    """Whole-episode integration: reference Env_rollout.iterations_rand (noise injected) vs the vectorised
    oracle pipeline on the same streams: buffers, routing quirk (PY:491), episodic choice reward, RTG."""
    import ref_rollout
    from oracle import oracle as O
    from oracle import ppo_oracle as PO
    ns, algo, env = ref
    sds = [_sd(n) for n in (algo.actor_net_cross, algo.actor_net_wait, algo.actor_net_choice)]
    for ep, (seed, env_id) in enumerate([(777, 5), (778, 9), (901, 123), (5, 1), (6, 2), (7, 3)]):
        want = ref_rollout.reference_episode(ns, algo, env, seed, env_id)
        venv = O.OracleVecEnv("coop_scalable", 1, 4, 3, 2, seed=seed, env_id0=env_id, store_f32=False, n_threads=1)
        b = PO.rollout_episode(venv, *sds, seed, [env_id], 3, 2)
        np.testing.assert_array_equal(b["exist"][:, 0], want["car_exist"])
        for name, r in (("cross", 0), ("wait", 1)):
            cars = [i for i in range(4) if b["route"][i, 0] == r]
            obs = np.concatenate([b["obs_c"][:, i, 0] for i in cars]) if cars else np.zeros((0, 13), np.float32)
            np.testing.assert_allclose(obs, want["obs_" + name], rtol=2e-5, atol=2e-5)
            for key, mine in (("acts", "act"), ("logp", "logp"), ("rews", "rew")):
                got = np.concatenate([b[mine][:, i, 0] for i in cars]) if cars else np.zeros(0)
                np.testing.assert_allclose(got, want[key + "_" + name], rtol=2e-5, atol=2e-5, err_msg="%s %s ep %d" % (key, name, ep))
            rtg = np.concatenate([PO.reward_to_go(b["rew"][:, i, 0]) for i in cars]) if cars else np.zeros(0)
            np.testing.assert_allclose(rtg, want["rtg_" + name], rtol=2e-5, atol=2e-5)
        cars = [i for i in range(4) if b["exist"][i, 0]]
        np.testing.assert_allclose(np.stack([b["obs_d"][i, 0] for i in cars]), want["obs_choice"], rtol=1e-5, atol=1e-5)
        np.testing.assert_array_equal(np.array([b["act_d"][i, 0] for i in cars]), want["acts_choice"])
        np.testing.assert_allclose(np.array([b["logp_d"][i, 0] for i in cars]), want["logp_choice"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(np.array([b["rew_d"][i, 0] for i in cars]), want["rews_choice"], rtol=1e-6, atol=1e-9)


def test_eval_rollout_matches_reference(ref):
    """Deterministic evaluation rollout: reference Env_rollout.iterations (env.step wrapped to record what it receives and
    returns) vs PO.eval_episode, both decision regimes (ped_traffic == nb_ped: decided once; < nb_ped: every step)."""
    import ref_rollout
    from oracle import oracle as O
    from oracle import ppo_oracle as PO
    ns, algo, env = ref
    sds = [_sd(n) for n in (algo.actor_net_cross, algo.actor_net_wait, algo.actor_net_choice)]
    regimes = set()
    for seed, env_id in [(777, 5), (778, 9), (901, 123), (7, 3), (8, 4)]:
        want = ref_rollout.reference_eval_episode(ns, algo, env, seed, env_id)
        venv = O.OracleVecEnv("coop_scalable", 1, 4, 3, 2, seed=seed, env_id0=env_id, store_f32=False, n_threads=1)
        b = PO.eval_episode(venv, *sds, 3, 2)
        regimes.add(want["rews_d"].shape[0])
        np.testing.assert_allclose(b["obs"][:, 0], want["obs"], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(b["acts"][:, 0], want["acts"], rtol=2e-5, atol=2e-5)
        np.testing.assert_array_equal(b["action_d"][:, 0], want["actions"][:, 4:])       # the whole decision vector is appended (PY:216)
        np.testing.assert_allclose(b["rew"][:, 0], want["rew"], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(b["rl"][:, 0], want["rl"], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(b["waiting"][:, 0], want["waiting"], rtol=0, atol=1e-9)
    assert regimes == {1, 80}


@pytest.mark.parametrize("nb,variant,cfg", [("coop", "coop", (2, 1, 2)), ("coop", "coop", (3, 3, 3)), ("naif", "naif", (1, 2, 1)), ("naif", "naif", (2, 2, 2))],
                         ids=["coop_212", "coop_333", "naif_121", "naif_222"])
def test_legacy_layout_rollout_matches_the_notebooks(nb, variant, cfg):
    """The two older drivers (Coop-MH-PPO.ipynb on `coop` = BASELINE configs[1], MH-PPO.ipynb on `naif` = configs[0]; same
    PPO cell): their first code cell is loaded UNMODIFIED and whole rollout episodes are compared with the oracle's legacy
    layout (6-float car rows, 3-float env row, 5 columns per other car, every pedestrian slot visited, closest pedestrian
    seeded with slot 0)."""
    import ref_rollout
    import refppo
    from oracle import oracle as O
    from oracle import ppo_oracle as PO
    ns = refppo.load_namespace(refppo.NB_COOP if nb == "coop" else refppo.NB_NAIF)
    c, p, l = cfg
    algo, env = refppo.make_algo(ns, variant, c, p, l, seed=1)
    assert algo.actor_net_choice.layer1.weight.shape[1] == 2 + 5 * (c - 1) + 10
    sds = [_sd(n) for n in (algo.actor_net_cross, algo.actor_net_wait, algo.actor_net_choice)]
    for seed, env_id in [(777, 5), (901, 123), (5, 1), (6, 2)]:
        want = ref_rollout.reference_episode(ns, algo, env, seed, env_id)
        venv = O.OracleVecEnv(variant, 1, c, p, l, seed=seed, env_id0=env_id, store_f32=False, n_threads=1)
        b = PO.rollout_episode(venv, *sds, seed, [env_id], p, l, legacy_nb_car=c)
        for name, r in (("cross", 0), ("wait", 1)):
            cars = [i for i in range(c) if b["route"][i, 0] == r]
            obs = np.concatenate([b["obs_c"][:, i, 0] for i in cars]) if cars else np.zeros((0, 13), np.float32)
            np.testing.assert_allclose(obs, want["obs_" + name], rtol=2e-5, atol=2e-5)
            for key, mine in (("acts", "act"), ("logp", "logp"), ("rews", "rew")):
                got = np.concatenate([b[mine][:, i, 0] for i in cars]) if cars else np.zeros(0)
                np.testing.assert_allclose(got, want[key + "_" + name], rtol=2e-5, atol=2e-5)
            rtg = np.concatenate([PO.reward_to_go(b["rew"][:, i, 0]) for i in cars]) if cars else np.zeros(0)
            np.testing.assert_allclose(rtg, want["rtg_" + name], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(np.stack([b["obs_d"][i, 0] for i in range(c)]), want["obs_choice"], rtol=1e-5, atol=1e-5)
        np.testing.assert_array_equal(np.array([b["act_d"][i, 0] for i in range(c)]), want["acts_choice"])
        np.testing.assert_allclose(np.array([b["logp_d"][i, 0] for i in range(c)]), want["logp_choice"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(np.array([b["rew_d"][i, 0] for i in range(c)]), want["rews_choice"], rtol=1e-6, atol=1e-9)


def _reference_get_average_prints(ep_car, ep_ped, ep_cross, nb_car, nb_ped):
    """Execute the reference's own `get_average` text (PY:1550-1675) and collect what it prints.  The CO2 helper needs
    LDV.csv, which the reference does not ship: it is stubbed out; `states` / `env` are the globals the cell reads."""
    import re
    import types
    src = open(os.path.join(rd.REF_ROOT, "Coop-MH-PPO-scalable.py")).read().split("\n")
    a = next(i for i, l in enumerate(src) if l.startswith("def get_average("))
    b = next(i for i in range(a + 1, len(src)) if src[i].startswith("get_average("))
    printed = []
    ns = {"torch": torch, "np": np, "env": types.SimpleNamespace(nb_car=nb_car, nb_ped=nb_ped), "states": torch.zeros(len(ep_cross), 10),
          "info_co2": lambda *a, **k: None, "LDV": None, "print": lambda *a, **k: printed.append(" ".join(str(x) for x in a))}
    exec(compile("\n".join(src[a:b]), "get_average", "exec"), ns)
    ns["get_average"](torch.as_tensor(ep_car), torch.as_tensor(ep_ped), torch.as_tensor(ep_cross))
    nums = [[float(x) for x in re.findall(r"-?\d+\.\d*(?:e-?\d+)?|nan|-?inf", l.split(":", 1)[-1] if ":" in l else l)] for l in printed]
    return printed, nums


@pytest.mark.parametrize("cfg", [(4, 3, 2), (2, 2, 1)], ids=["432", "221"])
def test_stats_oracle_matches_reference_cell(cfg):
    """oracle/stats_oracle.py against the reference's get_average cell on evaluation episodes of the reference itself."""
    import refppo
    from oracle import stats_oracle as SO
    c, p, l = cfg
    ns = refppo.load_namespace()
    algo, env = refppo.make_algo(ns, "coop_scalable", c, p, l, seed=3)
    env._mh_rng.set_stream(42, 7, 0)
    states = algo.rollout.iterations(algo.actor_net_cross, algo.actor_net_wait, algo.actor_net_choice, 6)[0].numpy()
    Cn = 2 * l
    assert Cn == c                                        # the cell reshapes by env.nb_car (PY:1372): only meaningful when nb_car == 2*nb_lines
    ep_car = states[:, :7 * Cn].reshape(-1, Cn, 7)
    ep_ped = states[:, 7 * Cn + 4:].reshape(-1, p, 9)
    ep_cross = states[:, 7 * Cn]
    printed, nums = _reference_get_average_prints(ep_car, ep_ped, ep_cross, c, p)
    o = SO.get_average(ep_car, ep_ped, ep_cross, c, p)
    flat = [v for row in nums for v in row]
    want = [o["mean_speed"], o["sqrt_speed"], o["mean_acc"], o["sqrt_acc"], o["mean_speed_p"], o["sqrt_speed_p"], o["mean_speed_cars"],
            o["sqrt_speed_cars"], o["interaction_cost"], o["max_finish_compare"], o["mean_temp_cars"], o["std_temp_cars"], o["mean_all_temp_cars"],
            o["std_all_temp_cars"], o["mean_temp_peds"], o["std_temp_peds"], o["yield_decision"], o["go_first_decision"], o["mean_waiting"],
            o["std_waiting"], o["could_stop"]]
    if c == 2:
        want += [o["scenario_11"], o["scenario_m1m1"], o["scenario_m11"], o["scenario_1m1"]]
    # the labels "(voiture 1)", "voiture 1 a 25 metres" contribute the digits 1 / 25 only as integers: the regex keeps decimals only
    assert len(flat) == len(want), (printed, flat, want)
    for got, w in zip(flat, want):
        if np.isnan(w) or np.isinf(w):                   # absent cars have v = 0: their free-flow time (25 - x)/v is inf (PY:1612)
            assert (np.isnan(got) and np.isnan(w)) or got == w
        else:
            assert abs(got - w) <= 6e-3 + 1e-4 * abs(w), (printed, flat, want)      # the cell prints 2 - 4 decimals
