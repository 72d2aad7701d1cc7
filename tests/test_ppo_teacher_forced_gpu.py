"""Teacher-forced single-step parity of the four policy kernels (k_choice_act, k_policy_act, k_choice_eval, k_policy_eval)
through the C ABI: identical observations in -> discrete actions bit-exact, means / actions / log-probs / features within
1e-5 (PY:400-453, 177-214).  The states are harvested from the env oracle at several points of an episode (>= 1e5 states
for the headline 4/3/2 config); the other configs cover the 18-, 42- and 54-input choice nets (KP 32 / 56) and the
update kernels of those widths; the shipped 1/1/1 checkpoints are replayed against episodes recorded from the reference."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu

# (nb_car, nb_ped, nb_lines, envs, harvest steps): D = 2 + 6 (2L - 1) + 10 = 30, 18, 42, 54
CONFIGS = [(4, 3, 2, 16384, (0, 3, 8, 15, 25, 40, 70)), (1, 1, 1, 4096, (0, 10, 30)), (6, 2, 3, 4096, (0, 12, 33)), (8, 4, 4, 2048, (0, 9, 21))]


@pytest.fixture(scope="module", params=["auto", "tc"])
def mh(request):
    import mhppo_b200
    mhppo_b200._lib.check(mhppo_b200.lib().mhppo_set_mlp_mode({"auto": 0, "tc": 2}[request.param]))
    mhppo_b200.mlp_mode = request.param
    yield mhppo_b200
    assert mhppo_b200.lib().mhppo_tc_failures() == 0
    mhppo_b200.lib().mhppo_set_mlp_mode(0)


def _nets(mh, D, seed=0):
    torch.manual_seed(seed)
    cross, wait, choice = mh.Model_PPO(13, 1, 1, mean=-1.0, std=3.0), mh.Model_PPO(13, 1, 1, mean=-1.0, std=3.0), mh.Model_PPO(D, 2, 2)
    # default init keeps tanh far from saturation; scale the last layers (moderately: the output error scales with it) so means / probabilities spread over their range
    for n, s in ((cross, 2.0), (wait, 2.0), (choice, 4.0)):
        sd = n.state_dict(); sd["layer4.weight"] = sd["layer4.weight"] * s; n.load_state_dict(sd)
    return cross, wait, choice


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: "%d%d%d" % c[:3])
def test_policy_kernels_teacher_forced(mh, oracle_mod, cfg):
    from oracle import ppo_oracle as PO
    from mhppo_b200._lib import RolloutCfg, check
    nb_car, P, L, N, steps = cfg
    Cn, D = 2 * L, 2 + 6 * (2 * L - 1) + 10
    seed, id0, it = 99, 700, 3
    cross, wait, choice = _nets(mh, D)
    sds = [{k: v.clone() for k, v in n.state_dict().items()} for n in (cross, wait, choice)]
    venv = oracle_mod.OracleVecEnv("coop_scalable", N, nb_car, P, L, seed=seed, env_id0=id0, store_f32=True)
    obs = venv.reset()
    rng = np.random.default_rng(5)
    rcfg = RolloutCfg(nb_ped=P, nb_lines=L, T=1, n_envs=N, seed=seed, env_id0=id0)
    lib = mh.lib()
    check(lib.mhppo_set_gaussian_head(-1.0, 3.0, 0.5, 2.0))
    dev = torch.device("cuda")
    f = lambda *s: torch.zeros(*s, device=dev)
    action_d, light = torch.zeros(Cn * P, N, dtype=torch.int8, device=dev), f(Cn, N)
    obs_d, act_d, logp_d = f(D, Cn * N), f(Cn * N), f(Cn * N)
    actions, obs_c, act, logp = f(2 * Cn, N), f(13, Cn * N), f(Cn * N), f(Cn * N)
    env_ids = np.arange(id0, id0 + N)
    n_states, t_now = 0, 0
    loose = mh.mlp_mode == "tc"
    for t_h in steps:
        while t_now < t_h:
            a = np.concatenate([rng.uniform(-4, 2, (N, Cn)), rng.choice([-1.0, 1.0], (N, Cn))], axis=1)
            obs = venv.step(a)[0]; t_now += 1
        obs = np.asarray(obs, np.float32)
        od = torch.as_tensor(obs.T.copy()).to(dev)
        # ---- k_choice_act (PY:400-428): Categorical sample per (car, pedestrian), light of the closest pedestrian
        u = np.stack([PO.policy_uniform(k, env_ids, seed, it) for k in range(Cn * P)], axis=1)
        acts, logps, feats, la, llp, lf, _ = PO.discrete_step(obs, sds[2], u.astype(np.float32), P, L)
        check(lib.mhppo_choice_act(C.byref(rcfg), od.data_ptr(), choice.flat.data_ptr(), it, action_d.data_ptr(), light.data_ptr(),
                                   obs_d.data_ptr(), act_d.data_ptr(), logp_d.data_ptr(), None))
        torch.cuda.synchronize()
        got = (action_d.cpu().numpy().T.astype(np.int64) + 1) // 2
        with torch.no_grad():
            p0 = np.stack([PO.mlp_forward(sds[2], feats[:, k], 2).numpy()[:, 0] for k in range(Cn * P)], axis=1)
        bad = got != acts
        assert (np.abs(p0 - u)[bad] < 1e-6).all() and bad.sum() <= 2, "choice sample differs away from u == p0: %d" % bad.sum()
        ok = ~bad.reshape(N, Cn, P).any(axis=2)                          # envs x cars whose decisions all agree
        np.testing.assert_array_equal(act_d.view(Cn, N).cpu().numpy().T[ok], la[ok])
        np.testing.assert_array_equal(light.cpu().numpy().T[ok], (2 * la - 1)[ok])
        # inputs reach 1e3 (absent cars sit at x = -1000, SC:654): two fp32 summation orders differ by ~1e-4 in a logit
        np.testing.assert_allclose(logp_d.view(Cn, N).cpu().numpy().T[ok], llp[ok], rtol=1e-4, atol=5e-5)
        np.testing.assert_allclose(obs_d.view(D, Cn, N).cpu().numpy().transpose(2, 1, 0)[ok], lf[ok], rtol=1e-6, atol=1e-6)
        # ---- k_policy_act (PY:434-453) on the oracle's decisions: min over existing pedestrians, N(mean, 0.5) sample, log-prob
        ad = (2 * acts - 1).astype(np.float64)
        action_d.copy_(torch.as_tensor(ad.T.astype(np.int8)).to(dev))
        light.copy_(torch.as_tensor((2 * la - 1).T.astype(np.float32)).to(dev))
        z = np.stack([PO.policy_normal(0 * Cn + i, env_ids, seed, it) for i in range(Cn)], axis=1)
        mean, a, lp, st = PO.continuous_step(obs, ad, sds[0], sds[1], z, P, L)
        check(lib.mhppo_policy_act(C.byref(rcfg), od.data_ptr(), cross.flat.data_ptr(), wait.flat.data_ptr(), action_d.data_ptr(),
                                   light.data_ptr(), 0, it, actions.data_ptr(), obs_c.data_ptr(), act.data_ptr(), logp.data_ptr(), None))
        torch.cuda.synchronize()
        # 3xTF32 (tc mode) carries ~5e-7 of the largest product: rows of absent cars hold x = -1000 (SC:654)
        tol = dict(rtol=1e-5, atol=1e-5) if not loose else dict(rtol=2e-3, atol=5e-3)
        np.testing.assert_allclose(act.view(Cn, N).cpu().numpy().T, a, **tol)
        np.testing.assert_allclose(logp.view(Cn, N).cpu().numpy().T, lp, rtol=tol["rtol"], atol=max(tol["atol"], 2e-5))
        np.testing.assert_array_equal(actions[:Cn].cpu().numpy(), act.view(Cn, N).cpu().numpy())
        np.testing.assert_array_equal(actions[Cn:].cpu().numpy().T, 2 * la - 1)
        # stored features = those of the arg-min pedestrian (PY:448-450); a tie between two pedestrians' means within fp32
        # rounding (both saturated) picks by rounding noise in the reference too: such rows must be rare
        oc = obs_c.view(13, Cn, N).cpu().numpy().transpose(2, 1, 0)
        differs = (np.abs(oc - st) > 1e-6 + 1e-6 * np.abs(st)).any(axis=2)
        assert differs.mean() < (2e-3 if not loose else 2e-2), differs.mean()
        # ---- k_choice_eval (PY:177-192): argmax decisions; k_policy_eval (PY:195-214)
        ea = PO.eval_discrete(obs, sds[2], P, L)
        check(lib.mhppo_choice_eval(C.byref(rcfg), od.data_ptr(), choice.flat.data_ptr(), 1, action_d.data_ptr(), None))
        torch.cuda.synchronize()
        got = (action_d.cpu().numpy().T.astype(np.int64) + 1) // 2
        bad = got != ea
        assert (np.abs(p0 - 0.5)[bad] < 1e-6).all() and bad.sum() <= 2, "argmax differs away from a tie: %d" % bad.sum()
        ead = (2 * ea - 1).astype(np.float64)
        action_d.copy_(torch.as_tensor(ead.T.astype(np.int8)).to(dev))
        want = PO.eval_continuous(obs, ead, sds[0], sds[1], P, L)
        check(lib.mhppo_policy_eval(C.byref(rcfg), od.data_ptr(), cross.flat.data_ptr(), wait.flat.data_ptr(), action_d.data_ptr(),
                                    0.3, 10.0, -4.0, 2.0, actions.data_ptr(), act.data_ptr(), None))
        torch.cuda.synchronize()
        np.testing.assert_allclose(act.view(Cn, N).cpu().numpy().T, want, **tol)
        np.testing.assert_array_equal(actions[Cn:].cpu().numpy().T, ead[:, :Cn])           # lights = first C decisions (PY:192, 216)
        n_states += N
    if (nb_car, P, L) == (4, 3, 2):
        assert n_states >= 100000


def test_gaussian_head_parameters_reach_the_kernels(mh, oracle_mod):
    """mean / std of the actors and Algo_PPO.value_std are honoured by rollout and update (PY:88-90, 726-729, 1044-1045):
    a non-default head gives the actions the oracle computes with the same parameters."""
    from oracle import ppo_oracle as PO
    N, seed, id0 = 2048, 5, 40
    env = mh.VecCrosswalkEnv("coop_scalable", N, nb_car=4, nb_ped=3, nb_lines=2, seed=seed, env_id0=id0)
    torch.manual_seed(1)
    algo = mh.Algo_PPO(mh.Model_PPO, env, num_states_c=13, num_states_d=30, num_actions=1, mean=-0.5, std=2.5, nb_cars=4, dt=0.3, value_std=0.3)
    r = algo.rollout
    sds = [{k: v.clone() for k, v in n.state_dict().items()} for n in (algo.actor_net_cross, algo.actor_net_wait, algo.actor_net_choice)]
    r.iterations_rand(algo.actor_net_cross, algo.actor_net_wait, algo.actor_net_choice)
    venv = oracle_mod.OracleVecEnv("coop_scalable", N, 4, 3, 2, seed=seed, env_id0=id0, store_f32=True)
    obs = np.asarray(venv.reset(), np.float32)
    ad = r.action_d.cpu().numpy().T.astype(np.float64)
    Cn, P = 4, 3
    mean = np.full((N, Cn), 2.0, np.float32)
    for i in range(Cn):
        for p in range(P):
            fe, ex = PO.obs_car_ped(obs, i, p, P, 2)
            with torch.no_grad():
                m = np.where(ad[:, i * P + p] <= 0, PO.mlp_forward(sds[0], fe, 1, -0.5, 2.5).numpy()[:, 0], PO.mlp_forward(sds[1], fe, 1, -0.5, 2.5).numpy()[:, 0])
            mean[:, i] = np.where(ex != 0, np.minimum(mean[:, i], m), mean[:, i])
    z = np.stack([PO.policy_normal(i, np.arange(id0, id0 + N), seed, 0) for i in range(Cn)], axis=1)
    a = mean + np.float32(np.sqrt(0.3)) * z.astype(np.float32)
    lp = -((a - mean) ** 2) / np.float32(0.6) - np.float32(0.5 * np.log(2 * np.pi * 0.3))
    # 1e-5, with an absolute floor of 2e-5 for the rows of absent cars (x = -1000 features, fp32 summation order)
    tol = dict(rtol=1e-5, atol=2e-5) if mh.mlp_mode != "tc" else dict(rtol=2e-3, atol=5e-3)
    np.testing.assert_allclose(r.act.view(r.T, Cn, N)[0].cpu().numpy().T, a, **tol)
    np.testing.assert_allclose(r.logp.view(r.T, Cn, N)[0].cpu().numpy().T, lp, rtol=tol["rtol"], atol=max(tol["atol"], 5e-5))
    algo.update(epochs=1)                                               # runs with the same head; must stay finite
    assert torch.isfinite(algo.actor_net_cross.flat).all()


def _sd(z, prefix):
    return {k[len(prefix) + 1:]: torch.as_tensor(z[k]) for k in z.files if k.startswith(prefix + ".")}


def test_shipped_checkpoint_111_eval_matches_reference(mh):
    """The shipped trained nets (load_model/weights/pappo-scalable-coop-*-111-*-step-1000.pth, 18-input choice net) on the
    1/1/1 scalable env: the GPU evaluation rollout reproduces episodes recorded from the unmodified reference."""
    z = np.load(os.path.join(GOLDEN_DIR, "ppo_eval_ckpt111.npz"))
    tol = dict(rtol=1e-4, atol=1e-4) if mh.mlp_mode != "tc" else dict(rtol=2e-3, atol=5e-3)
    for e, (seed, env_id) in enumerate(z["streams"]):
        env = mh.VecCrosswalkEnv("coop_scalable", 1, nb_car=1, nb_ped=1, nb_lines=1, seed=int(seed), env_id0=int(env_id))
        algo = mh.Algo_PPO(mh.Model_PPO, env, num_algo=111, num_states_c=13, num_states_d=18, num_actions=1, mean=-1.0, std=3.0, nb_cars=1, dt=0.3)
        for name, net in (("cross", algo.actor_net_cross), ("wait", algo.actor_net_wait), ("choice", algo.actor_net_choice)):
            net.load_state_dict(_sd(z, name))
        out = algo.rollout.iterations(algo.actor_net_cross, algo.actor_net_wait, algo.actor_net_choice, 1)
        want = {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith("ep%d." % e)}
        np.testing.assert_array_equal(out["action_d"][0, :, :, 0].cpu().numpy(), want["actions"][:, 2:])
        np.testing.assert_allclose(out["obs"][0, :, :, 0].cpu().numpy(), want["obs"], **tol)
        np.testing.assert_allclose(out["acts"][0, :, :, 0].cpu().numpy(), want["acts"], **tol)
        np.testing.assert_allclose(out["rews"][0, :, :, 0].cpu().numpy(), want["rew"], **tol)
        np.testing.assert_allclose(out["reward_light"][0, :, :, 0].cpu().numpy(), want["rl"], **tol)


@pytest.mark.parametrize("L", [1, 2, 3, 4], ids=lambda l: "D%d" % (2 + 6 * (2 * l - 1) + 10))
def test_choice_head_loss_and_gradient_match_literal_MxM_autograd(mh, L):
    """train_model_d (PY:818-851) for every choice-net width (18 / 30 / 42 / 54 inputs -> KP 32 / 32 / 56 / 56): the O(M)
    kernel's loss and flat gradients (actor head 2 and its critic) against torch autograd on the reference's literal (M, M)
    broadcast form."""
    from oracle import ppo_oracle as PO
    Cn, D, N = 2 * L, 2 + 6 * (2 * L - 1) + 10, 300
    env = mh.VecCrosswalkEnv("coop_scalable", N, nb_car=min(4, Cn), nb_ped=2, nb_lines=L, seed=2)
    torch.manual_seed(7)
    algo = mh.Algo_PPO(mh.Model_PPO, env, num_states_c=13, num_states_d=D, num_actions=1, mean=-1.0, std=3.0, nb_cars=4, dt=0.3)
    r = algo.rollout
    g = torch.Generator(device="cuda").manual_seed(3)
    r.obs_d.copy_(torch.randn(r.obs_d.shape, device="cuda", generator=g))
    r.act_d.copy_(torch.randint(0, 2, (r.M,), device="cuda", generator=g).float())
    r.logp_d.copy_(-torch.rand(r.M, device="cuda", generator=g) * 1.5)
    r.rew_d.copy_(torch.randn(r.M, device="cuda", generator=g) * 4 - 2)
    r.exist.copy_(torch.randint(0, 2, r.exist.shape, device="cuda", generator=g).to(torch.int8))
    actor, critic = algo.actor_net_choice, algo.critic_net_choice
    a_o, c_o = PO.Net(D, 2, 2), PO.Net(D, 1, 0)
    a_o.load_state_dict(actor.state_dict()); c_o.load_state_dict(critic.state_dict())
    sel = (r.exist.view(-1) != 0).cpu()
    s = r.obs_d.t().cpu()[sel]
    V = c_o(s).reshape(-1); rtg = r.rew_d.cpu()[sel]
    adv = rtg - V; adv = ((adv - adv.mean()) / (adv.std() + 1e-10)).detach()
    probs = a_o(s).reshape(-1, 2)
    logp = torch.distributions.Categorical(probs).log_prob(r.act_d.cpu()[sel].reshape(-1, 1))      # (M, M), PY:834-837
    ratio = torch.exp(logp - r.logp_d.cpu()[sel])
    la = (-torch.min(ratio * adv, torch.clamp(ratio, 0.8, 1.2) * adv)).mean()
    lc = torch.nn.functional.mse_loss(V, rtg)
    ga = torch.autograd.grad(la, list(a_o.parameters())); gc = torch.autograd.grad(lc, list(c_o.parameters()))
    assert algo.train_model_d(actor, critic, algo.optimizer_actor_choice, algo.optimizer_critic_choice)
    torch.cuda.synchronize()
    np.testing.assert_allclose(algo._loss.cpu().numpy(), [float(la), float(lc)], rtol=1e-5)
    for net, grads, ref, n_out in ((actor, ga, a_o, 2), (critic, gc, c_o, 1)):
        tmp = mh.Model_PPO(D, n_out, net.model_type, device="cpu")
        tmp.load_state_dict({k: gg for (k, _), gg in zip(ref.named_parameters(), grads)})
        want, got = tmp.flat.numpy(), net.grad.cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5 * np.abs(want).max())
    # the stand-alone critic forward of that width (k_value_stats<KP>) gives the same V and statistics
    from mhppo_b200._lib import check
    idx, K = algo._selection(None)
    st2, V2 = torch.zeros(3, dtype=torch.float64, device="cuda"), torch.zeros_like(r.V_d)
    check(mh.lib().mhppo_value_stats(D, r.obs_d.data_ptr(), D, r.M, idx.data_ptr(), K, r.M, c_o_flat(mh, c_o, D).data_ptr(), r.rew_d.data_ptr(),
                                     V2.data_ptr(), st2.data_ptr(), algo._ws.data_ptr(), None))
    torch.cuda.synchronize()
    np.testing.assert_allclose(V2.cpu()[sel].numpy(), V.detach().numpy(), rtol=1e-5, atol=1e-5)
    A = (rtg - V.detach()).double()
    np.testing.assert_allclose(st2.cpu().numpy(), [float(A.sum()), float((A * A).sum()), float(sel.sum())], rtol=1e-5)


def c_o_flat(mh, net, D):
    m = mh.Model_PPO(D, 1, 0)
    m.load_state_dict(net.state_dict())
    return m.flat


def test_two_rank_update_equals_one_rank(mh):
    """N-GPU run == 1-GPU run on the concatenated env set: spawns tests/dist_ppo_check.py on 2 ranks."""
    if mh.mlp_mode != "auto":
        pytest.skip("one mode is enough for the distributed check")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(here, "dist_ppo_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "OK" in r.stdout
