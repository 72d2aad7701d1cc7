"""Parity tests proper: the sm_100a env kernels, called through the C ABI (ctypes ->
libmhppo_b200.so), against the oracle in HBM semantics on the same seeded inputs, against the
golden fixtures generated from the unmodified reference, and -- at BASELINE.json's full sizes --
through size-independent properties.  Bar: every integer / flag / index bit-exact, fp32 values within
1e-5 relative per step (north_star)."""
import numpy as np
import pytest
import torch

from common import ALL_CONFIGS, ATOL_F32, EXTRA_CONFIGS, FLT_KEYS, INT_KEYS, assert_close, assert_equal, compare_vec_envs, inject_and_compare, random_actions
from conftest import golden_files, load_golden

pytestmark = pytest.mark.gpu
RTOL = 1e-5  # north_star: "fp32 kinematic state, rewards ... within 1e-5 relative per step"


@pytest.fixture(scope="module")
def cuda_env_cls():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from gpu_adapter import CudaEnv
    return CudaEnv


@pytest.mark.parametrize("cfg", ALL_CONFIGS, ids=lambda c: "%s_%d%d%d" % c)
def test_cuda_matches_oracle(oracle_mod, cuda_env_cls, cfg):
    v, c, p, l = cfg
    N = 2048
    ref = oracle_mod.OracleVecEnv(v, N, c, p, l, seed=42, env_id0=5000, store_f32=True)
    got = cuda_env_cls(v, N, c, p, l, seed=42, env_id0=5000)
    compare_vec_envs(ref, got, 165, np.random.default_rng(3), rtol=RTOL, check_state_every=4)


@pytest.mark.parametrize("cfg", EXTRA_CONFIGS, ids=lambda c: "%s_%d%d%d" % c)
def test_every_kernel_instantiation_matches_oracle(oracle_mod, cuda_env_cls, cfg):
    """The configs above leave some (variant, car slots, pedestrian slots) instantiations untouched: one config each,
    a ragged env count (tail CTA), one full episode with auto-reset."""
    v, c, p, l = cfg
    N = 300
    ref = oracle_mod.OracleVecEnv(v, N, c, p, l, seed=9, env_id0=77, store_f32=True)
    got = cuda_env_cls(v, N, c, p, l, seed=9, env_id0=77)
    compare_vec_envs(ref, got, 85, np.random.default_rng(5), rtol=RTOL, check_state_every=5)


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_cuda_tracks_reference_golden(cuda_env_cls, path):
    """Episodes recorded from the unmodified fp64 reference; the kernel keeps fp32 state, so values may
    drift by accumulated fp32 rounding over the 80 steps (tolerance 1e-4) while done stays exact."""
    g = load_golden(path)
    c, p, l = [int(x) for x in g["cfg"]]
    for ep in range(g["actions"].shape[0]):
        env = cuda_env_cls(g["variant"], 1, c, p, l, seed=int(g["seed"]), env_id0=int(g["env_id"][ep]))
        obs = env.reset()
        assert_close("obs", g["obs"][ep, 0], obs[0], 1e-5, 1e-5)
        for t in range(80):
            obs, rew, rl, done = env.step(g["actions"][ep, t][None], autoreset=False)
            ctx = "(episode %d step %d)" % (ep, t)
            assert_close("obs", g["obs"][ep, t + 1], obs[0], 1e-4, 1e-4, ctx)
            assert_close("rewards", g["rewards"][ep, t], rew[0], 1e-4, 1e-4, ctx)
            assert_close("reward_light", g["reward_light"][ep, t], rl[0], 1e-4, 1e-4, ctx)
            assert bool(done[0]) == bool(g["done"][ep, t])


def test_gym_layout_and_device_layout_agree(cuda_env_cls):
    """The kernel accepts either view layout (include/mhppo.h mhppo_view); results must be identical."""
    import ctypes as C
    import mhppo_b200
    from mhppo_b200._lib import View, check
    N = 1000  # not a multiple of the CTA size: exercises the tail
    a_env = cuda_env_cls("coop_scalable", N, 4, 3, 2, seed=7)
    b_env = cuda_env_cls("coop_scalable", N, 4, 3, 2, seed=7)
    a_env.reset(); b_env.reset()
    L = mhppo_b200.lib()
    rng = np.random.default_rng(0)
    e = b_env.env
    obs = torch.zeros(N, e.n_obs, device="cuda"); rew = torch.zeros(N, e.n_lead, device="cuda")
    rl = torch.zeros(N, e.n_lead, device="cuda"); done = torch.zeros(N, dtype=torch.uint8, device="cuda")
    for t in range(90):
        acts = random_actions(rng, N, e.n_action)
        ao, ar, al, ad = a_env.step(acts)
        at = torch.as_tensor(acts).cuda().t().contiguous()  # component-major actions this time
        rm = lambda x: View(x.data_ptr(), x.stride(0), x.stride(1))
        check(L.mhppo_env_step(e._h, View(at.data_ptr(), 1, N), rm(obs), rm(rew), rm(rl), done.data_ptr(), 1,
                               View(None, 0, 0), None))
        torch.cuda.synchronize()
        assert_equal("obs", ao, obs.cpu().numpy()); assert_equal("rew", ar, rew.cpu().numpy())
        assert_equal("rl", al, rl.cpu().numpy()); assert_equal("done", ad, done.cpu().numpy().astype(bool))


@pytest.mark.parametrize("cfg", [("coop", 777, 2, 1, 2), ("coop_scalable", 50001, 4, 3, 2)], ids=["one_slice", "three_slices"])
def test_host_buffer_entry_point(cuda_env_cls, cfg):
    """mhppo_env_step_host (numpy-in / numpy-out convention of env.step) == device entry point; the larger case runs
    through the sliced, three-stream H2D -> kernel -> D2H pipeline (ragged last slice)."""
    v, N, c, p, l = cfg
    a_env = cuda_env_cls(v, N, c, p, l, seed=11)
    b_env = cuda_env_cls(v, N, c, p, l, seed=11)
    o0 = a_env.reset()
    e = b_env.env
    obs_h = torch.zeros(N, e.n_obs).pin_memory(); rew_h = torch.zeros(N, e.n_lead).pin_memory()
    rl_h = torch.zeros(N, e.n_lead).pin_memory(); done_h = torch.zeros(N, dtype=torch.uint8).pin_memory()
    act_h = torch.zeros(N, e.n_action).pin_memory()
    e.reset_host(obs_h)
    assert_equal("obs0", o0, obs_h.numpy())
    rng = np.random.default_rng(1)
    for t in range(85):
        acts = random_actions(rng, N, e.n_action)
        ao, ar, al, ad = a_env.step(acts)
        act_h.copy_(torch.as_tensor(acts))
        e.step_host(act_h, obs_h, rew_h, rl_h, done_h)
        assert_equal("obs", ao, obs_h.numpy()); assert_equal("rew", ar, rew_h.numpy())
        assert_equal("rl", al, rl_h.numpy()); assert_equal("done", ad, done_h.numpy().astype(bool))


def test_state_export_import_roundtrip(cuda_env_cls):
    N = 600
    a_env = cuda_env_cls("coop_4cars2", N, 2, 2, 2, seed=5)
    b_env = cuda_env_cls("coop_4cars2", N, 2, 2, 2, seed=5)
    a_env.reset()
    rng = np.random.default_rng(2)
    for t in range(37):
        a_env.step(random_actions(rng, N, a_env.n_action))
    b_env.set_state(a_env.get_state())
    sa, sb = a_env.get_state(), b_env.get_state()
    for k in INT_KEYS + FLT_KEYS:
        assert_equal(k, sa[k], sb[k])
    for t in range(60):
        acts = random_actions(rng, N, a_env.n_action)
        ra, rb = a_env.step(acts), b_env.step(acts)
        for x, y in zip(ra, rb):
            assert_equal("out", x, y)


def test_masked_reset_only_touches_masked_envs(cuda_env_cls, oracle_mod):
    N = 512
    ref = oracle_mod.OracleVecEnv("stop", N, 2, 3, 2, seed=9, store_f32=True)
    got = cuda_env_cls("stop", N, 2, 3, 2, seed=9)
    ref.reset(); got.reset()
    rng = np.random.default_rng(4)
    for t in range(30):
        a = random_actions(rng, N, ref.n_action)
        ref.step(a.astype(np.float64)); got.step(a)
    mask = rng.random(N) < 0.3
    ro = ref.reset(mask=mask); go = got.reset(mask=mask)
    assert_close("obs", ro[mask], go[mask], RTOL, ATOL_F32)
    sr, sg = ref.get_state(), got.get_state()
    for k in INT_KEYS:
        assert_equal(k, sr[k], sg[k])
    for k in FLT_KEYS:
        assert_close(k, sr[k], sg[k], RTOL, ATOL_F32)


def test_shards_are_slices_of_one_env_set(cuda_env_cls):
    """Rank r holds envs [r*n, (r+1)*n): an R-way sharded run equals one run over the concatenation
    (SURVEY.md 8e), because the Philox stream is keyed by the GLOBAL env id."""
    N = 512
    whole = cuda_env_cls("coop_scalable", N, 4, 3, 2, seed=21, env_id0=0)
    parts = [cuda_env_cls("coop_scalable", N // 2, 4, 3, 2, seed=21, env_id0=r * (N // 2)) for r in range(2)]
    ow = whole.reset(); op = np.concatenate([p.reset() for p in parts])
    assert_equal("obs0", ow, op)
    rng = np.random.default_rng(6)
    for t in range(100):
        a = random_actions(rng, N, whole.n_action)
        rw = whole.step(a)
        rp = [p.step(a[i * (N // 2):(i + 1) * (N // 2)]) for i, p in enumerate(parts)]
        for k in range(4):
            assert_equal("out%d" % k, rw[k], np.concatenate([r[k] for r in rp]))


@pytest.mark.parametrize("cfg,N", [(("coop_scalable", 4, 3, 2), 262144), (("coop_4cars", 2, 2, 2), 65536), (("coop", 2, 1, 2), 4096)],
                         ids=["C4_scalable_262144", "C3_4cars_65536", "C2_coop_4096"])
def test_full_size_properties(oracle_mod, cuda_env_cls, cfg, N):
    """BASELINE.json sizes: (1) every episode is exactly 80 steps; (2) a run is a deterministic function
    of (seed, env ids, actions); (3) a window of envs in the middle of the batch equals the oracle on
    the same global env ids; (4) rewards are <= 0 and finite, flags are 0/1."""
    v, c, p, l = cfg
    W0, WN = N // 2 - 100, 512
    runs = []
    ref = oracle_mod.OracleVecEnv(v, WN, c, p, l, seed=123, env_id0=W0, store_f32=True)
    ref.reset()
    for rep in range(2):
        env = cuda_env_cls(v, N, c, p, l, seed=123, env_id0=0)
        env.reset()
        g = torch.Generator(device="cuda").manual_seed(5)
        trace = []
        for t in range(82):
            a = torch.empty(N, env.n_action, device="cuda")
            a[:, :env.n_action // 2] = torch.rand(N, env.n_action // 2, device="cuda", generator=g) * 6 - 4
            a[:, env.n_action // 2:] = (torch.rand(N, env.n_action // 2, device="cuda", generator=g) < 0.5).float() * 2 - 1
            _, rew, done, _, _ = env.env.step(a)
            assert bool(done.all()) == (t == 79) and bool(done.any()) == (t == 79)
            assert torch.isfinite(rew).all() and (rew <= 0).all()
            obs = env.env.flat_obs
            trace.append((obs[W0:W0 + WN].cpu().numpy(), rew[W0:W0 + WN].cpu().numpy(), env.env.reward_light[W0:W0 + WN].cpu().numpy()))
            if rep == 0:
                ro, rr, rl, rd = ref.step(a[W0:W0 + WN].cpu().numpy().astype(np.float64), autoreset=True)
                assert_close("obs", ro, trace[-1][0], RTOL, ATOL_F32, ctx="(step %d)" % t)
                assert_close("rewards", rr, trace[-1][1], RTOL, ATOL_F32, ctx="(step %d)" % t)
                assert_close("reward_light", rl, trace[-1][2], RTOL, ATOL_F32, ctx="(step %d)" % t)
        runs.append(trace)
        del env
    for a, b in zip(*runs):
        for x, y in zip(a, b):
            assert_equal("determinism", x, y)


def test_snapshot_file_roundtrip(tmp_path):
    """save_state / load_state: an env set restored from a snapshot file continues bit-identically (scenario datasets)."""
    import mhppo_b200
    N = 257
    a = mhppo_b200.VecCrosswalkEnv("coop_scalable", N, nb_car=4, nb_ped=3, nb_lines=2, seed=11, env_id0=40)
    a.reset()
    rng = np.random.default_rng(1)
    for t in range(23):
        a.step(torch.as_tensor(random_actions(rng, N, a.n_action)).cuda())
    a.save_state(str(tmp_path / "snap.npz"))
    b = mhppo_b200.VecCrosswalkEnv("coop_scalable", N, nb_car=4, nb_ped=3, nb_lines=2, seed=11, env_id0=40)
    b.reset()
    b.load_state(str(tmp_path / "snap.npz"))
    for t in range(70):
        act = torch.as_tensor(random_actions(rng, N, a.n_action)).cuda()
        oa, ra, da, _, _ = a.step(act)
        ob, rb, db, _, _ = b.step(act)
        for k in oa:
            assert torch.equal(oa[k], ob[k]), (t, k)
        assert torch.equal(ra, rb) and torch.equal(da, db)
    c = mhppo_b200.VecCrosswalkEnv("coop", N, nb_car=2, nb_ped=1, nb_lines=2, seed=11)
    with pytest.raises(ValueError):
        c.load_state(str(tmp_path / "snap.npz"))


@pytest.mark.parametrize("cfg", [("coop_scalable", 4, 3, 2), ("coop_scalable", 1, 1, 1), ("coop", 2, 2, 2), ("stop", 2, 3, 2), ("naif", 2, 2, 2),
                                 ("coop_4cars", 2, 2, 2), ("coop_4cars2", 2, 2, 2)], ids=lambda c: "%s_%d%d%d" % c)
def test_state_injection_matches_oracle(oracle_mod, cuda_env_cls, cfg):
    """reset_pedestrian / reset_cars / get_state (SC:948-969, NA:897-901) through the C ABI (mhppo_env_reset_pedestrian,
    mhppo_env_reset_cars, mhppo_env_observe) with per-env parameters on half of 2048 envs, then 80 free-running steps
    against the oracle (whose injection is pinned against the unmodified reference, tests/test_oracle_vs_reference.py)."""
    v, c, p, l = cfg
    N = 2048
    ref = oracle_mod.OracleVecEnv(v, N, c, p, l, seed=7, env_id0=900, store_f32=True)
    got = cuda_env_cls(v, N, c, p, l, seed=7, env_id0=900)
    inject_and_compare(ref, got, v, np.random.default_rng(1))


def test_observation_stride_out_of_range_is_rejected():
    """mhppo_env_step keeps the observation component stride in one 32-bit register: a stride >= 2^31 is MHPPO_EINVAL, not a
    silent truncation (include/mhppo.h, mhppo_view)."""
    import mhppo_b200
    from mhppo_b200 import _lib
    from mhppo_b200.vec_env import View, _view
    env = mhppo_b200.VecCrosswalkEnv("coop_scalable", 64, nb_car=4, nb_ped=3, nb_lines=2, seed=1)
    env.reset()
    act = torch.zeros(64, env.n_action, device="cuda")
    good = _view(env._obs, True)
    bad = View(good.ptr, good.env_stride, 1 << 31)
    none = View(None, 0, 0)
    rc = env._L.mhppo_env_step(env._h, _view(act, False), bad, none, none, None, 0, none, env._stream())
    assert rc != 0 and "comp_stride" in env._L.mhppo_last_error().decode()
    with pytest.raises(_lib.MhppoError):
        _lib.check(rc)
    env.step(act)     # the handle stays usable
