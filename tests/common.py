"""Shared helpers of the parity tests."""
import numpy as np

ALL_CONFIGS = [
    ("coop_scalable", 4, 3, 2), ("coop_scalable", 1, 1, 1), ("coop_scalable", 8, 4, 4), ("coop_scalable", 3, 2, 3),
    ("coop", 2, 1, 2), ("coop", 3, 3, 3), ("stop", 1, 2, 1), ("stop", 2, 3, 2), ("naif", 1, 2, 1), ("naif", 3, 3, 2),
    ("coop_4cars", 2, 2, 2), ("coop_4cars", 1, 1, 1), ("coop_4cars2", 2, 2, 2), ("coop_4cars2", 3, 2, 3),
]
# one more config per remaining (variant, car slots, pedestrian slots) kernel instantiation (csrc/env_inst_*.cu)
EXTRA_CONFIGS = [
    ("coop_scalable", 2, 4, 1), ("coop_scalable", 3, 4, 2), ("coop", 2, 3, 2), ("coop", 6, 4, 3), ("stop", 3, 4, 2), ("stop", 5, 3, 3),
    ("naif", 2, 4, 2), ("naif", 7, 2, 4), ("coop_4cars", 2, 3, 2), ("coop_4cars", 3, 4, 3), ("coop_4cars", 4, 2, 4),
    ("coop_4cars2", 2, 4, 2), ("coop_4cars2", 4, 3, 2),
]
INT_KEYS = ("car_i", "ped_i", "env_i")
FLT_KEYS = ("car_f", "ped_f", "env_f")


def random_actions(rng, n, n_action, light_values=(-1.0, 0.0, 1.0)):
    """acc ~ U(-4.5, 2.5) (exercises the clamp), light from `light_values`; fp32-representable."""
    a = np.zeros((n, n_action), np.float32)
    a[:, :n_action // 2] = rng.uniform(-4.5, 2.5, (n, n_action // 2))
    a[:, n_action // 2:] = rng.choice(np.asarray(light_values, np.float32), (n, n_action // 2))
    return a


def rel_err(ref, got):
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64)
    err = np.abs(ref - got) / np.maximum(np.abs(ref), 1e-30)
    err[ref == got] = 0.0
    return err


def assert_close(name, ref, got, rtol, atol=0.0, ctx=""):
    """|ref - got| <= atol + rtol*|ref|.  When `got` is fp32 (kernel outputs) the fp64 reference value is
    first rounded to fp32, so magnitudes below the fp32 range compare as the zeros they are stored as."""
    got_is_f32 = np.asarray(got).dtype == np.float32
    ref = np.asarray(ref, np.float64)
    if got_is_f32:
        with np.errstate(over="ignore", under="ignore"):
            ref = ref.astype(np.float32).astype(np.float64)
    got = np.asarray(got, np.float64)
    assert ref.shape == got.shape, "%s shape %s vs %s %s" % (name, ref.shape, got.shape, ctx)
    bad = np.abs(ref - got) > (atol + rtol * np.abs(ref))
    if bad.any():
        idx = tuple(np.argwhere(bad)[0])
        raise AssertionError("%s mismatch at %s: ref %r got %r (%d bad) %s" % (name, idx, ref[idx], got[idx], bad.sum(), ctx))


def assert_equal(name, ref, got, ctx=""):
    ref, got = np.asarray(ref), np.asarray(got)
    assert ref.shape == got.shape, "%s shape %s vs %s %s" % (name, ref.shape, got.shape, ctx)
    if not np.array_equal(ref, got):
        idx = tuple(np.argwhere(ref != got)[0])
        raise AssertionError("%s mismatch at %s: ref %r got %r (%d bad) %s" % (name, idx, ref[idx], got[idx], (ref != got).sum(), ctx))


ATOL_F32 = 1e-30  # below ~1.2e-38 fp32 is denormal and no longer carries 1e-5 relative precision itself;
                  # exp(-4*dl) shaping terms reach that range, so values that small compare absolutely


def compare_vec_envs(ref_env, got_env, n_steps, rng, rtol=1e-5, autoreset=True, to_np=lambda x: np.asarray(x),
                     light_values=(-1.0, 0.0, 1.0), check_state_every=1, atol=ATOL_F32):
    """Free-run two vectorised envs on the same actions; flags/ints bit-exact, floats within rtol."""
    o0, g0 = ref_env.reset(), got_env.reset()
    assert_close("obs(reset)", o0, to_np(g0), rtol, atol)
    for t in range(n_steps):
        a = random_actions(rng, ref_env.N, ref_env.n_action, light_values)
        ro, rr, rl, rd = ref_env.step(a.astype(np.float64), autoreset=autoreset)
        go, gr, gl, gd = got_env.step(a, autoreset=autoreset)
        ctx = "(step %d)" % t
        assert_equal("done", rd, to_np(gd), ctx)
        assert_close("obs", ro, to_np(go), rtol, atol, ctx=ctx)
        assert_close("rewards", rr, to_np(gr), rtol, atol, ctx=ctx)
        assert_close("reward_light", rl, to_np(gl), rtol, atol, ctx=ctx)
        if check_state_every and (t % check_state_every == 0 or t == n_steps - 1):
            sr, sg = ref_env.get_state(), got_env.get_state()
            for k in INT_KEYS:
                assert_equal("state." + k, sr[k], to_np(sg[k]), ctx)
            for k in FLT_KEYS:
                assert_close("state." + k, sr[k], to_np(sg[k]), rtol, atol, ctx=ctx)
