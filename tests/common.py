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


def inject_and_compare(ref_env, got_env, variant, rng, n_steps=80, rtol=1e-5, to_np=lambda x: np.asarray(x)):
    """State injection (reset_pedestrian / reset_cars / get_state, SC:948-969) with per-env parameters into half of the envs
    of two vectorised envs after a common reset, then `n_steps` free-running steps: ints bit-exact, floats within rtol."""
    N = ref_env.N
    ref_env.reset(); got_env.reset()
    mask = (np.arange(N) % 2 == 0).astype(np.uint8)
    f32 = lambda a: np.asarray(a, np.float32)
    if variant == "naif":
        cross = f32(rng.uniform(2.5, 3.0, N))
        for j in range(ref_env.P):
            vals = (f32(rng.uniform(-0.05, 0.05, N)), f32(rng.uniform(0.75, 1.75, N)), f32(rng.uniform(0, 4, N)), f32(rng.uniform(-3, -0.5, N)),
                    0.0, f32(rng.choice([-1.0, 1.0], N)), cross, 0.0, 0.0)
            ref_env.reset_pedestrian(j, *vals[:7], mask=mask); got_env.reset_pedestrian(j, *vals[:7], mask=mask)
    else:
        d = f32(rng.choice([-1.0, 1.0], N))
        vals = (f32(rng.uniform(-0.05, 0.05, N)), f32(rng.uniform(0.75, 1.75, N)) * d, f32(rng.uniform(0, 4, N)), f32(rng.uniform(-3.5, -1.0, N)) * d,
                0.0, 0.0, 0.0, 1.0, d)
        ref_env.reset_pedestrian(0, *vals, mask=mask); got_env.reset_pedestrian(0, *vals, mask=mask)
    for i in range(min(2, ref_env.C)):
        lanes = int(ref_env.cfg.nb_lines)
        cv = (f32(rng.uniform(5, 10, N)), f32(rng.uniform(-50, -15, N)), 0.0, float(i % 2) if lanes > 1 else 0.0)   # a valid lane: line < nb_lines
        ref_env.reset_cars(i, *cv, mask=mask); got_env.reset_cars(i, *cv, mask=mask)
    assert_close("obs(get_state)", ref_env.observe(), to_np(got_env.observe()), rtol, ATOL_F32)
    for t in range(n_steps):
        a = random_actions(rng, N, ref_env.n_action)
        ro, rr, rl, rd = ref_env.step(a.astype(np.float64), autoreset=False)
        go, gr, gl, gd = got_env.step(a, autoreset=False)
        ctx = "(step %d after injection)" % t
        assert_close("obs", ro, to_np(go), rtol, ATOL_F32, ctx=ctx)
        assert_close("rewards", rr, to_np(gr), rtol, ATOL_F32, ctx=ctx)
        assert_close("reward_light", rl, to_np(gl), rtol, ATOL_F32, ctx=ctx)
        if t % 8 == 0 or t == n_steps - 1:
            sr, sg = ref_env.get_state(), got_env.get_state()
            for k in INT_KEYS:
                assert_equal("state." + k, sr[k], to_np(sg[k]), ctx)
            for k in FLT_KEYS:
                assert_close("state." + k, sr[k], to_np(sg[k]), rtol, ATOL_F32, ctx=ctx)
