"""Memory / undefined-behaviour check of the kernel logic.  compute-sanitizer is closed on the GPU pool
(profiles/round2_sanitizer_closed_on_pool.txt), so the kernel's own thread body - the shared headers env_step.cuh /
env_core.cuh / env_state.cuh compiled for the host (tests/hostsim) - runs under AddressSanitizer + UBSan instead: whole episodes
with in-step auto-reset, masked resets, state injection and snapshot export / import for every env variant and every template
instantiation the host build has.  Any report aborts the child process."""
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests/hostsim")
import hostsim as H
rng = np.random.default_rng(3)
cases = [("coop_scalable", 4, 3, 2), ("coop_scalable", 1, 1, 1), ("coop_scalable", 6, 4, 4), ("coop", 2, 1, 2), ("coop", 3, 4, 3),
         ("stop", 1, 2, 1), ("naif", 2, 2, 2), ("coop_4cars", 2, 2, 2), ("coop_4cars", 3, 4, 3), ("coop_4cars2", 2, 2, 2), ("coop_4cars2", 1, 1, 1)]
for v, c, p, l in cases:
    for soa in (False, True):
        N = 97
        e = H.HostSimEnv(v, N, c, p, l, seed=5, env_id0=1000, soa=soa)
        e.reset()
        for t in range(170):                      # two auto-resets inside the step
            a = np.empty((N, e.n_action), np.float32)
            a[:, :e.n_action // 2] = rng.uniform(-4, 2, (N, e.n_action // 2)); a[:, e.n_action // 2:] = rng.choice([-1.0, 0.0, 1.0], (N, e.n_action // 2))
            e.step(a, autoreset=True, want_term_obs=(t % 7 == 0))
            if t == 40: e.reset(mask=rng.random(N) < 0.5)
            if t == 60:
                s = e.get_state(); e.set_state(s)
            if t == 90:
                m = rng.random(N) < 0.5
                e.reset_cars(0, rng.uniform(0, 10, N), rng.uniform(-60, -5, N), rng.choice([-1.0, 1.0], N), np.zeros(N), mask=m)
                e.reset_pedestrian(p - 1, rng.uniform(-.05, .05, N), rng.uniform(.75, 1.75, N) * rng.choice([-1.0, 1.0], N), rng.uniform(0, 4, N),
                                   rng.uniform(-3, -.5, N), rng.integers(0, 2, N), rng.integers(0, 3, N), rng.integers(0, 2, N),
                                   rng.uniform(0, 1, N), rng.uniform(0, 1, N), mask=m)
                e.observe()
        del e
print("SANITIZED-OK")
"""


def test_kernel_logic_is_clean_under_asan_and_ubsan():
    asan = sorted(glob.glob("/usr/lib/x86_64-linux-gnu/libasan.so.*"))
    if not asan:
        pytest.skip("no AddressSanitizer runtime in this image")
    sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
    import build as B
    B.build_sanitized()
    env = dict(os.environ, HOSTSIM_SANITIZE="1", LD_PRELOAD=asan[0], ASAN_OPTIONS="detect_leaks=0:abort_on_error=0:halt_on_error=1",
               UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "SANITIZED-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
    assert "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-4000:]
