"""Dev tool: host build of the kernel logic vs the oracle in HBM semantics (store_f32=1)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
from oracle import oracle as O
import hostsim as H

CONFIGS = [("coop_scalable", 4, 3, 2), ("coop_scalable", 1, 1, 1), ("coop_scalable", 8, 4, 4), ("coop_scalable", 3, 2, 3),
           ("coop", 2, 1, 2), ("coop", 3, 3, 3), ("stop", 1, 2, 1), ("stop", 2, 3, 2), ("naif", 1, 2, 1), ("naif", 3, 3, 2),
           ("coop_4cars", 2, 2, 2), ("coop_4cars", 1, 1, 1), ("coop_4cars2", 2, 2, 2), ("coop_4cars2", 3, 2, 3)]

def run(variant, c, p, l, N=512, T=170, seed=99, soa=False):
    o = O.OracleVecEnv(variant, N, c, p, l, seed=seed, env_id0=1000, store_f32=True)
    h = H.HostSimEnv(variant, N, c, p, l, seed=seed, env_id0=1000, soa=soa)
    rng = np.random.default_rng(5)
    oo, ho = o.reset(), h.reset()
    worst = 0.0
    def cmp(name, a, b, t, exact=False):
        nonlocal worst
        a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
        if exact:
            if not np.array_equal(a, b):
                idx = np.argwhere(a != b)[0]; raise SystemExit("%s %s t=%d idx=%s %r %r" % (variant, name, t, idx, a[tuple(idx)], b[tuple(idx)]))
            return
        err = np.abs(a - b) / np.maximum(np.abs(a), 1e-30)
        err[a == b] = 0
        worst = max(worst, err.max())
        if err.max() > 1e-6:
            idx = np.unravel_index(err.argmax(), err.shape); raise SystemExit("%s %s t=%d idx=%s %r %r" % (variant, name, t, idx, a[idx], b[idx]))
    cmp("obs0", oo, ho, -1)
    for t in range(T):
        A = o.n_action
        acts = np.zeros((N, A), np.float32)
        acts[:, :A // 2] = rng.uniform(-4, 2, (N, A // 2))
        acts[:, A // 2:] = rng.choice([-1.0, 0.0, 1.0], (N, A // 2))
        ob, orw, orl, od = o.step(acts.astype(np.float64), autoreset=True)
        hb, hrw, hrl, hd = h.step(acts, autoreset=True)
        cmp("done", od, hd, t, True); cmp("obs", ob, hb, t); cmp("rew", orw, hrw, t); cmp("rl", orl, hrl, t)
        so, sh = o.get_state(), h.get_state()
        for k in ("car_i", "ped_i", "env_i"): cmp(k, so[k], sh[k], t, True)
        for k in ("car_f", "ped_f", "env_f"): cmp(k, so[k], sh[k], t)
    return worst

if __name__ == "__main__":
    for cfg in CONFIGS:
        print(cfg, "max rel err %.2e" % run(*cfg, soa=(cfg[1] % 2 == 0)))
