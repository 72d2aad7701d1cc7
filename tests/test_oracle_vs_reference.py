"""Live pin of the oracle against the unmodified reference (only where /root/reference exists, i.e.
the build container; skipped on the GPU box, where the golden fixtures stand in)."""
import numpy as np
import pytest

import refdriver as rd

pytestmark = pytest.mark.skipif(not rd.available(), reason="reference sources not present")


@pytest.mark.parametrize("cfg", [("coop_scalable", 4, 3, 2), ("coop", 2, 2, 2), ("stop", 2, 3, 2), ("naif", 2, 2, 2),
                                 ("coop_4cars", 2, 2, 2), ("coop_4cars2", 2, 2, 2)], ids=lambda c: "%s_%d%d%d" % c)
def test_live_reference_episode(cfg):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import pin_oracle
    rng = np.random.default_rng(hash(cfg) % 2**32)
    for ep in range(4):
        errs = pin_oracle.compare_episode(*cfg, seed=77 + ep, env_id=5 * ep + 1, rng=rng, light_mode=["episode", "step"][ep % 2])
        assert not errs, errs[:3]
