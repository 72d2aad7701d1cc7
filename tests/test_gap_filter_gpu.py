"""The step kernel decides `car_time + light < CG` (SC:167-170 with CG_score SC:419-428) through a three-level filtered
predicate (env_step.cuh gap_eval): no draw / fp32 draw on the SFU / exact fp64.  Every level must agree with the exact
decision.  mhppo_gap_selftest runs the kernel's own gap_eval next to the exact evaluation; the inputs put the time gap on the
critical gap and at relative distances from 1e-9 to 3e-3 either side of it, where a wrong error bound would show."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(lib, dx, vden, light, size, v0y, gender, age, ctr, env_id=12345, k0=0x9E3779B9, k1=0x1234567):
    n = dx.numel()
    fast = torch.empty(n, dtype=torch.uint8, device="cuda"); level = torch.empty_like(fast); exact = torch.empty_like(fast)
    cg = torch.empty(n, dtype=torch.float64, device="cuda")
    import mhppo_b200
    mhppo_b200._lib.check(lib.mhppo_gap_selftest(n, dx.data_ptr(), vden.data_ptr(), light.data_ptr(), size.data_ptr(), v0y.data_ptr(),
                                                 gender.data_ptr(), age.data_ptr(), ctr.data_ptr(), env_id, k0, k1, fast.data_ptr(),
                                                 level.data_ptr(), exact.data_ptr(), cg.data_ptr(), None))
    torch.cuda.synchronize()
    return fast, level, exact, cg


def test_filtered_gap_decision_equals_exact_decision():
    import mhppo_b200
    lib = mhppo_b200.lib()
    g = torch.Generator(device="cuda").manual_seed(3)
    n = 1 << 22
    U = lambda lo, hi: torch.rand(n, generator=g, device="cuda", dtype=torch.float64) * (hi - lo) + lo
    vden = torch.where(U(0, 1) < 0.3, torch.full((n,), 0.01, device="cuda", dtype=torch.float64), U(0.01, 12.0))   # Vc + 0.01, stopped cars included
    light = torch.randint(-1, 2, (n,), generator=g, device="cuda").double()
    size = torch.randint(1, 5, (n,), generator=g, device="cuda").double() * U(2.5, 4.5)                              # |line_pos - line| * cross
    v0y = U(0.4, 2.2) * torch.where(U(0, 1) < 0.5, -1.0, 1.0)
    gender = torch.randint(0, 2, (n,), generator=g, device="cuda", dtype=torch.int32)
    age = torch.randint(0, 3, (n,), generator=g, device="cuda", dtype=torch.int32)
    ctr = torch.randint(0, 2 ** 31 - 1, (n,), generator=g, device="cuda", dtype=torch.int32)      # read as uint32 by the kernel
    args = (vden, light, size, v0y, gender, age, ctr)
    # pass 1: free-running gaps (what the env produces), and the exact critical gap of every sample
    dx = U(-120.0, 0.0)
    fast, level, exact, cg = _run(lib, dx, *args)
    assert torch.equal(fast, exact)
    share = [float((level == k).double().mean()) for k in range(3)]
    assert share[2] < 5e-3, share            # the exact fallback stays rare
    # pass 2: the time gap sits at relative distance eps from the critical gap
    worst = 0.0
    for eps in (0.0, 1e-9, 1e-7, 1e-6, 1e-5, 1e-4, 2e-4, 3e-4, 5e-4, 1e-3, 3e-3):
        for sign in (-1.0, 1.0):
            lhs = cg * (1.0 + sign * eps)
            dxa = -(lhs - light) * vden       # |dx / vden| + light = lhs (up to rounding); lhs - light > 0 whenever CG > 1
            ok = (lhs - light) > 0
            fast, level, exact, _ = _run(lib, torch.where(ok, dxa, dx), *args)
            assert torch.equal(fast, exact), (eps, sign, int((fast != exact).sum()))
            worst = max(worst, float((level == 2).double().mean()))
    assert worst > 0.5                        # on the boundary the filter does hand over to the exact path
