"""tcgen05 / TMEM building blocks (mh-ppo_b200/csrc/tc.cuh): one-CTA GEMM against torch fp64."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(64, 32), (32, 64), (64, 16), (16, 32)])
@pytest.mark.parametrize("mode", [0, 1])
def test_tcgen05_gemm_selftest(N, K, mode):
    import mhppo_b200
    g = torch.Generator(device="cuda").manual_seed(N * 100 + K + mode)
    A = torch.randn(128, K, device="cuda", generator=g)
    B = torch.randn(N, K, device="cuda", generator=g)
    D = torch.zeros(128, N, device="cuda")
    mhppo_b200._lib.check(mhppo_b200.lib().mhppo_tc_selftest(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, mode, None))
    torch.cuda.synchronize()
    want = (A.double() @ B.double().t()).cpu().numpy()
    got = D.cpu().numpy()
    assert np.isfinite(got).all(), "tcgen05 self-test timed out (NaN marker)"
    scale = np.abs(want).max()
    err = np.abs(got - want).max() / scale
    assert err < (2e-3 if mode == 0 else 2e-6), err
