"""tcgen05 / TMEM plumbing (descriptors, unswizzled K-major tiles, accumulation, commit, TMEM loads) against numpy."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tc_selftest_matches_numpy():
    from asr_b200._lib import lib, check
    rng = np.random.default_rng(5)
    a1 = rng.standard_normal((128, 32)).astype(np.float16)
    b1 = rng.standard_normal((32, 32)).astype(np.float16)
    a2 = (rng.standard_normal((128, 32)) * 2.0 ** -9).astype(np.float16)
    b2 = rng.standard_normal((32, 32)).astype(np.float16)
    dev = [torch.from_numpy(x).cuda() for x in (a1, b1, a2, b2)]
    d = torch.zeros((2, 128, 32), dtype=torch.float32, device="cuda")
    check(lib.asr_tc_selftest(*[t.data_ptr() for t in dev], d.data_ptr(), None), "asr_tc_selftest")
    torch.cuda.synchronize()
    want = a1.astype(np.float64) @ b1.astype(np.float64).T + a2.astype(np.float64) @ b2.astype(np.float64).T
    got = d.cpu().numpy()
    assert np.abs(got[0] - want).max() < 1e-4, np.abs(got[0] - want).max()      # float32 accumulation of exact products
    assert np.array_equal(got[0], got[1])
