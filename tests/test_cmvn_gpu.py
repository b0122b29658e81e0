"""standardize_dataset (VDR/attacks.py:48-69) on the GPU vs sklearn's StandardScaler."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_golden_cmvn():
    from asr_b200.voice_digit import attacks
    g = np.load(os.path.join(GOLD, "cmvn.npz"))
    sa, sb, sc = attacks.standardize_dataset(g["a"], g["b"], g["c"])
    for got, ref in ((sa, g["sa"]), (sb, g["sb"]), (sc, g["sc"])):
        assert got.dtype == np.float64 and got.shape == ref.shape
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)
    assert (sa[:, 5] == 0).all()          # constant column: scale_ -> 1, value - mean = 0


@pytest.mark.parametrize("shape", [(23665, 880), (1000, 2020), (3, 5), (1, 7)])
def test_stats_vs_sklearn(shape):
    import asr_b200 as A
    from oracle import cmvn_ref as cr
    rng = np.random.default_rng(shape[0])
    X = rng.standard_normal(shape) * rng.uniform(0.1, 50, shape[1]) + rng.uniform(-100, 100, shape[1])
    X[:, 0] = 3.25                         # constant feature
    mean, var, scale = cr.column_stats(X)
    xd = torch.from_numpy(X).cuda()
    cut = shape[0] // 3
    st = A.Standardizer(shape[1]).fit([xd[:cut], xd[cut:]])          # two row blocks, like train/dev/test
    np.testing.assert_allclose(st.mean.cpu().numpy(), mean, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(st.var.cpu().numpy(), var, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(st.scale.cpu().numpy(), scale, rtol=1e-9, atol=1e-12)
    from sklearn.preprocessing import StandardScaler
    ref = StandardScaler().fit_transform(X)
    got = st.transform(xd).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9)
    # float32 features (the kernel's native output) give the same statistics as their float64 widening
    x32 = xd.to(torch.float32)
    st32 = A.Standardizer(shape[1]).fit([x32])
    m64, _, _ = cr.column_stats(x32.cpu().numpy().astype(np.float64))
    np.testing.assert_allclose(st32.mean.cpu().numpy(), m64, rtol=1e-12, atol=1e-12)


def test_idempotence_full_size():
    """Standardising standardised data is (numerically) the identity: mean 0, variance 1."""
    import asr_b200 as A
    X = torch.randn(23665, 880, dtype=torch.float64, device="cuda") * 7 + 3
    st = A.Standardizer(880).fit([X])
    Y = st.transform(X)
    st2 = A.Standardizer(880).fit([Y])
    assert st2.mean.abs().max().item() < 1e-12
    assert (st2.var - 1).abs().max().item() < 1e-10
