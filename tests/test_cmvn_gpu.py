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


@pytest.mark.parametrize("kind", ["white", "mixture"])
def test_noisy_test_rows_fused_into_standardisation(kind):
    """VDR/attacks.py:433-491: MFCC-domain noise on the test rows, then standardize_dataset over train + dev + noisy test.
    The fused form (noise mixed inside the statistics and apply kernels) equals the reference's two steps."""
    from asr_b200.voice_digit import attacks
    from oracle import noise_ref as nr, cmvn_ref as cr
    rng = np.random.default_rng(4)
    train, val, test = (rng.standard_normal((n, 880)) * 5 + 2 for n in (700, 200, 100))
    if kind == "white":
        np.random.seed(77)
        noisy = np.stack([nr.add_white_noise(r, 0.5) for r in test])
        np.random.seed(77)
        got = attacks.standardize_dataset_with_noisy_test(train, val, test, sigma=0.5)
    else:
        np.random.seed(78)
        noisy = np.stack([nr.add_noise(r, 0.01, 0.3) for r in test])
        np.random.seed(78)
        got = attacks.standardize_dataset_with_noisy_test(train, val, test, p=0.01, alpha=0.3)
    want = cr.standardize_dataset(train, val, noisy)
    for g, w in zip(got, want):
        np.testing.assert_allclose(g, w, rtol=1e-9, atol=1e-9)


def test_fused_single_gpu_form_equals_block_form():
    """fit_transform (3 launches, statistics finished inside the apply launch) == fit + transform; also the merge of
    several ranks' messages (emulated on one GPU: shards fitted separately, messages stacked) == the single-rank fit."""
    import asr_b200 as A
    from asr_b200._lib import lib, check
    X = torch.randn(5000, 1313, dtype=torch.float32, device="cuda") * 3 + 1
    a = A.Standardizer(1313)
    ya = a.fit_transform(X, out_dtype=torch.float32)
    b = A.Standardizer(1313).fit([X])
    yb = b.transform(X, out_dtype=torch.float32)
    assert torch.equal(a.mean, b.mean) and torch.equal(a.var, b.var) and torch.equal(a.scale, b.scale) and torch.equal(ya, yb)
    assert b.n_total == 5000 and a.n_total == 5000
    msgs = []
    for lo, hi in ((0, 1250), (1250, 2500), (2500, 3750), (3750, 5000)):
        s = A.Standardizer(1313)
        s.local_stats([X[lo:hi]]); s.local_message()
        msgs.append(s.msg.clone())
    m = A.Standardizer(1313)
    m.msgs = torch.stack(msgs)
    m.merge()
    torch.cuda.synchronize()
    assert m.n_total == 5000
    np.testing.assert_allclose(m.mean.cpu().numpy(), b.mean.cpu().numpy(), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(m.var.cpu().numpy(), b.var.cpu().numpy(), rtol=1e-11, atol=1e-13)
    ym = m.transform(X, out_dtype=torch.float64)
    np.testing.assert_allclose(ym.cpu().numpy(), b.transform(X).cpu().numpy(), rtol=1e-9, atol=1e-9)
