"""the classifier forward pass (VDR/train_constraints.py:63-88) and the accuracy-vs-SNR sweep
(VDR/attacks.py:401-422).  The trained .h5 weights are not in the reference tree: synthetic weights, parity against the
numpy restatement in oracle/mlp_ref.py."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
from oracle import mlp_ref as mr
from synth import synth_clips


def test_oracle_softmax_rows_and_architecture():
    layers = mr.random_weights(1)
    assert [ly["kernel"].shape for ly in layers] == [(880, 1024), (1024, 512), (512, 256), (256, 128), (128, 64), (64, 10)]
    assert all((ly["kernel"] >= 0).all() for ly in layers)            # kernel_constraint=NonNeg()
    assert ["gamma" in ly for ly in layers] == [True] * 5 + [False]
    x = np.random.default_rng(0).standard_normal((17, 880)).astype(np.float32)
    p = mr.predict(x, layers)
    assert p.shape == (17, 10) and p.dtype == np.float32
    np.testing.assert_allclose(p.sum(axis=1), 1.0, atol=1e-6)
    onehot = np.eye(10, dtype=np.float32)[np.argmax(p, axis=1)]
    assert mr.accuracy(p, onehot) == 1.0


def test_batchnorm_folding_is_exact_algebra():
    """(s*h + t) @ K + b == h @ (s[:,None]*K) + (t @ K + b): the fold the device class applies, checked in float64."""
    rng = np.random.default_rng(3)
    h = np.maximum(rng.standard_normal((5, 8)), 0)
    s, t = rng.uniform(0.5, 2, 8), rng.standard_normal(8)
    K, b = np.abs(rng.standard_normal((8, 4))), rng.standard_normal(4)
    np.testing.assert_allclose((s * h + t) @ K + b, h @ (s[:, None] * K) + (t @ K + b), rtol=1e-13, atol=1e-13)


@pytest.mark.gpu
def test_predict_matches_oracle():
    import torch
    import asr_b200 as A
    layers = mr.random_weights(7)
    x = (3.0 * np.random.default_rng(1).standard_normal((2048, 880))).astype(np.float32)   # standardised features
    model = A.DenseStack(layers)
    got = model.predict(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = mr.predict(x, layers)
    assert got.shape == ref.shape == (2048, 10)
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-5)          # float32 GEMMs, different summation order, folded BN
    margin = np.sort(ref, axis=1)
    sure = margin[:, -1] - margin[:, -2] > 1e-4                       # rows whose decision is not a numerical tie
    assert sure.mean() > 0.95 and (np.argmax(got, 1)[sure] == np.argmax(ref, 1)[sure]).all()
    # get_weights() order round trip
    flat = []
    for ly in layers:
        flat += [ly["kernel"], ly["bias"]] + ([ly["gamma"], ly["beta"], ly["moving_mean"], ly["moving_var"]] if "gamma" in ly else [])
    again = A.DenseStack.from_keras_weights(flat).predict(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(again, got)
    labels = torch.from_numpy(np.argmax(ref, axis=1))
    assert model.accuracy(torch.from_numpy(x).cuda(), labels) >= sure.mean()


@pytest.mark.gpu
def test_accuracy_vs_snr_sweep_runs_on_device():
    """The SNR sweep of attacks.py with the reference's parameters (REF_VDR rows of 880): per SNR, fused noisy MFCC ->
    standardise with the fitted statistics -> predict; checked against the same steps taken one by one."""
    import torch
    import asr_b200 as A
    clips = synth_clips(24, 22050, 22050, 5)
    batch = A.ClipBatch.from_arrays([c.astype(np.float32) / np.float32(32768.0) for c in clips])
    plan = A.MfccPlan(A.REF_VDR)
    clean, _ = plan.mfcc(batch, out_frames=44)
    rows = clean.reshape(24, -1)
    assert rows.shape[1] == 880
    std = A.Standardizer(880).fit([rows])
    model = A.DenseStack(mr.random_weights(9))
    labels = model.logits(std.transform(rows, out_dtype=torch.float32)).argmax(dim=1)      # the clean decisions
    snrs = [60, 30, 20, 15, 10, 5, 0]                                                       # VDR/attacks.py:319
    acc = A.accuracy_vs_snr([model], batch, labels, snrs, plan, std, seed=3, out_frames=44)
    assert list(acc) == snrs and all(0.0 <= a[0] <= 1.0 for a in acc.values())
    # one SNR by hand
    z = A.randn(3 + snrs.index(20), 0, batch.audio.shape[0])
    noise = A.Noise.white(z, A.snr_sigma_device(A.clip_power(batch), 20.0))
    f, _ = plan.mfcc(batch, out_frames=44, noise=noise)
    r = std.transform(f.reshape(24, -1), out_dtype=torch.float32)
    assert acc[20][0] == model.accuracy(r, labels)
    assert acc[60][0] >= acc[0][0]                       # 60 dB of SNR perturbs the decisions less than 0 dB


@pytest.mark.gpu
def test_mlp_forward_kernel_against_library_gemms_and_odd_shapes():
    """asr_mlp_forward (one fused launch of this repository's kernel) against the cuBLAS form of the same folded network:
    row counts that are not multiples of the 16-row tile, a strided input view, narrow and wide layers, 1 and 8 layers."""
    import torch
    import asr_b200 as A
    rng = np.random.default_rng(11)
    for sizes, n in (((880, 1024, 512, 256, 128, 64, 10), 1037), ((2020, 1024, 512, 256, 128, 64, 20), 50),
                     ((7, 3), 1), ((33, 1000, 5), 17), ((16, 32, 48, 64, 80, 96, 112, 128, 9), 129)):
        layers = mr.random_weights(3, sizes)             # (2020 inputs: the speaker network's rows)
        model = A.DenseStack(layers)
        wide = torch.from_numpy((2.0 * rng.standard_normal((n, sizes[0] + 5))).astype(np.float32)).cuda()
        x = wide[:, 2:2 + sizes[0]]                      # leading dimension > width, unaligned start
        lg = model.logits(x)
        ref = model.logits_library(x.contiguous())
        scale = float(ref.abs().max()) + 1e-6
        assert float((lg - ref).abs().max()) <= 2e-5 * scale + 1e-6
        pr = model.predict(x)
        assert torch.allclose(pr.sum(1), torch.ones(n, device=pr.device), atol=1e-5)
        assert torch.equal(model.decide(x).long(), lg.argmax(1))
        np.testing.assert_allclose(pr.cpu().numpy(), mr.predict(x.cpu().numpy(), layers), rtol=0, atol=3e-5)
