"""tcgen05 / TMEM plumbing (descriptors, unswizzled K-major tiles, accumulation, commit, TMEM loads) against numpy."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from synth import synth_clips, to_f32

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


def test_tc_path_selection():
    import asr_b200 as A
    plan = A.MfccPlan(A.C1, path="tc")
    assert plan.path_used(np.int16, noisy=False) == "tc" and plan.path_used(np.int16, noisy=True) == "tc"
    assert plan.path_used(np.float32, noisy=False) == "tiles"           # float audio keeps the FP32 tile kernel
    assert A.MfccPlan(A.C3, path="tc").path_used(np.int16, False) == "tc"
    assert A.MfccPlan(A.C5, path="tc").path_used(np.int16, False) == "clip"    # n_fft = 1024
    assert plan.launches(True) == 3


@pytest.mark.parametrize("preset", ["c1", "c3"])
@pytest.mark.parametrize("noisy", [False, True])
def test_tc_small_batches_against_oracle(preset, noisy):
    """A handful of clips (fewer frames than one 128-frame tile, then a few tiles): every value against the oracle."""
    import asr_b200 as A
    from oracle import librosa_ref as lr, noise_ref as nr
    P = A.PRESETS[preset]
    for n, L in ((1, 16000), (3, 16000), (9, 12345)):
        clips = synth_clips(n, L, 16000, 90 + n)
        batch = A.ClipBatch.from_arrays(clips)
        noise = None
        if noisy:
            z = A.randn(17, 0, batch.audio.shape[0])
            sig = torch.from_numpy(A.snr_sigma_host(A.clip_power(batch).cpu().numpy(), 10)).cuda()
            noise = A.Noise.white(z, sig)
        out, st = A.MfccPlan(P, path="tc").mfcc(batch, noise=noise)
        torch.cuda.synchronize()
        assert int(st.max()) == 0 and torch.isfinite(out).all()
        zs = batch.unpack(noise.z) if noisy else None
        for i in range(n):
            x = to_f32([clips[i]])[0]
            if noisy:
                x = nr.add_white_noise_with_snr_z(x, 10, zs[i])
            r = lr.mfcc(np.asarray(x), lr.PRESETS[preset])
            err = np.abs(out[i, :, :r.shape[1]].cpu().numpy() - r).max()
            assert err < 3e-3, (preset, noisy, n, i, err)


@pytest.mark.parametrize("noisy", [False, True])
def test_tc_many_ctas_ragged_against_tiles_and_oracle(noisy):
    """3000 ragged clips: every CTA gets a range, tiles hold pieces of several clips, quiet and loud clips mixed."""
    import asr_b200 as A
    from oracle import librosa_ref as lr, noise_ref as nr
    rng = np.random.default_rng(11)
    lengths = rng.integers(3200, 20800, size=3000).tolist()
    base = synth_clips(64, 20800, 16000, 12)
    clips = []
    for i, n in enumerate(lengths):
        c = np.roll(base[i % 64], 17 * i)[:n]
        if i % 5 == 0:
            c = (c.astype(np.int32) >> (i % 11)).astype(np.int16)        # quiet clips: down to a few LSB of amplitude
        clips.append(c)
    batch = A.ClipBatch.from_arrays(clips)
    noise = None
    if noisy:
        z = A.randn(5, 0, batch.audio.shape[0])
        sig = torch.from_numpy(A.snr_sigma_host(A.clip_power(batch).cpu().numpy(), 5)).cuda()
        noise = A.Noise.white(z, sig)
    out_t, st_t = A.MfccPlan(A.C1, path="tc").mfcc(batch, noise=noise)
    out_c, st_c = A.MfccPlan(A.C1, path="tiles").mfcc(batch, noise=noise)
    torch.cuda.synchronize()
    assert int(st_t.max()) == 0 and torch.equal(st_t, st_c)
    assert torch.isfinite(out_t).all()
    assert float((out_t - out_c).abs().max()) < 3e-3
    zs = batch.unpack(noise.z) if noisy else None
    for i in (0, 1, 5, 10, 55, 517, 1499, 2995, 2998, 2999):
        x = to_f32([clips[i]])[0]
        if noisy:
            x = nr.add_white_noise_with_snr_z(x, 5, zs[i])
        r = lr.mfcc(np.asarray(x), lr.C1)
        err = np.abs(out_t[i, :, :r.shape[1]].cpu().numpy() - r).max()
        assert err < 3e-3, (i, err)
        assert (out_t[i, :, r.shape[1]:] == 0).all()
