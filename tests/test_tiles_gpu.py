"""The TMA-staged "tiles" path of asr_mfcc_batch (n_fft = 512): path selection and cases that stress its block
structure (clip boundaries inside 32-frame blocks, ranges of many CTAs, element tails of the bulk copies)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
from synth import synth_clips, to_f32

pytestmark = pytest.mark.gpu


def test_path_selection():
    import asr_b200 as A
    plan = A.MfccPlan(A.C1)
    for dt in (np.int16, np.float32, np.float64):
        assert plan.path_used(dt, noisy=False) == "tiles"
    assert plan.path_used(np.int16, noisy=True) == "tiles"
    # wider raw samples need more shared memory for the staged block: float64 audio + float64 noise does not fit
    assert plan.path_used(np.float32, noisy=True) in ("tiles", "frames")
    assert plan.path_used(np.float64, noisy=True) == "frames"
    assert plan.launches(True) == 3 and plan.launches(False) == 3
    assert A.MfccPlan(A.C1, path="clip").path_used(np.int16, True) == "clip"
    assert A.MfccPlan(A.C1, path="frames").path_used(np.int16, False) == "frames"
    assert A.MfccPlan(A.C3).path_used(np.int16, False) == "tiles"
    # pre-emphasis and other FFT sizes are outside the tiles path
    assert A.MfccPlan(A.C1.replace(preemph=0.97)).path_used(np.int16, True) == "frames"
    assert A.MfccPlan(A.C1.replace(preemph=0.97)).path_used(np.int16, False) == "clip"
    assert A.MfccPlan(A.C5).path_used(np.int16, True) == "clip"
    assert A.MfccPlan(A.REF_VDR).path_used(np.float32, False) == "clip"


@pytest.mark.parametrize("noisy", [False, True])
def test_many_ctas_ragged_against_oracle_and_clip_kernel(noisy):
    """3000 clips of 0.2-1.3 s (lengths not multiples of 8): every CTA gets a range, blocks hold pieces of 2-3 clips,
    the bulk copies end in element tails.  Checked against the per-clip kernel everywhere and the oracle on a sample."""
    import asr_b200 as A
    from oracle import librosa_ref as lr, noise_ref as nr
    rng = np.random.default_rng(11)
    lengths = rng.integers(3200, 20800, size=3000).tolist()
    base = synth_clips(64, 20800, 16000, 12)
    clips = [np.roll(base[i % 64], 17 * i)[:n] for i, n in enumerate(lengths)]
    batch = A.ClipBatch.from_arrays(clips)
    noise = None
    if noisy:
        z = A.randn(5, 0, batch.audio.shape[0])
        sig = A.snr_sigma_device(A.clip_power(batch), 5.0)
        noise = A.Noise.white(z, sig)
    out_t, st_t = A.MfccPlan(A.C1, path="tiles").mfcc(batch, noise=noise)
    out_c, st_c = A.MfccPlan(A.C1, path="clip").mfcc(batch, noise=noise)
    torch.cuda.synchronize()
    assert int(st_t.max()) == 0 and torch.equal(st_t, st_c)
    assert torch.isfinite(out_t).all()
    # two float32 implementations of the same arithmetic (different FFT rounding, fast log2 in both)
    assert float((out_t - out_c).abs().max()) < 3e-3
    zs = batch.unpack(noise.z) if noisy else None
    sg = noise.sigma.cpu().numpy() if noisy else None
    for i in (0, 1, 517, 1499, 2998, 2999):
        x = to_f32([clips[i]])[0]
        if noisy:
            x = nr.add_white_noise_z(x, sg[i], zs[i])
        r = lr.mfcc(np.asarray(x), lr.C1)
        err = np.abs(out_t[i, :, :r.shape[1]].cpu().numpy() - r).max()
        assert err < 3e-3, (i, err)
        assert (out_t[i, :, r.shape[1]:] == 0).all()


def test_rows_do_not_depend_on_batch_composition():
    """A clip's features are the same alone and inside a large batch (flat frame list, block and CTA boundaries)."""
    import asr_b200 as A
    clips = synth_clips(700, 16000, 16000, 21)
    plan = A.MfccPlan(A.C1, path="tiles")
    full, _ = plan.mfcc(A.ClipBatch.from_arrays(clips))
    for i in (0, 3, 350, 699):
        one, _ = plan.mfcc(A.ClipBatch.from_arrays([clips[i]]))
        assert torch.equal(one[0], full[i])


@pytest.mark.parametrize("dtype", [np.int16, np.float32])
def test_fused_mix_staged_samples_are_bit_exact(dtype):
    """The samples the TILES kernel stages with noise fused - what its FFT sees - equal float32(reference signal) bit
    for bit: float32(add_white_noise_with_snr(x, snr)) of VDR/attacks.py:222-245 given the same z, with sigma from the
    reference's own chain.  Ragged lengths: clip edges, element tails, blocks holding pieces of several clips."""
    import asr_b200 as A
    from oracle import noise_ref as nr
    rng = np.random.default_rng(3)
    lengths = [16000, 16000, 4001, 9999, 12345, 16000, 257, 15992] + rng.integers(3000, 17000, size=56).tolist()
    base = synth_clips(len(lengths), 17000, 16000, 77)
    clips = [b[:n] for b, n in zip(base, lengths)]
    if dtype == np.float32:
        clips = to_f32(clips)
    batch = A.ClipBatch.from_arrays(clips)
    plan = A.MfccPlan(A.C1, path="tiles")
    assert plan.path_used(dtype, noisy=True) == "tiles"
    xs = clips if dtype == np.float32 else to_f32(clips)
    for snr in (0, 10, 20):
        z = A.randn(40 + snr, 0, batch.audio.shape[0])
        sigma = A.snr_sigma_host(A.clip_power(batch).cpu().numpy(), snr)
        staged = torch.full((batch.audio.shape[0],), float("nan"), dtype=torch.float32, device="cuda")
        plan.set_stage_probe(staged)
        out, status = plan.mfcc(batch, noise=A.Noise.white(z, torch.from_numpy(sigma).cuda()))
        torch.cuda.synchronize()
        plan.set_stage_probe(None)
        assert int(status.max()) == 0
        got = batch.unpack(staged)
        zs = batch.unpack(z)
        bad = 0
        for i, (x, zz) in enumerate(zip(xs, zs)):
            want = nr.add_white_noise_with_snr_z(x, snr, zz).astype(np.float32)   # float64 reference signal, rounded once
            bad += int((got[i].view(np.uint32) != want.view(np.uint32)).sum())
        assert bad == 0, (snr, bad)
