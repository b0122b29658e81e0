"""Bit-exact parity of the noise path with the reference's numpy code (VDR/attacks.py:73-86,145-183,222-245)."""
import os

import numpy as np
import pytest
import torch

from synth import synth_clips, to_f32

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def test_clip_power_bit_exact_many_lengths():
    import asr_b200 as A
    lengths = [1, 2, 7, 8, 9, 15, 16, 100, 127, 128, 129, 255, 256, 257, 1000, 4097, 16000, 16384, 16385, 22050,
               40001, 100003, 220500]
    clips = to_f32(synth_clips(len(lengths), 0, 16000, 11, lengths=lengths))
    got = A.clip_power(A.ClipBatch.from_arrays(clips)).cpu().numpy()
    ref = np.array([np.mean(c ** 2) for c in clips], dtype=np.float32)
    assert got.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (got, ref)
    # int16 input decodes to the same float32 samples
    got16 = A.clip_power(A.ClipBatch.from_arrays(synth_clips(len(lengths), 0, 16000, 11, lengths=lengths))).cpu().numpy()
    assert np.array_equal(got16.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("L", [1, 7, 8, 9, 100, 128, 129, 1000, 4097, 16000, 22050, 70001, 90000])
def test_clip_power_bit_exact_equal_length_batches(L):
    """Batches of one length take the table-replay path of the power kernel (90 000 samples exceed the table and
    walk the tree); both must reproduce numpy's pairwise order bit for bit."""
    import asr_b200 as A
    clips16 = synth_clips(19, L, 16000, 12 + L)
    ref = np.array([np.mean(c ** 2) for c in to_f32(clips16)], dtype=np.float32)
    for clips in (clips16, to_f32(clips16)):
        got = A.clip_power(A.ClipBatch.from_arrays(clips)).cpu().numpy()
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_clip_power_narrow_cta_shape_bit_exact():
    """ASR_B200_POW_WARPS=4 (three 4-warp CTAs per SM instead of one of 12: the shape whose CTAs fit beside a step's tail
    kernels) is read when the library first launches the power pass, so it is checked in a fresh process: same bits."""
    import subprocess, sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import asr_b200 as A\nfrom synth import synth_clips, to_f32\n"
        "for L in (7, 129, 1000, 16000, 22050, 90000):\n"
        "    c16 = synth_clips(301, L, 16000, 30 + L)\n"
        "    ref = np.array([np.mean(c ** 2) for c in to_f32(c16)], dtype=np.float32)\n"
        "    for clips in (c16, to_f32(c16)):\n"
        "        got = A.clip_power(A.ClipBatch.from_arrays(clips)).cpu().numpy()\n"
        "        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), L\n"
        "lengths = [1, 2, 7, 8, 9, 100, 127, 128, 129, 257, 1000, 4097, 16000, 16385, 40001]\n"
        "c16 = synth_clips(len(lengths), 0, 16000, 21, lengths=lengths)\n"
        "ref = np.array([np.mean(c ** 2) for c in to_f32(c16)], dtype=np.float32)\n"
        "got = A.clip_power(A.ClipBatch.from_arrays(c16)).cpu().numpy()\n"
        "assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))\nprint('narrow ok')\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    for shape in ("4", "12"):
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, ASR_B200_POW_WARPS=shape), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "narrow ok" in r.stdout, (shape, r.stdout[-1500:], r.stderr[-1500:])


def test_copy_mapped_both_directions():
    """asr_copy_mapped: the small power read-back / sigma upload as a kernel over mapped pinned memory (no DMA-engine queueing)."""
    import asr_b200 as A
    from asr_b200.frontend import copy_mapped
    for dtype in (torch.float32, torch.float64):
        for n in (1, 3, 257, 8192):
            src = torch.randn(n, dtype=dtype)
            host_in = src.clone().pin_memory()
            dev = torch.zeros(n, dtype=dtype, device="cuda")
            copy_mapped(dev, host_in)
            host_out = torch.zeros(n, dtype=dtype).pin_memory()
            copy_mapped(host_out, dev)
            torch.cuda.current_stream().synchronize()
            assert torch.equal(dev.cpu(), src) and torch.equal(host_out, src)
    with pytest.raises(ValueError):
        copy_mapped(torch.zeros(4, device="cuda"), torch.zeros(4))          # pageable host memory is not device-accessible
    with pytest.raises(ValueError):
        copy_mapped(torch.zeros(4, device="cuda"), torch.zeros(5).pin_memory())


def test_snr_sigma_device_matches_host_chain():
    import asr_b200 as A
    from oracle import noise_ref as nr
    clips = to_f32(synth_clips(256, 16000, 16000, 12))
    batch = A.ClipBatch.from_arrays(clips)
    P = A.clip_power(batch)
    for snr in (0, 5, 10, 20, 60, 0.5):
        host = A.snr_sigma_host(P.cpu().numpy(), snr)
        oracle = np.array([float(nr.snr_sigma(c, snr)) for c in clips])
        assert np.array_equal(_bits(host), _bits(oracle))            # the safe contract is exact
        dev = A.snr_sigma_device(P, snr).cpu().numpy()
        mism = int((_bits(dev) != _bits(host)).sum())
        assert np.allclose(dev, host, rtol=1e-5, atol=0)             # glibc log10f/powf are not correctly rounded: a few float32 ulps
        print("snr", snr, "device-vs-host sigma mismatches:", mism, "of", len(clips))


@pytest.mark.parametrize("snr", [0, 5, 10, 20])
def test_golden_snr_mix_bit_exact(snr):
    from asr_b200.voice_digit import attacks
    g = np.load(os.path.join(GOLD, "noise.npz"))
    np.random.seed(1234 + snr)
    got = np.stack([attacks.add_white_noise_with_snr(c, snr) for c in g["audio_f32"]])
    assert got.dtype == np.float64
    assert np.array_equal(_bits(got), _bits(g[f"snr{snr}"]))


def test_golden_white_and_mixture_bit_exact():
    from asr_b200.voice_digit import attacks
    g = np.load(os.path.join(GOLD, "noise.npz"))
    np.random.seed(77)
    got = np.stack([attacks.add_white_noise(c, 0.01) for c in g["audio_f32"]])
    assert np.array_equal(_bits(got), _bits(g["white"]))
    np.random.seed(78)
    got = np.stack([attacks.add_noise(c, 0.01, 0.002) for c in g["audio_f32"]])
    assert np.array_equal(_bits(got), _bits(g["mixture"]))
    import asr_b200 as A
    P = A.clip_power(A.ClipBatch.from_arrays(list(g["audio_f32"]))).cpu().numpy()
    assert np.array_equal(P.view(np.uint32), g["power"].view(np.uint32))


def test_mixtgauss_and_float64_audio():
    from asr_b200.voice_digit import attacks
    from oracle import noise_ref as nr
    np.random.seed(5); got = attacks.mixtgauss(5000, 0.01, 0.003, 0.03)
    np.random.seed(5); ref = nr.mixtgauss(5000, 0.01, 0.003, 0.03)
    assert np.array_equal(_bits(got), _bits(ref))
    x = to_f32(synth_clips(1, 3001, 16000, 6))[0].astype(np.float64)
    np.random.seed(6); got = attacks.add_white_noise_with_snr(x, 10)
    np.random.seed(6); ref = nr.add_white_noise_with_snr(x, 10)
    assert np.array_equal(_bits(got), _bits(ref))


def test_feature_domain_noise_on_dataset():
    from asr_b200.voice_digit import attacks
    from oracle import noise_ref as nr
    X = np.random.default_rng(1).standard_normal((50, 880)) * 30
    np.random.seed(9); got = attacks.add_white_noise_on_dataset(X, 0.5)
    np.random.seed(9); ref = np.array(X)
    for i in range(ref.shape[0]):
        ref[i] = nr.add_white_noise(ref[i], 0.5)
    assert np.array_equal(_bits(got), _bits(ref))
    np.random.seed(10); got = attacks.add_noise_mixture_on_dataset(X, 0.01, 0.2)
    np.random.seed(10); ref = np.array(X)
    for i in range(ref.shape[0]):
        ref[i] = nr.add_noise(ref[i], 0.01, 0.2)
    assert np.array_equal(_bits(got), _bits(ref))


def test_batched_mix_full_size_roundtrip_properties():
    """BASELINE-size batch: x + 0*z == float64(x) exactly; mixing is per-clip independent; device randn is
    shard-invariant (value depends only on (seed, index))."""
    import asr_b200 as A
    clips = synth_clips(1024, 16000, 16000, 21)
    batch = A.ClipBatch.from_arrays(clips)
    n = batch.audio.shape[0]
    z = A.randn(42, 0, n)
    zero = torch.zeros(1024, dtype=torch.float64, device="cuda")
    out = A.mix_white(batch, z, zero)
    ref = batch.audio.to(torch.float64) / 32768.0
    assert torch.equal(out, ref)
    z2 = torch.cat([A.randn(42, 0, 1000), A.randn(42, 1000, n - 1000)])
    assert torch.equal(z, z2)
    zs = z[: 16000 * 64].cpu().numpy()
    assert abs(zs.mean()) < 5e-3 and abs(zs.std() - 1) < 5e-3 and np.abs(zs).max() < 7
    sig = torch.rand(1024, dtype=torch.float64, device="cuda")
    full = A.mix_white(batch, z, sig)
    host_x = np.stack(clips).astype(np.float32) / np.float32(32768.0)
    for i in (0, 517, 1023):
        o = int(batch.offsets_host[i])
        ref_i = host_x[i].astype(np.float64) + float(sig[i]) * z[o:o + 16000].cpu().numpy()
        assert np.array_equal(_bits(full[o:o + 16000].cpu().numpy()), _bits(ref_i))


def test_fused_noise_mfcc_matches_oracle_pipeline():
    """test_dataset_to_add_noise path: SNR mix fused into the MFCC launch vs oracle mix -> oracle MFCC."""
    from asr_b200.voice_digit import attacks
    from oracle import pipeline_ref as pr, librosa_ref as lr
    import asr_b200 as A
    waves = to_f32(synth_clips(6, 0, 22050, 33, lengths=[22050, 20000, 22050, 9000, 22050, 12345]))
    for snr in (0, 10, 20):
        np.random.seed(1234 + snr)
        got = attacks.black_box_attack_on_waveforms_snr(waves, snr)
        np.random.seed(1234 + snr)
        zs = [np.random.standard_normal(len(w)) for w in waves]
        ref = pr.black_box_attack_on_audio_dataset_snr(waves, snr, zs)
        assert got.shape == ref.shape == (6, 880) and got.dtype == np.float64
        err = np.abs(got - ref).max()
        assert err <= 3e-3, err
    np.random.seed(3); got = attacks.black_box_attack_on_waveforms(waves, sigma=0.02)
    np.random.seed(3); zs = [np.random.standard_normal(len(w)) for w in waves]
    ref = np.stack([pr.black_box_attack_on_audio(w, 44, sigma=0.02, z=z).flatten() for w, z in zip(waves, zs)])
    assert np.abs(got - ref).max() <= 3e-3
    np.random.seed(4); got = attacks.black_box_attack_on_waveforms(waves, p=0.01, alpha=0.004)
    np.random.seed(4)
    ref = []
    for w in waves:
        q = np.random.standard_normal(len(w)); g = np.random.standard_normal(len(w))
        ref.append(pr.black_box_attack_on_audio(w, 44, p_peak=0.01, alpha=0.004, q=q, g=g).flatten())
    assert np.abs(got - np.stack(ref)).max() <= 3e-3
    # C2 shape: C1 front end + SNR noise, fused, int16 input
    clips = synth_clips(8, 16000, 16000, 44)
    for snr in (0, 5, 10, 20):
        np.random.seed(1234 + snr)
        got = attacks.black_box_attack_on_waveforms_snr(clips, snr, utterance_length=101, params=A.C1)
        np.random.seed(1234 + snr)
        zs = [np.random.standard_normal(16000) for _ in clips]
        ref = pr.black_box_attack_on_audio_dataset_snr(to_f32(clips), snr, zs, utterance_length=101, p=lr.C1)
        assert np.abs(got - ref).max() <= 3e-3


def test_babble_stream_and_mix():
    """BASELINE configs[1] babble (no reference implementation; recipe of SURVEY.md 8(d)): the stream equals the oracle's
    sum bit for bit, its power to 1e-13, and the mix - the white-noise formula on that stream - is bit-exact given the gain."""
    import asr_b200 as A
    from oracle import noise_ref as nr
    rng = np.random.default_rng(9)
    lengths = [16000] * 20 + rng.integers(3000, 17000, size=90).tolist()
    clips = to_f32(synth_clips(len(lengths), 17000, 16000, 31))
    clips = [c[:n] for c, n in zip(clips, lengths)]
    batch = A.ClipBatch.from_arrays(clips)
    b_dev, pb_dev = A.babble_stream(batch)
    bs = batch.unpack(b_dev)
    pb = pb_dev.cpu().numpy()
    for i in range(len(clips)):
        want = nr.babble_stream(clips, i)
        assert np.array_equal(bs[i].view(np.uint64), want.view(np.uint64)), i
        assert abs(pb[i] - np.mean(want ** 2)) <= 1e-13 * np.mean(want ** 2)
    sigma = A.snr_sigma_host(A.clip_power(batch).cpu().numpy(), 5)
    gain = A.babble_gain_host(sigma, pb)
    noisy = batch.unpack(A.mix_white(batch, b_dev, torch.from_numpy(gain).cuda()))
    for i in range(len(clips)):
        want = nr.add_babble_with_snr(clips, i, 5, gain=gain[i])
        assert np.array_equal(noisy[i].view(np.uint64), want.view(np.uint64)), i
        free = nr.add_babble_with_snr(clips, i, 5)                     # the oracle's own gain: same signal to 1e-12
        assert np.abs(free - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
        # realised SNR is the target (the point of the sigma law)
        snr = 10 * np.log10(np.mean(clips[i].astype(np.float64) ** 2) / np.mean((want - clips[i]) ** 2))
        assert abs(snr - 5) < 1e-3
