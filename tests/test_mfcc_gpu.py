"""GPU parity of the fused MFCC kernel against the CPU oracle (librosa 0.9 restatement).

Tolerance (stated per north_star): the kernel computes in fp32 (fp32 FFT, fast log2), the oracle's FFT
runs in float64; per coefficient we require
    |gpu - oracle| <= ATOL + RTOL * max|oracle_clip|
with ATOL = 2e-3 (dB-scale cepstra span several hundred units) and RTOL = 2e-5.  The direct-DFT path
(n_fft = 441, 128 mel filters on 221 bins, many filters with one weak bin) gets ATOL = 5e-3.
"""
import os

import numpy as np
import pytest
import torch

from synth import synth_clips, to_f32

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ATOL, RTOL = 2e-3, 2e-5
PATH = "auto"


@pytest.fixture(params=["clip", "frames", "tiles", "tc"], autouse=True)
def _kernel_path(request):
    """Every test runs on both kernel paths: one CTA per clip, and the block-pipelined n_fft = 512 path
    (presets with another n_fft only have the first; "frames" then falls back to it)."""
    global PATH
    PATH = request.param
    yield
    PATH = "auto"



def _o():
    from oracle import librosa_ref as lr
    return lr


def _close(got, ref, atol=ATOL, rtol=RTOL):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.abs(got - ref)
    tol = atol + rtol * np.abs(ref).reshape(ref.shape[0], -1).max(axis=1).reshape((-1,) + (1,) * (ref.ndim - 1))
    worst = float((err - tol).max())
    assert worst <= 0, f"max abs err {err.max():.3e} (tolerance exceeded by {worst:.3e})"
    return float(err.max())


def _run(preset, clips, out_frames=None, dtype=None, **kw):
    import asr_b200 as A
    plan = A.MfccPlan(A.PRESETS[preset].replace(**kw) if kw else A.PRESETS[preset], path=PATH)
    batch = A.ClipBatch.from_arrays(clips, dtype=dtype)
    out, status = plan.mfcc(batch, out_frames=out_frames)
    torch.cuda.synchronize()
    return out.cpu().numpy(), status.cpu().numpy()


@pytest.mark.parametrize("name", ["ref_vdr", "c1", "c3", "c5", "ref_sr"])
def test_golden(name):
    g = np.load(os.path.join(GOLD, f"mfcc_{name}.npz"))
    clips = list(g["audio_i16"])
    preset = name
    if name == "ref_sr":
        x = [c.astype(np.float64) / 32768.0 for c in clips]       # float64 windows, as the reference passes them
        got, st = _run(preset, x)
        err = _close(got, g["mfcc"], atol=5e-3)
    else:
        got, st = _run(preset, clips)                            # int16 in: value/32768 == the golden float32 input
        err = _close(got, g["mfcc"])
    assert (st == 0).all()
    print(name, "max abs err", err)


@pytest.mark.parametrize("name,dtype", [("c1", np.float32), ("c1", np.float64), ("ref_vdr", np.float32)])
def test_input_dtypes(name, dtype):
    lr = _o()
    p = lr.PRESETS[name]
    clips = synth_clips(5, p.sr, p.sr, 31)
    x = [c.astype(dtype) / dtype(32768.0) for c in clips]
    got, st = _run(name, x)
    ref = np.stack([lr.mfcc(v.astype(np.float32), p) for v in x])
    _close(got, ref)
    assert (st == 0).all()


def test_ragged_lengths_and_feature_padding():
    """Variable-length clips in one batch; frames >= T are zeros, frames >= out_frames are cut
    (VDR/extract_features_construct_dataset.py:33-37)."""
    from oracle import pipeline_ref as pr
    lr = _o()
    p = lr.REF_VDR
    lengths = [22050, 1025, 15000, 30000, 2047, 2048, 2049, 22049]
    clips = to_f32(synth_clips(len(lengths), 0, p.sr, 5, lengths=lengths))
    got, st = _run("ref_vdr", clips, out_frames=44)
    assert (st == 0).all()
    ref = np.stack([pr.extract_features(c, 44, p) for c in clips])
    _close(got, ref)
    for i, L in enumerate(lengths):
        T = 1 + L // 512
        if T < 44:
            assert (got[i][:, T:] == 0).all()


def test_too_short_clip_is_flagged_not_fatal():
    """np.pad(reflect) raises for len <= n_fft//2; the batched kernel flags the clip and zero-fills."""
    p = _o().C1
    clips = to_f32(synth_clips(3, 0, p.sr, 9, lengths=[16000, 256, 257]))
    got, st = _run("c1", clips, out_frames=101)
    assert st.tolist() == [0, 1, 0]
    assert (got[1] == 0).all()
    ref = _o().mfcc(clips[2], p)
    _close(got[2:3, :, :ref.shape[1]], ref[None])


def test_delta_needs_nine_frames():
    p = _o().C3
    clips = to_f32(synth_clips(2, 0, p.sr, 10, lengths=[160 * 7 + 10, 160 * 9]))
    got, st = _run("c3", clips, out_frames=12)
    assert st.tolist() == [2, 0]


@pytest.mark.parametrize("kw", [
    dict(top_db=-1.0), dict(lifter=0.0), dict(preemph=0.97), dict(pad_mode="constant"), dict(center=False),
    dict(window="hann", win_length=512), dict(n_mels=40, n_mfcc=40), dict(fmin=20.0, fmax=7600.0),
    dict(hop_length=161), dict(delta_orders=2, delta_width=5),
])
def test_keyword_coverage_c1(kw):
    lr = _o()
    p = lr.C1.replace(**kw)
    clips = to_f32(synth_clips(3, 16000, p.sr, 77))
    got, st = _run("c1", clips, **kw)
    assert (st == 0).all()
    ref = np.stack([lr.mfcc(c, p) for c in clips])
    _close(got, ref)


@pytest.mark.parametrize("n_fft,hop", [(1024, 256), (2048, 512), (441, 220), (400, 160), (600, 200)])
def test_fft_sizes(n_fft, hop):
    lr = _o()
    p = lr.REF_VDR.replace(n_fft=n_fft, win_length=n_fft, hop_length=hop, n_mels=64)
    clips = to_f32(synth_clips(2, 12000, p.sr, 3))
    got, st = _run("ref_vdr", clips, n_fft=n_fft, win_length=n_fft, hop_length=hop, n_mels=64)
    ref = np.stack([lr.mfcc(c, p) for c in clips])
    _close(got, ref, atol=5e-3)


def test_rfftfreq_switch_matters_for_odd_nfft():
    lr = _o()
    clips = [c.astype(np.float64) for c in to_f32(synth_clips(1, 22050, 22050, 4))]
    p = lr.REF_SR.replace(fftfreq_mode="rfftfreq")
    got, _ = _run("ref_sr", clips, fftfreq_mode="rfftfreq")
    _close(got, np.stack([lr.mfcc(c, p) for c in clips]), atol=5e-3)
    other = np.stack([lr.mfcc(c, lr.REF_SR) for c in clips])
    assert np.abs(got - other).max() > 0.1


def test_logmel_stage():
    lr = _o()
    import asr_b200 as A
    clips = to_f32(synth_clips(2, 16000, 16000, 8))
    plan = A.MfccPlan(A.C1, path=PATH)
    out, st = plan.logmel(A.ClipBatch.from_arrays(clips))
    ref = np.stack([lr.log_mel(c, lr.C1) for c in clips])
    _close(out.cpu().numpy(), ref, atol=1e-3, rtol=0)


def test_plan_tables_match_oracle():
    lr = _o()
    import asr_b200 as A
    import scipy.fftpack
    for name in ("ref_vdr", "ref_sr", "c1", "c3", "c5"):
        p = lr.PRESETS[name]
        t = A.MfccPlan(A.PRESETS[name]).tables()
        np.testing.assert_allclose(t["window"], lr.fft_window(p).astype(np.float32), atol=1e-7)
        np.testing.assert_allclose(t["mel"], lr.mel_filterbank(p), rtol=2e-6, atol=1e-9)
        D = scipy.fftpack.dct(np.eye(p.n_mels), axis=0, type=2, norm="ortho")[:p.n_mfcc]
        if p.lifter > 0:
            D = D * (1 + (p.lifter / 2) * np.sin(np.pi * np.arange(1, 1 + p.n_mfcc) / p.lifter))[:, None]
        np.testing.assert_allclose(t["dct"], D, atol=2e-6)
        if p.delta_orders:
            import scipy.signal
            for o in range(1, p.delta_orders + 1):
                c = scipy.signal.savgol_coeffs(p.delta_width, o, deriv=o, use="dot")
                np.testing.assert_allclose(t["delta_taps"][o - 1], c, atol=1e-7)


def test_large_batch_properties_c1():
    """BASELINE-size batch (1024 clips): order independence and determinism instead of a CPU oracle pass."""
    import asr_b200 as A
    clips = synth_clips(1024, 16000, 16000, 123)
    plan = A.MfccPlan(A.C1, path=PATH)
    a, st = plan.mfcc(A.ClipBatch.from_arrays(clips))
    perm = np.random.default_rng(0).permutation(1024)
    b, _ = plan.mfcc(A.ClipBatch.from_arrays([clips[i] for i in perm]))
    c, _ = plan.mfcc(A.ClipBatch.from_arrays(clips))
    assert (st == 0).all()
    assert torch.equal(a, c)                       # bit-deterministic
    assert torch.equal(a[perm], b)                 # a clip's features do not depend on its batch slot
    assert torch.isfinite(a).all()
    lr = _o()
    for i in (0, 511, 1023):
        _close(a[i:i + 1].cpu().numpy(), lr.mfcc(to_f32([clips[i]])[0], lr.C1)[None])


@pytest.mark.parametrize("dtype", [np.int16, np.float32, np.float64])
def test_fused_noise_c1(dtype):
    """White (per-clip sigma) and Gaussian-mixture noise fused into the launch == oracle mix -> oracle MFCC, on both
    kernel paths and for every input dtype; ragged lengths so blocks of the frames path straddle clips."""
    import asr_b200 as A
    from oracle import noise_ref as nr
    lr = _o()
    lengths = [16000, 8000, 16000, 4321, 12000]
    clips = synth_clips(len(lengths), 0, 16000, 91, lengths=lengths)
    if dtype == np.int16:
        host = clips
        xs = to_f32(clips)
    else:
        xs = [c.astype(dtype) for c in to_f32(clips)]
        host = xs
    rng = np.random.default_rng(5)
    zs = [rng.standard_normal(n) for n in lengths]
    gs = [rng.standard_normal(n) for n in lengths]
    sig = np.array([0.01, 0.02, 0.005, 0.03, 0.0], dtype=np.float64)
    plan = A.MfccPlan(A.C1, path=PATH)
    batch = A.ClipBatch.from_arrays(host)
    zd = A.ClipBatch.from_arrays(zs).audio
    gd = A.ClipBatch.from_arrays(gs).audio
    out, st = plan.mfcc(batch, noise=A.Noise.white(zd, torch.from_numpy(sig).cuda()))
    assert int(st.max()) == 0
    ref = [lr.mfcc(nr.add_white_noise_z(x, s, z).astype(np.float64), lr.C1) for x, s, z in zip(xs, sig, zs)]
    for i, r in enumerate(ref):
        _close(out[i:i + 1, :, :r.shape[1]].cpu().numpy(), r[None], atol=3e-3)
        assert (out[i, :, r.shape[1]:] == 0).all()
    out, st = plan.mfcc(batch, noise=A.Noise.mixture(zd, gd, 0.01, 0.004))
    ref = [lr.mfcc(nr.add_noise_z(x, 0.01, 0.004, q, g).astype(np.float64), lr.C1) for x, q, g in zip(xs, zs, gs)]
    for i, r in enumerate(ref):
        _close(out[i:i + 1, :, :r.shape[1]].cpu().numpy(), r[None], atol=3e-3)


def test_many_tiny_and_empty_clips():
    """Clips of a few frames, clips that cannot be framed and long clips in one batch: the frames path packs up to
    four runs into a 32-frame block and must neither drop nor duplicate a frame."""
    import asr_b200 as A
    lr = _o()
    rng = np.random.default_rng(17)
    lengths = [300, 100, 16000, 257, 400, 0, 700, 320, 5000, 258, 100, 100, 100, 900, 16000, 480, 330, 260] * 3
    clips = synth_clips(len(lengths), 0, 16000, 92, lengths=[max(1, n) for n in lengths])
    clips = [c[:n] for c, n in zip(clips, lengths)]
    plan = A.MfccPlan(A.C1, path=PATH)
    out, st = plan.mfcc(A.ClipBatch.from_arrays(clips), out_frames=101)
    st = st.cpu().numpy()
    for i, (c, n) in enumerate(zip(clips, lengths)):
        if n <= 256:
            assert st[i] == 1 and (out[i] == 0).all()
        else:
            assert st[i] == 0
            r = lr.mfcc(to_f32([c])[0], lr.C1)
            _close(out[i:i + 1, :, :r.shape[1]].cpu().numpy(), r[None])
            assert (out[i, :, r.shape[1]:] == 0).all()


@pytest.mark.parametrize("noisy", [False, True])
def test_unaligned_packing(noisy):
    """Clips packed back to back at odd offsets (no 8-sample alignment): the vector / asynchronous staging paths
    must fall back to their scalar forms and still match."""
    import asr_b200 as A
    from oracle import noise_ref as nr
    lr = _o()
    lengths = [16000, 7001, 12345, 16000, 3333]
    clips = synth_clips(len(lengths), 0, 16000, 93, lengths=lengths)
    offsets = np.zeros(len(lengths), dtype=np.int64)
    offsets[1:] = np.cumsum(lengths[:-1]) + 1                       # +1: first clip at 0, the others at odd / arbitrary offsets
    total = int(offsets[-1] + lengths[-1]) + 3
    host = np.zeros(total, dtype=np.int16)
    zhost = np.zeros(total, dtype=np.float64)
    rng = np.random.default_rng(6)
    zs = [rng.standard_normal(n) for n in lengths]
    for c, z, o in zip(clips, zs, offsets):
        host[o:o + len(c)] = c
        zhost[o:o + len(c)] = z
    lens = np.asarray(lengths, dtype=np.int32)
    batch = A.ClipBatch(torch.from_numpy(host).cuda(), torch.from_numpy(offsets).cuda(), torch.from_numpy(lens).cuda(),
                        int(lens.max()), offsets, lens)
    plan = A.MfccPlan(A.C1, path=PATH)
    noise = None
    sig = np.array([0.01, 0.0, 0.02, 0.005, 0.03])
    if noisy:
        noise = A.Noise.white(torch.from_numpy(zhost).cuda(), torch.from_numpy(sig).cuda())
    out, st = plan.mfcc(batch, noise=noise)
    assert int(st.max()) == 0
    for i, (c, z) in enumerate(zip(to_f32(clips), zs)):
        x = nr.add_white_noise_z(c, sig[i], z) if noisy else c
        r = lr.mfcc(np.asarray(x), lr.C1)
        _close(out[i:i + 1, :, :r.shape[1]].cpu().numpy(), r[None], atol=3e-3)


@pytest.mark.parametrize("form", ["chunks", "clusters"])
@pytest.mark.parametrize("variant", ["mfcc", "logmel", "no_delta"])
def test_c5_full_length_cluster_path(variant, form, monkeypatch):
    """BASELINE configs[4] at its stated size: 10 s clips (160 000 samples, 1001 frames x 80 mel = 320 KB of log-mel rows,
    more than one SM holds).  Default form ("chunks"): the per-clip kernel takes chunks of frames per CTA, writes the log-mel
    rows to the workspace and the cepstra kernels of the pipelined paths finish the clip.  "clusters"
    (ASR_B200_CLIP_CLUSTERS): a clip spread over a thread-block cluster, clip maximum and delta halo through distributed
    shared memory.  Plus one shorter clip in the same batch (ragged: idle chunks / idle CTAs of its cluster)."""
    import asr_b200 as A
    lr = _o()
    if form == "clusters":
        monkeypatch.setenv("ASR_B200_CLIP_CLUSTERS", "1")
    else:
        monkeypatch.delenv("ASR_B200_CLIP_CLUSTERS", raising=False)
    clips = synth_clips(3, 160000, 16000, 55, lengths=[160000, 160000, 47111])
    P = A.C5 if variant != "no_delta" else A.C5.replace(delta_orders=0)
    Pr = lr.C5 if variant != "no_delta" else lr.C5.replace(delta_orders=0)
    plan = A.MfccPlan(P, path=PATH)
    batch = A.ClipBatch.from_arrays(clips)
    T = P.num_frames(160000)
    assert T == 1001
    if variant == "logmel":
        out, status = plan.logmel(batch, out_frames=T)
    else:
        out, status = plan.mfcc(batch, out_frames=T)
    torch.cuda.synchronize()
    assert int(status.max()) == 0
    for i, c in enumerate(clips):
        x = to_f32([c])[0]
        ref = lr.log_mel(x, Pr) if variant == "logmel" else lr.mfcc(x, Pr)
        t = ref.shape[1]
        # log-mel: the usual tolerance.  Cepstra: this clip has near-silent edge frames (Hann envelope), where the float32
        # FFT's log-mel error is largest (7e-4), and the lifter (22) scales coefficients 4..16 by up to 12 - hence 1e-2.
        if variant == "logmel":
            _close(out[i, :, :t].cpu().numpy(), ref)
        else:
            _close(out[i, :, :t].cpu().numpy(), ref, atol=1e-2)
        assert (out[i, :, t:] == 0).all()
