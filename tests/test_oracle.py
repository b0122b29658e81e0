"""CPU tests of the oracle itself (no GPU, no product code).

The reference ships no tests, golden vectors or feature matrices and its MFCC arithmetic lives in
librosa, which is not in the image - PARITY UNPINNED.  What can be pinned is done here:

* the restatement against two INDEPENDENT implementations that are in the image
  (``torchaudio.transforms.MFCC`` with librosa-compatible settings - fp32 FFT, so a tolerance of 1e-3 - and the numpy
  ``transformers.audio_utils`` spectrogram / Slaney bank / power_to_db, written after librosa - float64, agreeing to
  1e-5 dB on the log-mel matrix);
* the reference's own pure-numpy code (noise mixers, StandardScaler standardisation) restated verbatim
  in behaviour, against numpy / sklearn run directly;
* the structural pins the reference text holds (880 = 20 x 44, 2020 = 20 x 101, T = 1 + L // hop,
  zero padding in the feature domain, float64 row-major rows);
* the committed golden fixtures (regression pin of the oracle).
"""
import os

import numpy as np
import pytest
import scipy.signal

from oracle import librosa_ref as lr, noise_ref as nr, cmvn_ref as cr, pipeline_ref as pr
from synth import synth_clips, to_f32

GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ---- golden fixtures (oracle regression pin) ------------------------------------------------------
@pytest.mark.parametrize("name", ["ref_vdr", "ref_sr", "c1", "c3", "c5"])
def test_oracle_reproduces_golden_mfcc(name):
    g = np.load(os.path.join(GOLD, f"mfcc_{name}.npz"))
    p = lr.PRESETS[name]
    for clip, want in zip(g["audio_i16"], g["mfcc"]):
        x = clip.astype(np.float32) / np.float32(32768.0)
        if name == "ref_sr":
            x = x.astype(np.float64)
        got = lr.mfcc(x, p)
        assert got.dtype == want.dtype and got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-4)   # BLAS summation order may differ across hosts


def test_oracle_reproduces_golden_noise_and_cmvn():
    g = np.load(os.path.join(GOLD, "noise.npz"))
    clips = list(g["audio_f32"])
    for snr in (0, 5, 10, 20):
        np.random.seed(1234 + snr)
        got = np.stack([nr.add_white_noise_with_snr(c, snr) for c in clips])
        assert np.array_equal(got.view(np.uint64), g[f"snr{snr}"].view(np.uint64))
    np.random.seed(77)
    assert np.array_equal(np.stack([nr.add_white_noise(c, 0.01) for c in clips]), g["white"])
    np.random.seed(78)
    assert np.array_equal(np.stack([nr.add_noise(c, 0.01, 0.002) for c in clips]), g["mixture"])
    assert np.array_equal(np.array([nr.mean_power_f32(c) for c in clips], dtype=np.float32), g["power"])
    c = np.load(os.path.join(GOLD, "cmvn.npz"))
    sa, sb, sc = cr.standardize_dataset(c["a"], c["b"], c["c"])
    for got, want in ((sa, c["sa"]), (sb, c["sb"]), (sc, c["sc"])):
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-13)


# ---- independent implementation: torchaudio (fp32 FFT) --------------------------------------------
def _torchaudio_mfcc(x_f32, p):
    import torch
    import torchaudio
    t = torchaudio.transforms.MFCC(
        sample_rate=p.sr, n_mfcc=p.n_mfcc, dct_type=2, norm="ortho", log_mels=False,
        melkwargs=dict(n_fft=p.n_fft, win_length=p.win_length or p.n_fft, hop_length=p.hop_length, f_min=p.fmin,
                       f_max=p.fmax or p.sr / 2, n_mels=p.n_mels, center=True, pad_mode="reflect", power=2.0,
                       norm="slaney", mel_scale="slaney",
                       window_fn=torch.hann_window if p.window == "hann" else torch.hamming_window))
    t.amplitude_to_DB.top_db = p.top_db
    out = t(torch.from_numpy(x_f32)[None])[0].numpy()
    if p.lifter > 0:
        out = out * (1 + (p.lifter / 2) * np.sin(np.pi * np.arange(1, 1 + p.n_mfcc) / p.lifter))[:, None]
    return out


@pytest.mark.parametrize("name,tol", [("ref_vdr", 1e-3), ("c1", 1e-3)])
def test_oracle_vs_torchaudio(name, tol):
    p = lr.PRESETS[name]
    for x in to_f32(synth_clips(3, p.sr, p.sr, 5)):
        want = _torchaudio_mfcc(x, p)
        got = lr.mfcc(x, p)[:p.n_mfcc]
        assert got.shape == want.shape
        assert np.abs(got - want).max() < tol


@pytest.mark.parametrize("name", ["ref_vdr", "ref_sr", "c1", "c3", "c5"])
def test_mel_filterbank_vs_torchaudio(name):
    import torchaudio
    p = lr.PRESETS[name]
    fb = torchaudio.functional.melscale_fbanks(1 + p.n_fft // 2, p.fmin, p.fmax or p.sr / 2, p.n_mels, p.sr,
                                               norm="slaney", mel_scale="slaney").numpy().T
    W = lr.mel_filterbank(p)        # torchaudio uses the linspace frequency grid = librosa 0.9
    assert W.dtype == np.float32 and W.shape == fb.shape
    assert np.abs(W - fb).max() < 1e-6
    # <= 2 filters per bin (what the sparse kernel layout relies on)
    assert int((W != 0).sum(axis=0).max()) <= 2


def test_fftfreq_mode_switch_matters_for_odd_nfft():
    a = lr.mel_filterbank(lr.REF_SR)
    b = lr.mel_filterbank(lr.REF_SR.replace(fftfreq_mode="rfftfreq"))
    assert np.abs(a - b).max() > 1e-3          # librosa 0.9 vs >= 0.10 differ for n_fft = 441
    c = lr.mel_filterbank(lr.C1)
    d = lr.mel_filterbank(lr.C1.replace(fftfreq_mode="rfftfreq"))
    assert np.abs(c - d).max() < 1e-6          # and agree for even n_fft


# ---- structural pins the reference text holds ------------------------------------------------------
def test_reference_shapes():
    # VDR: 22 050 samples, hop 512 -> 44 frames; 20 x 44 = 880 (VDR/train_constraints.py:66)
    assert lr.num_frames(lr.REF_VDR, 22050) == 44
    # SR: 1-s windows of 22 050 samples, n_fft 441 / hop 220 -> 101 frames; 20 x 101 = 2020
    assert lr.num_frames(lr.REF_SR, 22050) == 101
    x = to_f32(synth_clips(1, 22050, 22050, 3))[0]
    assert lr.mfcc(x, lr.REF_VDR).shape == (20, 44)
    assert lr.mfcc(x.astype(np.float64), lr.REF_SR).shape == (20, 101)
    rows = pr.compute_mfcc_all_files([x, x[:9000]], 44)
    assert rows.shape == (2, 880) and rows.dtype == np.float64
    # zero padding in the feature domain (VDR/extract...py:36-37): 9 000 samples -> 18 frames, rest zeros
    blk = rows[1].reshape(20, 44)
    assert np.all(blk[:, 18:] == 0) and np.all(np.abs(blk[:, :18]).sum(axis=0) > 0)


def test_dtype_flow():
    x = to_f32(synth_clips(1, 8000, 16000, 4))[0]
    assert lr.mfcc(x, lr.C1).dtype == np.float32                    # complex64 STFT, float32 mel basis
    assert lr.mfcc(x.astype(np.float64), lr.C1).dtype == np.float64
    with pytest.raises(TypeError):
        lr.mfcc((x * 32768).astype(np.int16), lr.C1)               # librosa.util.valid_audio


def test_too_short_raises_like_np_pad():
    with pytest.raises(ValueError):
        lr.mfcc(np.zeros(256, np.float32), lr.C1)                   # len <= n_fft//2 with reflect padding
    with pytest.raises(ValueError):
        lr.mfcc(np.zeros(16000, np.float32), lr.C3.replace(hop_length=4000))   # 5 frames < delta width 9


def test_top_db_is_clip_wide():
    p = lr.C1
    x = to_f32(synth_clips(1, 16000, 16000, 6))[0]
    x[:8000] *= 1e-4                                                # first half 80 dB down: clamped by the loud half
    L = lr.log_mel(x, p)
    assert np.isclose(L.min(), L.max() - 80.0, atol=1e-4)
    Lhalf = lr.log_mel(x[:8000], p)                                 # alone it is not clamped at that level
    assert Lhalf.max() < L.max() - 60


def test_delta_is_savgol():
    rng = np.random.default_rng(0)
    C = rng.standard_normal((5, 40))
    d1 = lr.delta(C, 9, 1)
    taps1 = np.arange(-4, 5) / 60.0
    assert np.allclose(d1[:, 4:-4], np.stack([np.correlate(r, taps1, "valid") for r in C]))
    taps2 = np.array([28, 7, -8, -17, -20, -17, -8, 7, 28]) / 462.0
    assert np.allclose(lr.delta(C, 9, 2)[:, 4:-4], np.stack([np.correlate(r, taps2, "valid") for r in C]))
    assert np.allclose(d1, scipy.signal.savgol_filter(C, 9, polyorder=1, deriv=1, axis=-1, mode="interp"))


# ---- the reference's own numpy code ------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 7, 8, 100, 128, 129, 1000, 4097, 16000, 22050])
def test_pairwise_sum_matches_numpy(n):
    rng = np.random.default_rng(n)
    a = (rng.standard_normal(n) ** 2).astype(np.float32)
    assert nr.pairwise_sum_f32(a) == np.sum(a)
    x = rng.standard_normal(n).astype(np.float32)
    assert nr.mean_power_f32(x) == np.mean(x ** 2)


def test_snr_mix_restatement_is_the_reference_text():
    """VDR/attacks.py:222-245 executed literally (numpy global RNG) equals the z-parameterised restatement."""
    x = to_f32(synth_clips(1, 5000, 16000, 9))[0]
    for snr in (60, 30, 20, 15, 10, 5, 0):
        np.random.seed(snr)
        sample = x
        signal_avg_watts = np.mean(sample ** 2)
        signal_avg_db = 10 * np.log10(signal_avg_watts)
        noise_avg_db = signal_avg_db - snr
        noise_avg_watts = 10 ** (noise_avg_db / 10)
        noise = 1 * np.random.normal(0, np.sqrt(noise_avg_watts), len(sample))
        want = sample + noise
        np.random.seed(snr)
        z = np.random.standard_normal(len(x))
        got = nr.add_white_noise_with_snr_z(x, snr, z)
        assert want.dtype == np.float64 and np.array_equal(got.view(np.uint64), want.view(np.uint64))
        np.random.seed(snr)
        assert np.array_equal(nr.add_white_noise_with_snr(x, snr), want)


def test_mixture_restatement_is_the_reference_text():
    """VDR/attacks.py:145-183: two normal draws per sample, all selectors first."""
    x = to_f32(synth_clips(1, 3000, 16000, 10))[0]
    N, p, alpha = len(x), 0.01, 0.003
    np.random.seed(5)
    q = np.random.randn(N)
    u = q.copy()
    flag1 = abs(q) < p
    flag0 = abs(q) >= p
    u[flag1] = 1
    u[flag0] = 0
    noise = (alpha * (1 - u) + 10 * alpha * u) * np.random.randn(N)
    want = x + noise
    np.random.seed(5)
    q2 = np.random.standard_normal(N)
    g2 = np.random.standard_normal(N)
    got = nr.add_noise_z(x, p, alpha, q2, g2)
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def test_standardize_is_sklearn():
    from sklearn.preprocessing import StandardScaler
    rng = np.random.default_rng(1)
    a, b, c = rng.standard_normal((50, 33)) * 5 + 2, rng.standard_normal((20, 33)), rng.standard_normal((9, 33))
    a[:, 3] = b[:, 3] = c[:, 3] = 0.0
    allx = np.concatenate([a, b, c])
    want = StandardScaler().fit_transform(allx)
    sa, sb, sc = cr.standardize_dataset(a, b, c)
    assert np.array_equal(np.concatenate([sa, sb, sc]), want)
    mean, var, scale = cr.column_stats(allx)
    assert scale[3] == 1.0 and np.allclose(mean, allx.mean(0)) and np.allclose(var, allx.var(0))


def test_sr_trim_split():
    """SR/extract...py:211-222: drop the first second and everything after (floor(len/sr)-1)*sr, 1-s windows."""
    sr = 100
    y = np.arange(537, dtype=np.float32)
    w = pr.sr_trim_split(y, sr)
    assert len(w) == 3 and all(len(v) == sr for v in w)
    assert w[0][0] == 100 and w[-1][-1] == 399


# ---- second independent implementation: transformers.audio_utils (numpy, float64 FFT, written after librosa) ----
@pytest.mark.parametrize("name", ["ref_vdr", "ref_sr", "c1", "c3", "c5"])
def test_log_mel_vs_transformers_audio_utils(name):
    """``transformers.audio_utils.spectrogram`` + ``mel_filter_bank(norm='slaney', mel_scale='slaney')`` restate the same
    librosa stages (reflect-padded centred frames, periodic window, |rfft|^2, Slaney bank, power_to_db with an 80 dB
    range over the whole call) in plain numpy: a witness that shares no code with the oracle or with torchaudio."""
    au = pytest.importorskip("transformers.audio_utils")
    p = lr.PRESETS[name]
    win_length = p.win_length if p.win_length > 0 else p.n_fft
    window = scipy.signal.get_window(p.window, win_length, fftbins=True)
    bank = au.mel_filter_bank(1 + p.n_fft // 2, p.n_mels, p.fmin, p.fmax if p.fmax > 0 else p.sr / 2, p.sr,
                              norm="slaney", mel_scale="slaney")
    np.testing.assert_allclose(bank.T, lr.mel_filterbank(p), rtol=0, atol=1e-7)      # float32 storage of the oracle's bank
    for x in to_f32(synth_clips(2, p.sr, p.sr, 31)):
        if name == "ref_sr":
            x = x.astype(np.float64)      # the speaker script hands float64 windows to librosa (complex128 STFT; odd n_fft = 441)
        want = au.spectrogram(x.astype(np.float64), window, frame_length=win_length, hop_length=p.hop_length, fft_length=p.n_fft,
                              power=2.0, center=True, pad_mode="reflect", mel_filters=bank, mel_floor=p.amin, log_mel="dB",
                              reference=1.0, min_value=p.amin, db_range=p.top_db, dtype=np.float64)
        got = lr.log_mel(x, p)
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-4)    # measured 9e-6 dB: float32 storage of |D|^2 in the oracle (librosa's dtype flow)


def test_oracle_reproduces_librosa_docstring_examples():
    """Known answers PUBLISHED by the dependency that holds the algorithm (librosa 0.9 docstrings; librosa itself is not in
    the image): `librosa.hz_to_mel(60)` -> 0.9, `hz_to_mel([110, 220, 440])` -> [1.65, 3.3, 6.6]; `librosa.mel_to_hz(3)` ->
    200., `mel_to_hz([1, 2, 3, 4, 5])` -> [66.667, 133.333, 200., 266.667, 333.333]; `librosa.mel_frequencies(n_mels=40)`
    (fmin = 0, fmax = 11025) -> the 40 values below; `librosa.filters.mel(sr=22050, n_fft=2048)` prints
    `[[0., 0.016, ..., 0., 0.], ...]` (128 x 1025).  These pin the Slaney mel scale and the filter bank of the reference's
    own call (VDR/extract_features_construct_dataset.py:30 uses exactly sr = 22050, n_fft = 2048, 128 filters) to the
    printed precision; the STFT / dB / DCT stages stay pinned only through the independent implementations above."""
    assert np.allclose(lr.hz_to_mel(np.array([60.0])), [0.9], atol=5e-4)
    assert np.allclose(lr.hz_to_mel(np.array([110.0, 220.0, 440.0])), [1.65, 3.3, 6.6], atol=5e-4)
    assert np.allclose(lr.mel_to_hz(np.array([3.0])), [200.0], atol=5e-4)
    assert np.allclose(lr.mel_to_hz(np.array([1.0, 2, 3, 4, 5])), [66.667, 133.333, 200.0, 266.667, 333.333], atol=5e-4)
    doc = np.array([0., 85.317, 170.635, 255.952, 341.269, 426.586, 511.904, 597.221, 682.538, 767.855, 853.173, 938.49,
                    1024.856, 1119.114, 1222.042, 1334.436, 1457.167, 1591.187, 1737.532, 1897.337, 2071.84, 2262.393,
                    2470.47, 2697.686, 2945.799, 3216.731, 3512.582, 3835.643, 4188.417, 4573.636, 4994.285, 5453.621,
                    5955.205, 6502.92, 7101.009, 7754.107, 8467.272, 9246.028, 10096.408, 11025.])
    mels = np.linspace(lr.hz_to_mel(np.array([0.0]))[0], lr.hz_to_mel(np.array([11025.0]))[0], 40)
    assert np.allclose(lr.mel_to_hz(mels), doc, atol=6e-4)
    fb = lr.mel_filterbank(lr.PRESETS["ref_vdr"])
    assert fb.shape == (128, 1025)
    assert [round(float(v), 3) + 0.0 for v in (fb[0, 0], fb[0, 1], fb[0, -2], fb[0, -1])] == [0.0, 0.016, 0.0, 0.0]
    assert not np.round(fb[[1, -2, -1]][:, [0, 1, -2, -1]], 3).any()      # the other printed corners: 0. to three decimals
