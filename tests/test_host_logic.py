"""CPU tests of the host-side logic above the C-ABI (no kernel is launched)."""
import os
import wave

import numpy as np
import pytest

import asr_b200 as A
from asr_b200 import audio_io, sharding
from asr_b200.speaker import extract_features_construct_dataset as sr_efcd
from asr_b200.voice_digit import extract_features_construct_dataset as vdr_efcd
from oracle import librosa_ref as lr, noise_ref as nr, pipeline_ref as pr
from synth import synth_clips, to_f32


def test_presets_match_the_oracle_presets():
    for name, p in A.PRESETS.items():
        o = lr.PRESETS[name]
        for f in ("sr", "n_fft", "win_length", "hop_length", "window", "center", "pad_mode", "fftfreq_mode", "n_mels",
                  "fmin", "fmax", "n_mfcc", "top_db", "amin", "lifter", "preemph", "delta_orders", "delta_width"):
            assert getattr(p, f) == getattr(o, f), (name, f)
    # the reference's real parameter sets (VDR/extract...py:30 ; SR/extract...py:227-228)
    assert (A.REF_VDR.n_fft, A.REF_VDR.hop_length, A.REF_VDR.n_mels, A.REF_VDR.n_mfcc, A.REF_VDR.window) == (2048, 512, 128, 20, "hann")
    assert (A.REF_SR.n_fft, A.REF_SR.win_length, A.REF_SR.hop_length) == (441, 441, 220)


@pytest.mark.parametrize("name", ["ref_vdr", "ref_sr", "c1", "c3", "c5"])
def test_num_frames_rule(name):
    p, o = A.PRESETS[name], lr.PRESETS[name]
    for L in (0, 1, p.n_fft // 2, p.n_fft // 2 + 1, p.n_fft, 9000, 16000, 22050, 160000):
        want = 0 if (o.pad_mode == "reflect" and L <= o.n_fft // 2) else lr.num_frames(o, L)
        assert p.num_frames(L) == want
    assert A.REF_VDR.num_frames(22050) == 44 and A.REF_SR.num_frames(22050) == 101      # comment at VDR/extract...py:17
    assert A.C3.feature_rows == 60 and A.C1.feature_rows == 13


def test_params_struct_round_trip():
    c = A.C5.to_c()
    assert (c.sr, c.n_fft, c.win_length, c.hop_length, c.window, c.n_mels, c.n_mfcc) == (16000, 1024, 1024, 160, 1, 80, 40)
    assert (c.delta_orders, c.delta_width, c.center, c.pad_mode, c.fftfreq_mode) == (1, 9, 1, 0, 0)
    assert abs(c.lifter - 22.0) < 1e-6 and abs(c.top_db - 80.0) < 1e-6


def test_clip_layout_is_aligned_and_disjoint():
    lengths = [1, 7, 8, 9, 16000, 3]
    off, total = A.ClipBatch.layout(lengths)
    assert off.dtype == np.int64 and (off % 8 == 0).all()
    ends = off + np.asarray(lengths)
    assert (off[1:] >= ends[:-1]).all() and total >= ends[-1] and total % 8 == 0
    off0, total0 = A.ClipBatch.layout([])
    assert len(off0) == 0 and total0 == 8


def test_snr_sigma_host_is_the_reference_chain():
    clips = to_f32(synth_clips(4, 4000, 16000, 3))
    P = np.array([np.mean(c ** 2) for c in clips], dtype=np.float32)
    for snr in (60, 30, 20, 15, 10, 5, 0):
        got = A.snr_sigma_host(P, snr)
        want = np.array([float(nr.snr_sigma(c, snr)) for c in clips])
        assert got.dtype == np.float64 and np.array_equal(got, want)


def test_snr_sigma_host_vectorised_equals_scalar_chain_on_a_million_powers():
    """The batch form used by the benchmarked pipeline (numpy log10 vectorised + libm powf chain in C) against the
    reference's scalar lines: zero mismatches over >= 10^6 float32 powers spanning 1e-8 .. 1 (and the edge values)."""
    rng = np.random.default_rng(11)
    mism = 0
    total = 0
    for snr in (0, 5, 10, 20):
        P = (10.0 ** rng.uniform(-8.0, 0.0, 262144)).astype(np.float32)
        P[:6] = np.array([0.0, 1.0, 1e-8, np.float32(2.0) ** -126, 3.0517578e-05, 0.99999994], dtype=np.float32)
        with np.errstate(divide="ignore"):
            want = A.snr_sigma_host_scalar(P, snr)
        got = A.snr_sigma_host(P, snr)
        mism += int((got.view(np.uint64) != want.view(np.uint64)).sum())
        total += P.shape[0]
    assert total >= 10 ** 6 and mism == 0
    # python-float and float32 targets keep the chain in float32; a float64 numpy scalar promotes it (literal text)
    P = (10.0 ** rng.uniform(-6.0, 0.0, 512)).astype(np.float32)
    assert np.array_equal(A.snr_sigma_host(P, 7.5), A.snr_sigma_host_scalar(P, 7.5))
    assert np.array_equal(A.snr_sigma_host(P, np.float32(7.3)), A.snr_sigma_host_scalar(P, np.float32(7.3)))
    assert np.array_equal(A.snr_sigma_host(P, np.float64(7.3)), A.snr_sigma_host_scalar(P, np.float64(7.3)))


def test_sr_window_index_matches_reference_trim_split():
    """SR/extract...py:211-222 via index arithmetic == slicing the waveform."""
    sr = 50
    lengths = [0, 49, 100, 149, 150, 151, 537, 1000]
    fid, start = sr_efcd.window_index(lengths, sr)
    k = 0
    for i, n in enumerate(lengths):
        y = np.arange(n, dtype=np.float32) + 1000 * i
        for w in pr.sr_trim_split(y, sr):
            assert fid[k] == i and np.array_equal(y[start[k]:start[k] + sr], w)
            k += 1
    assert k == len(fid)


def test_file_listing_and_labels(tmp_path):
    """The reference's fixed class list, in list order, other folders ignored (VDR/extract...py:118-140, SR :113-137)."""
    for d, n in (("zero", 2), ("one", 1), ("two", 3), ("bed", 2), ("nine", 1)):
        os.makedirs(tmp_path / d)
        for i in range(n):
            (tmp_path / d / f"a{i}.wav").write_bytes(b"")
    files, labels = vdr_efcd.get_file_names_and_labels(str(tmp_path))
    assert labels.dtype == np.int32 and len(files) == 7          # "bed" is not a digit: ignored
    by = {os.path.basename(os.path.dirname(f)): l for f, l in zip(files, labels)}
    assert by == {"zero": 0, "one": 1, "two": 2, "nine": 3}       # position among the PRESENT classes, list order
    assert list(labels) == sorted(labels)
    # all ten present -> label = digit value
    for d in vdr_efcd.DIGITS:
        os.makedirs(tmp_path / d, exist_ok=True)
        (tmp_path / d / "z.wav").write_bytes(b"")
    files, labels = vdr_efcd.get_file_names_and_labels(str(tmp_path))
    by = {os.path.basename(os.path.dirname(f)): l for f, l in zip(files, labels)}
    assert by == {d: i for i, d in enumerate(vdr_efcd.DIGITS)}
    # speaker corpus: 20 listed IDs, in list order
    sp = tmp_path / "spk"
    for d in ("420", "006", "999", "105"):
        os.makedirs(sp / d)
        (sp / d / "u.wav").write_bytes(b"")
    files2, labels2 = sr_efcd.get_file_names_and_labels(str(sp))
    by2 = {os.path.basename(os.path.dirname(f)): l for f, l in zip(files2, labels2)}
    assert by2 == {"006": 0, "105": 1, "420": 2} and labels2.dtype == np.int32


def test_wav_decode_and_resample(tmp_path):
    sr = 16000
    x = synth_clips(1, sr, sr, 2)[0]
    path = str(tmp_path / "a.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes(x.tobytes())
    y, got_sr = audio_io.load(path, sr=None)
    assert got_sr == sr and y.dtype == np.float32 and np.array_equal(y, x.astype(np.float32) / 32768.0)
    y2, sr2 = audio_io.load(path, sr=22050)                 # librosa.load default: resample to 22 050 Hz
    assert sr2 == 22050 and abs(len(y2) - 22050) <= 1 and y2.dtype == np.float32
    st = np.stack([x, x], axis=1)                           # stereo -> mono mean
    with wave.open(path, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes(st.tobytes())
    y3, _ = audio_io.load(path, sr=None)
    assert np.allclose(y3, y)


@pytest.mark.parametrize("n,world", [(0, 1), (1, 2), (7, 2), (8192, 8), (1000003, 8), (5, 8)])
def test_shard_bounds_partition(n, world):
    prev = 0
    sizes = []
    for r in range(world):
        lo, hi = sharding.shard_bounds(n, r, world)
        assert lo == prev and hi >= lo
        sizes.append(hi - lo)
        prev = hi
    assert prev == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(n, world, world)


def test_first_sample_index_matches_the_packed_layout():
    lengths = [16000, 15999, 3, 8, 22050, 100]
    off, _ = A.ClipBatch.layout(lengths)
    for world in (1, 2, 3):
        for r in range(world):
            lo, hi = sharding.shard_bounds(len(lengths), r, world)
            want = int(off[lo]) if lo < len(lengths) else int(off[-1] + (lengths[-1] + 7) // 8 * 8)
            assert sharding.first_sample_index(lengths, r, world) == want


@pytest.mark.parametrize("orig,target", [(16000, 22050), (44100, 22050), (8000, 22050), (48000, 22050), (22050, 16000)])
def test_resample_filter_design_is_scipy_firwin(orig, target):
    """asr_resample_design (host code of the library, no GPU) == the taps scipy.signal.resample_poly builds for float32."""
    import ctypes as C
    from math import gcd
    import scipy.signal
    from asr_b200._lib import lib
    u, d, n, pre = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    assert lib.asr_resample_design(target, orig, 5.0, None, 0, C.byref(u), C.byref(d), C.byref(n), C.byref(pre)) == 0
    g = gcd(orig, target)
    assert (u.value, d.value) == (target // g, orig // g)
    taps = np.zeros(n.value, np.float32)
    assert lib.asr_resample_design(target, orig, 5.0, taps.ctypes.data, n.value, C.byref(u), C.byref(d), C.byref(n), C.byref(pre)) == 0
    mr = max(u.value, d.value)
    half = 10 * mr
    h = scipy.signal.firwin(2 * half + 1, 1.0 / mr, window=("kaiser", 5.0)).astype(np.float32)
    h *= u.value
    n_pre_pad = d.value - half % d.value
    ref = np.concatenate([np.zeros(n_pre_pad, np.float32), h])
    assert n.value == len(ref) and pre.value == (half + n_pre_pad) // d.value
    np.testing.assert_allclose(taps, ref, rtol=2e-7, atol=1e-12)
    for n_in in (1, 2, 159, 160, 16000, 22051):
        assert lib.asr_resample_out_len(n_in, target, orig) == len(scipy.signal.resample_poly(np.zeros(n_in, np.float32), u.value, d.value))


def test_processed_dataset_files_round_trip(tmp_path):
    """The six .npy files the reference's training / attack scripts load: names, dtypes, C-order float64 rows."""
    from asr_b200 import dataset_io
    rng = np.random.default_rng(3)
    files = np.array([f"data\\\\seven\\\\f{i}.wav" for i in range(23)])
    labels = np.arange(23, dtype=np.int32) % 10
    (f_tr, f_dev, f_te), (l_tr, l_dev, l_te) = dataset_io.split_70_20_10(files, labels)
    assert (len(f_tr), len(f_dev), len(f_te)) == (16, 4, 2)            # int(23*0.7), int(23*0.9)-16, int(23*0.1)
    assert list(f_te) == list(files[-2:]) and list(l_te) == list(labels[-2:])
    data = [rng.standard_normal((len(f), 880)).astype(np.float32) for f in (f_tr, f_dev, f_te)]
    d = str(tmp_path / "processed_google_dataset")
    dataset_io.save_processed_dataset(d, data, (l_tr, l_dev, l_te), str(tmp_path / "test_dataset_to_add_noise"), f_te)
    got = dataset_io.load_npy_dataset(d + os.sep)
    for name, x, y, gx, gy in zip(("train", "dev", "test"), data, (l_tr, l_dev, l_te), got[0::2], got[1::2]):
        assert gx.dtype == np.float64 and gx.flags["C_CONTIGUOUS"] and gx.shape == x.shape
        assert np.array_equal(gx, x.astype(np.float64)) and np.array_equal(gy, y)
        raw = open(os.path.join(d, f"{name}_data.npy"), "rb").read()
        assert len(raw) == 128 + 8 * x.size                              # 128-byte header + float64 payload (LFS pointer size rule)
    assert np.array_equal(np.load(str(tmp_path / "test_dataset_to_add_noise" / "test_filenames.npy")), f_te)
    # shard merge by index == single-process row order
    full = rng.standard_normal((10, 7))
    parts = [full[slice(*sharding.shard_bounds(10, r, 3))] for r in range(3)]
    assert np.array_equal(dataset_io.merge_shards(parts), full)
