"""Device-side audio ingest: batched polyphase resampling against scipy.signal.resample_poly,
and the file-based drop-in entry point end to end (WAV on disk -> float64 (N, 880) rows)."""
import wave

import numpy as np
import pytest
import scipy.signal
import torch

from synth import synth_clips, to_f32

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("orig,target", [(16000, 22050), (44100, 22050), (8000, 22050), (22050, 16000)])
@pytest.mark.parametrize("as_int16", [True, False])
def test_resample_matches_scipy(orig, target, as_int16):
    import asr_b200 as A
    from math import gcd
    lengths = [orig, 1, 37, orig // 3, 2 * orig + 5, 160]
    clips = synth_clips(len(lengths), 0, orig, 71, lengths=lengths)
    xs = to_f32(clips)
    rs = A.Resampler(orig, target)
    out = rs(A.ClipBatch.from_arrays(clips if as_int16 else xs))
    got = out.unpack()
    g = gcd(orig, target)
    exact = 0
    for x, y in zip(xs, got):
        ref = scipy.signal.resample_poly(x, target // g, orig // g)
        assert y.dtype == np.float32 and y.shape == ref.shape
        # same taps, same summation order, separate multiply and add: equal up to the host compiler's FMA contraction
        assert np.abs(y - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
        exact += int(np.array_equal(y, ref))
    print("bit-identical clips:", exact, "of", len(xs))


def _write_wav(path, x_i16, sr):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes(x_i16.tobytes())


def test_compute_mfcc_all_files_end_to_end(tmp_path):
    """16 kHz PCM files -> device resample to 22 050 Hz -> REF-VDR MFCC -> (N, 880) float64, against the host path
    (scipy resample + oracle MFCC)."""
    from asr_b200 import audio_io
    from asr_b200.voice_digit import extract_features_construct_dataset as efcd
    from oracle import pipeline_ref as pr, librosa_ref as lr
    lengths = [16000, 12000, 16000, 9000]
    clips = synth_clips(len(lengths), 0, 16000, 72, lengths=lengths)
    paths = []
    for i, c in enumerate(clips):
        p = tmp_path / f"c{i}.wav"
        _write_wav(p, c, 16000)
        paths.append(str(p))
    got = efcd.compute_mfcc_all_files(paths)
    assert got.shape == (4, 880) and got.dtype == np.float64
    waves = [audio_io.load(p, sr=22050)[0] for p in paths]        # host: scipy.signal.resample_poly
    ref = pr.compute_mfcc_all_files(waves, 44, lr.REF_VDR)
    assert np.abs(got - ref).max() <= 3e-3
    b = audio_io.load_batch(paths, sr=22050)
    for w, y in zip(waves, b.unpack()):
        assert np.abs(w - y).max() <= 2e-6
