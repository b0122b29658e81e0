"""Host glue in front of the hot path: WAV decode + resample (the reference's ``librosa.load``).

``librosa.load(path, mono=True)`` (VDR/extract_features_construct_dataset.py:27) decodes with
libsndfile, averages channels and resamples to 22 050 Hz with resampy's ``kaiser_best`` filter.
Neither libsndfile nor resampy is in the image, so this decodes PCM WAV with ``scipy.io.wavfile``
and resamples with ``scipy.signal.resample_poly`` - same contract (float32 mono in [-1, 1) at
``sr``), not bit-identical samples.  SURVEY.md 8(f) lists a device-side ingest as the next row.
"""
from __future__ import annotations

from math import gcd

import numpy as np
import scipy.io.wavfile
import scipy.signal


def load(path, sr: int = 22050, mono: bool = True):
    native_sr, data = scipy.io.wavfile.read(path)
    if data.dtype == np.int16:
        y = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        y = (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    elif data.dtype == np.uint8:
        y = (data.astype(np.float32) - 128.0) / 128.0
    else:
        y = data.astype(np.float32)
    if y.ndim == 2:
        y = y.mean(axis=1) if mono else y.T
    if sr is not None and native_sr != sr:
        g = gcd(int(sr), int(native_sr))
        y = scipy.signal.resample_poly(y, sr // g, native_sr // g, axis=-1).astype(np.float32)
        native_sr = sr
    return np.ascontiguousarray(y, dtype=np.float32), native_sr
