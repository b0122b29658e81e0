"""Audio ingest in front of the hot path: WAV decode + resample (the reference's ``librosa.load``).

``librosa.load(path, mono=True)`` (VDR/extract_features_construct_dataset.py:27) decodes with libsndfile, averages
channels and resamples to 22 050 Hz with resampy's ``kaiser_best`` filter.  Neither libsndfile nor resampy's
filter table is in the image, so:

* decode = ``scipy.io.wavfile`` on the host (PCM header parse + memory copy; int16 stays int16),
* resample = ``scipy.signal.resample_poly`` semantics.  ``load`` does it on the host for one file (scipy itself);
  ``load_batch`` uploads the decoded PCM of many files once and resamples them in ONE launch on the device
  (``frontend.Resampler`` -> ``asr_resample_batch``), which is what the batched drop-in entry points use.

Same contract as ``librosa.load`` (float32 mono in [-1, 1) at ``sr``); not bit-identical samples to resampy.
"""
from __future__ import annotations

from functools import lru_cache
from math import gcd

import numpy as np
import scipy.io.wavfile
import scipy.signal


def _decode(path, mono: bool = True):
    """PCM WAV -> (native_sr, int16 or float32 mono array)."""
    native_sr, data = scipy.io.wavfile.read(path)
    if data.ndim == 2 and mono:
        if data.dtype == np.int16:
            data = data.astype(np.float32).mean(axis=1) / np.float32(32768.0)     # librosa: to_mono on floats
        else:
            data = data.mean(axis=1)
    if data.dtype == np.int16:
        return native_sr, np.ascontiguousarray(data)
    if data.dtype == np.int32:
        return native_sr, (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    if data.dtype == np.uint8:
        return native_sr, (data.astype(np.float32) - 128.0) / 128.0
    return native_sr, np.ascontiguousarray(data, dtype=np.float32)


def load(path, sr: int = 22050, mono: bool = True):
    """One file, on the host (``librosa.load`` signature)."""
    native_sr, data = _decode(path, mono)
    y = data.astype(np.float32) / np.float32(32768.0) if data.dtype == np.int16 else data
    if y.ndim == 2:
        y = y.T
    if sr is not None and native_sr != sr:
        g = gcd(int(sr), int(native_sr))
        y = scipy.signal.resample_poly(y, sr // g, native_sr // g, axis=-1).astype(np.float32)
        native_sr = sr
    return np.ascontiguousarray(y, dtype=np.float32), native_sr


@lru_cache(maxsize=16)
def _resampler(native_sr: int, sr: int, device_index: int):
    from .frontend import Resampler
    return Resampler(native_sr, sr, device=f"cuda:{device_index}")


def load_batch(paths, sr: int = 22050):
    """Many files -> one packed float32 ``ClipBatch`` at ``sr`` on the current device, in the order of ``paths``.

    Files are decoded on the host, uploaded as PCM (int16 stays int16: half the PCIe bytes of float32) and
    resampled on the device, one launch per distinct source rate."""
    import torch
    from .frontend import ClipBatch
    paths = list(paths)
    if not paths:
        return ClipBatch.from_arrays([], dtype=np.float32)
    decoded = [_decode(p) for p in paths]
    dev = torch.cuda.current_device()
    groups = {}
    for i, (nsr, data) in enumerate(decoded):
        groups.setdefault((nsr, data.dtype), []).append(i)
    pieces = [None] * len(paths)
    for (nsr, _), idx in groups.items():
        b = ClipBatch.from_arrays([decoded[i][1] for i in idx])
        r = _resampler(int(nsr), int(sr), dev)(b) if nsr != sr or b.audio.dtype != torch.float32 else b
        for k, i in enumerate(idx):
            o, n = int(r.offsets_host[k]), int(r.lengths_host[k])
            pieces[i] = r.audio[o:o + n]
    if len(groups) == 1:                        # common case: one source rate -> the resampler's batch is the answer
        return r
    lengths = np.array([p.numel() for p in pieces], dtype=np.int32)
    offsets, total = ClipBatch.layout(lengths)
    audio = torch.zeros(total, dtype=torch.float32, device=pieces[0].device if pieces else "cuda")
    for p, o in zip(pieces, offsets):
        audio[o:o + p.numel()] = p
    return ClipBatch(audio, torch.from_numpy(offsets).to(audio.device), torch.from_numpy(lengths).to(audio.device),
                     int(lengths.max()) if len(lengths) else 0, offsets, lengths)
