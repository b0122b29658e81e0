"""Drop-in for the hot-path functions of ``Speaker recognition/attacks.py``.

``black_box_attack_on_audio_dataset(filenames, labels, sigma, p, alpha)`` :97-146 and
``black_box_attack_on_audio_snr(filenames, labels, target_snr_db)`` :254-295 add the noise to the
WHOLE file first, then trim / split into 1-s windows, then take the 441/220 MFCC of each window.
Here the per-file sigma comes from the bit-exact power kernel, the windows are index ranges into
the packed file audio and the mix is fused into the MFCC launch (the float64 noisy signal is never
materialised).  The noise-only helpers are shared with the voice-digit task (the reference files
duplicate them verbatim: SR/attacks.py:81-94,149-251 == VDR/attacks.py:73-86,145-245).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import audio_io
from ..frontend import ClipBatch, Noise, clip_power, snr_sigma_host
from ..voice_digit.attacks import (load_npy_dataset, standardize_dataset, add_white_noise, mixtgauss, add_noise,  # noqa: F401
                                   add_white_noise_on_dataset, add_noise_mixture_on_dataset, add_white_noise_with_snr,
                                   _as_audio, _draw_like)
from . import extract_features_construct_dataset as efcd


def _window_sigma(sigma_file: np.ndarray, fid: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(sigma_file[fid], dtype=np.float64)).to(device)


def black_box_attack_on_waveforms_dataset(waves, labels, sigma=0, p=0, alpha=0, params=None):
    batch = waves if isinstance(waves, ClipBatch) else ClipBatch.from_arrays([_as_audio(w) for w in waves])
    prm = efcd.PARAMS if params is None else params
    _, fid = efcd.windows_of(batch, prm.sr)
    noise = None
    if sigma != 0:
        (z,) = _draw_like(batch)
        noise = Noise.white(z, torch.full((len(fid),), float(sigma), dtype=torch.float64, device=z.device))
    elif p != 0 and alpha != 0:
        q, g = _draw_like(batch, 2)
        noise = Noise.mixture(q, g, p, alpha)
    mfcc, fid = efcd.mfcc_windows(batch, noise, params)
    return mfcc, np.copy(np.asarray(labels))[fid]


def black_box_attack_on_waveforms_snr(waves, labels, target_snr_db, params=None):
    batch = waves if isinstance(waves, ClipBatch) else ClipBatch.from_arrays([_as_audio(w) for w in waves])
    prm = efcd.PARAMS if params is None else params
    sigma_file = snr_sigma_host(clip_power(batch).cpu().numpy(), target_snr_db)
    (z,) = _draw_like(batch)
    _, fid = efcd.windows_of(batch, prm.sr)
    mfcc, fid = efcd.mfcc_windows(batch, Noise.white(z, _window_sigma(sigma_file, fid, z.device)), params)
    return mfcc, np.copy(np.asarray(labels))[fid]


def _load_all(filenames):
    """Decode on the host, resample on the device: one packed batch for the whole file list."""
    return audio_io.load_batch(list(filenames), sr=efcd.PARAMS.sr)


def black_box_attack_on_audio_dataset(filenames, labels, sigma=0, p=0, alpha=0):
    return black_box_attack_on_waveforms_dataset(_load_all(filenames), labels, sigma, p, alpha)


def black_box_attack_on_audio_snr(filenames, labels, target_snr_db):
    return black_box_attack_on_waveforms_snr(_load_all(filenames), labels, target_snr_db)
