"""Drop-in for the hot-path functions of ``Voice digit recogniton/attacks.py``.

Mirrored (same names / argument order / return layout):
``load_npy_dataset`` :27-45, ``standardize_dataset`` :48-69, ``add_white_noise`` :73-86,
``black_box_attack_on_audio`` :89-121, ``black_box_attack_on_audio_dataset`` :124-142,
``mixtgauss`` :145-162, ``add_noise`` :165-183, ``add_white_noise_on_dataset`` :186-201,
``add_noise_mixture_on_dataset`` :204-219, ``add_white_noise_with_snr`` :222-245,
``black_box_attack_on_audio_snr`` :248-274, ``black_box_attack_on_audio_dataset_snr`` :277-294.

Random numbers: like the reference these functions draw from numpy's GLOBAL legacy RNG, in the
same order and amount (``np.random.normal(0, s, n)`` is bit-identical to ``s *
np.random.standard_normal(n)``), so ``np.random.seed(k)`` before a call gives the reference's noise.
The arithmetic (power reduction, mix, MFCC, standardisation) runs on the GPU.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import audio_io
from ..frontend import (ClipBatch, Noise, Standardizer, clip_power, snr_sigma_host, mix_white, mix_mixture,
                        mix_rows_white, mix_rows_mixture)
from . import extract_features_construct_dataset as efcd

UTTERANCE_LENGTH = 44


def load_npy_dataset(path):
    train_data = np.load(path + "train_data.npy")
    train_label = np.load(path + "train_label.npy")
    val_label = np.load(path + "dev_label.npy")
    val_data = np.load(path + "dev_data.npy")
    test_data = np.load(path + "test_data.npy")
    test_label = np.load(path + "test_label.npy")
    return train_data, train_label, val_data, val_label, test_data, test_label


def standardize_dataset(train_data, val_data, test_data, group=None, distributed=False):
    """StandardScaler over train+dev+test rows; with ``distributed=True`` the rows given are this
    rank's shard and the statistics are all-reduced over the process group (NCCL)."""
    blocks = [torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()
              for a in (train_data, val_data, test_data)]
    st = Standardizer(blocks[0].shape[1], group=group, distributed=distributed).fit(blocks)
    return tuple(st.transform(b).cpu().numpy() for b in blocks)


def standardize_dataset_with_noisy_test(train_data, val_data, test_data, sigma=0, p=0, alpha=0, group=None, distributed=False):
    """The MFCC-domain black-box sweep step in one pass (reference :433-491): ``add_white_noise_on_dataset(test, sigma)``
    (or ``add_noise_mixture_on_dataset(test, p, alpha)``) followed by ``standardize_dataset(train, val, noisy_test)``.
    The noise is drawn from numpy's global RNG in the reference's order and mixed inside the statistics and apply kernels
    (float64, two roundings): the noisy test matrix is never written.  Returns the three standardised arrays."""
    from ..frontend import Noise
    blocks = [torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda() for a in (train_data, val_data, test_data)]
    n, d = blocks[2].shape
    noise = None
    if sigma != 0:
        z = torch.from_numpy(np.random.standard_normal((n, d))).cuda()
        noise = Noise.rows_white(z, float(sigma))
    elif p != 0 and alpha != 0:
        qg = np.random.standard_normal((n, 2, d))
        noise = Noise.rows_mixture(torch.from_numpy(np.ascontiguousarray(qg[:, 0])).cuda(),
                                   torch.from_numpy(np.ascontiguousarray(qg[:, 1])).cuda(), p, alpha)
    noises = [None, None, noise]
    st = Standardizer(d, group=group, distributed=distributed).fit(blocks, noises=noises)
    return tuple(st.transform(b, noise=nz).cpu().numpy() for b, nz in zip(blocks, noises))


def _f64_host(x):
    return np.ascontiguousarray(np.asarray(x), dtype=np.float64)


def _as_audio(x):
    """int16 / float32 / float64 stay as they are; anything else goes through float64 like numpy would."""
    x = np.asarray(x)
    return np.ascontiguousarray(x if x.dtype in (np.int16, np.float32, np.float64) else x.astype(np.float64))


# Type one black-box attack
def add_white_noise(array, sigma):
    array = _as_audio(array)
    z = np.random.standard_normal(array.shape[0])
    batch = ClipBatch.from_arrays([array])
    zt = ClipBatch.from_arrays([z]).audio
    sig = torch.tensor([float(sigma)], dtype=torch.float64, device=zt.device)
    return batch.unpack(mix_white(batch, zt, sig))[0]


def mixtgauss(N, p, sigma0, sigma1):
    q = np.random.normal(0, 1, N)
    g = np.random.normal(0, 1, N)
    zero = ClipBatch.from_arrays([np.zeros(N, dtype=np.float64)])
    qd, gd = ClipBatch.from_arrays([q]).audio, ClipBatch.from_arrays([g]).audio
    from .._lib import lib, check
    out = torch.zeros_like(zero.audio)
    check(lib.asr_mix_mixture(zero.audio.data_ptr(), zero.dtype_code, zero.offsets.data_ptr(), zero.lengths.data_ptr(),
                              1, qd.data_ptr(), gd.data_ptr(), float(p), float(sigma0), float(sigma1), out.data_ptr(),
                              torch.cuda.current_stream().cuda_stream), "asr_mix_mixture")
    return zero.unpack(out)[0]


def add_noise(x, p, alpha):
    x = _as_audio(x)
    N = x.shape[0]
    q = np.random.normal(0, 1, N)
    g = np.random.normal(0, 1, N)
    batch = ClipBatch.from_arrays([x])
    qd, gd = ClipBatch.from_arrays([q]).audio, ClipBatch.from_arrays([g]).audio
    return batch.unpack(mix_mixture(batch, qd, gd, p, alpha))[0]


def add_white_noise_on_dataset(dataset, sigma):
    """Feature-domain white noise, row by row (reference :186-201): one launch for the matrix."""
    x = _f64_host(dataset)
    z = np.random.standard_normal(x.shape)           # row-major draw order == the reference's per-row loop
    xd, zd = torch.from_numpy(x).cuda(), torch.from_numpy(z).cuda()
    return mix_rows_white(xd, zd, float(sigma)).cpu().numpy()


def add_noise_mixture_on_dataset(dataset, p, alpha):
    x = _f64_host(dataset)
    n, d = x.shape
    qg = np.random.standard_normal((n, 2, d))        # per row: q then g (reference :218 -> add_noise -> mixtgauss)
    q, g = np.ascontiguousarray(qg[:, 0]), np.ascontiguousarray(qg[:, 1])
    xd = torch.from_numpy(x).cuda()
    return mix_rows_mixture(xd, torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda(), p, alpha).cpu().numpy()


def add_white_noise_with_snr(audio, target_snr_db):
    sample = _as_audio(np.asanyarray(audio))
    if sample.dtype == np.float64:
        # float64 audio: the reference's chain runs in float64 on the host; only the mix is offloaded
        sigma = np.sqrt(10 ** ((10 * np.log10(np.mean(sample ** 2)) - target_snr_db) / 10))
        z = np.random.standard_normal(len(sample))
        batch = ClipBatch.from_arrays([sample])
        sig = torch.tensor([float(sigma)], dtype=torch.float64, device=batch.audio.device)
        return batch.unpack(mix_white(batch, ClipBatch.from_arrays([z]).audio, sig))[0]
    batch = ClipBatch.from_arrays([sample])
    sigma = snr_sigma_host(clip_power(batch).cpu().numpy(), target_snr_db)
    z = np.random.standard_normal(len(sample))
    sig = torch.from_numpy(sigma).to(batch.audio.device)
    return batch.unpack(mix_white(batch, ClipBatch.from_arrays([z]).audio, sig))[0]


# ---- batched waveform-in variants (one launch per dataset pass) ------------------------------------------
def _features(batch, noise, utterance_length, params=None):
    plan = efcd.get_plan(params)
    out, status = plan.mfcc(batch, out_frames=utterance_length, noise=noise, out_dtype=torch.float64)
    if int(status.max()) != 0:
        raise ValueError("a clip is too short to be framed (librosa / np.pad would raise here)")
    return out.reshape(out.shape[0], -1).cpu().numpy()


def _draw_like(batch, n_streams=1):
    """Per clip, in order, `n_streams` standard-normal draws of the clip's length (the reference's order), written at
    the AUDIO batch's own offsets: the kernels index the noise streams exactly like the audio, whatever its packing."""
    total = int(batch.audio.shape[0])
    hosts = [torch.zeros(total, dtype=torch.float64).pin_memory() for _ in range(n_streams)]
    views = [h.numpy() for h in hosts]
    for o, n in zip(batch.offsets_host, batch.lengths_host):
        for v in views:
            v[int(o):int(o) + int(n)] = np.random.standard_normal(int(n))
    return [h.to(batch.audio.device, non_blocking=True) for h in hosts]


def black_box_attack_on_waveforms(waves, sigma=0, p=0, alpha=0, utterance_length=UTTERANCE_LENGTH, params=None):
    batch = waves if isinstance(waves, ClipBatch) else ClipBatch.from_arrays([_as_audio(w) for w in waves])
    noise = None
    if sigma != 0:
        (z,) = _draw_like(batch)
        sig = torch.full((batch.n_clips,), float(sigma), dtype=torch.float64, device=z.device)
        noise = Noise.white(z, sig)
    elif (p != 0) and (alpha != 0):
        q, g = _draw_like(batch, 2)
        noise = Noise.mixture(q, g, p, alpha)
    return _features(batch, noise, utterance_length, params)


def black_box_attack_on_waveforms_snr(waves, target_snr_db, utterance_length=UTTERANCE_LENGTH, params=None):
    batch = waves if isinstance(waves, ClipBatch) else ClipBatch.from_arrays([_as_audio(w) for w in waves])
    sigma = snr_sigma_host(clip_power(batch).cpu().numpy(), target_snr_db)
    (z,) = _draw_like(batch)
    return _features(batch, Noise.white(z, torch.from_numpy(sigma).to(z.device)), utterance_length, params)


# ---- reference-named, path-based entry points ---------------------------------------------------------------
def _load_all(filenames):
    """Decode on the host, resample on the device: one packed batch for the whole file list."""
    return audio_io.load_batch(list(filenames), sr=efcd.PARAMS.sr)


def black_box_attack_on_audio(file_path, utterance_length, sigma=0, p=0, alpha=0):
    flat = black_box_attack_on_waveforms(_load_all([file_path]), sigma, p, alpha, utterance_length)
    return flat.reshape(-1, utterance_length).astype(np.float32 if sigma == 0 and (p == 0 or alpha == 0) else np.float64)


def black_box_attack_on_audio_dataset(filenames, sigma, p, alpha):
    return black_box_attack_on_waveforms(_load_all(filenames), sigma, p, alpha, UTTERANCE_LENGTH)


def black_box_attack_on_audio_snr(file_path, utterance_length, target_snr_db):
    return black_box_attack_on_waveforms_snr(_load_all([file_path]), target_snr_db, utterance_length).reshape(
        -1, utterance_length)


def black_box_attack_on_audio_dataset_snr(filenames, target_snr_db):
    return black_box_attack_on_waveforms_snr(_load_all(filenames), target_snr_db, UTTERANCE_LENGTH)
