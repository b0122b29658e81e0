"""Drop-in for ``Speaker recognition/extract_features_construct_dataset.py`` (hot-path functions).

* ``load_audio_dataset_and_labels(filenames, labels)``   reference :203-233
* ``extract_features(file_path, utterance_length)``      reference :21-35
* ``compute_mfcc_all_files(filenames)``                  reference :141-147 (L = 500)
* ``get_file_names_and_labels(data_dir)``

Every file is trimmed (first second and the tail dropped, :211-214) and cut into non-overlapping
1-s windows (:216-222); each window's MFCC uses ``win_length=441, n_fft=441, hop_length=220``
(:227-228) -> (20, 101) -> row of 2020.  All windows of a call go through ONE fused launch; the
windows are index ranges into the packed file audio (no per-window copies).
"""
from __future__ import annotations

import os
from functools import lru_cache

import numpy as np
import torch

from .. import audio_io
from ..frontend import ClipBatch, MfccPlan
from ..params import MfccParams, REF_SR, REF_VDR

maximum = 0
STANDARD_UTTERANCE_LENGTH = 500        # reference :15 (used by compute_mfcc_all_files only)
PARAMS: MfccParams = REF_SR            # per-window MFCC (:227-228)
PARAMS_FILE: MfccParams = REF_VDR      # extract_features uses librosa defaults (:28)


@lru_cache(maxsize=16)
def _plan(params: MfccParams, device: int) -> MfccPlan:
    return MfccPlan(params, device)


def get_plan(params: MfccParams = None) -> MfccPlan:
    return _plan(PARAMS if params is None else params, torch.cuda.current_device())


def window_index(lengths, sampling_rate):
    """Trim/split bookkeeping of reference :209-222 for files of the given lengths.

    Returns (file_id, start) per 1-s window: window j of file i covers
    ``raw_w[start : start + sampling_rate]`` of the ORIGINAL (untrimmed) file.
    """
    file_id, start = [], []
    window_length = 1 * sampling_rate
    for i, n in enumerate(lengths):
        audio_length = int(n / window_length)
        lo, hi = window_length, (audio_length - 1) * window_length     # raw_w[window_length:(audio_length-1)*window_length]
        kept = max(0, min(hi, n) - lo) if hi > lo else 0
        for index in range(int(kept / window_length)):
            file_id.append(i)
            start.append(lo + index * window_length)
    return np.asarray(file_id, dtype=np.int64), np.asarray(start, dtype=np.int64)


def windows_of(batch: ClipBatch, sampling_rate) -> tuple:
    """ClipBatch whose clips are the 1-s windows inside `batch`'s packed audio (zero copy) + file ids."""
    fid, start = window_index(batch.lengths_host, sampling_rate)
    offsets = batch.offsets_host[fid] + start if len(fid) else np.zeros(0, dtype=np.int64)
    lengths = np.full(len(fid), sampling_rate, dtype=np.int32)
    dev = batch.audio.device
    wb = ClipBatch(batch.audio, torch.from_numpy(offsets).to(dev), torch.from_numpy(lengths).to(dev),
                   int(sampling_rate) if len(fid) else 0, offsets, lengths)
    return wb, fid


def mfcc_windows(batch: ClipBatch, noise=None, params: MfccParams = None):
    """(n_windows, n_mfcc*T) float64 features of every 1-s window of the files in `batch`."""
    prm = PARAMS if params is None else params
    plan = get_plan(prm)
    wb, fid = windows_of(batch, prm.sr)
    if wb.n_clips == 0:
        return np.zeros((0, 0)), fid
    out, status = plan.mfcc(wb, noise=noise, out_dtype=torch.float64)
    if int(status.max()) != 0:
        raise ValueError("a window is too short to be framed")
    return out.reshape(out.shape[0], -1).cpu().numpy(), fid


def load_waveforms_and_labels(waves, labels, params: MfccParams = None):
    """``load_audio_dataset_and_labels`` on decoded waveforms (a list of arrays, or a packed ``ClipBatch``)."""
    batch = waves if isinstance(waves, ClipBatch) else ClipBatch.from_arrays([np.ascontiguousarray(w) for w in waves])
    mfcc, fid = mfcc_windows(batch, None, params)
    return mfcc, np.asarray(labels)[fid]


def load_audio_dataset_and_labels(filenames, labels):
    # decode on the host, upload the PCM once, resample to 22 050 Hz on the device; windows are index ranges
    return load_waveforms_and_labels(audio_io.load_batch(list(filenames), sr=PARAMS.sr), labels)


def extract_features(file_path, utterance_length):
    global maximum
    raw_w, _ = audio_io.load(file_path, sr=PARAMS_FILE.sr, mono=True)
    plan = get_plan(PARAMS_FILE)
    batch = ClipBatch.from_arrays([raw_w])
    maximum = max(maximum, plan.num_frames(batch.max_length))
    out, status = plan.mfcc(batch, out_frames=utterance_length)
    if int(status.max()) != 0:
        raise ValueError("clip too short to be framed")
    return out[0].cpu().numpy()


def compute_mfcc_all_files(filenames):
    waves = [audio_io.load(f, sr=PARAMS_FILE.sr, mono=True)[0] for f in filenames]
    plan = get_plan(PARAMS_FILE)
    out, status = plan.mfcc(ClipBatch.from_arrays(waves), out_frames=STANDARD_UTTERANCE_LENGTH, out_dtype=torch.float64)
    if int(status.max()) != 0:
        raise ValueError("clip too short to be framed")
    return out.reshape(out.shape[0], -1).cpu().numpy()


# the 20 RoDigits speaker folders the reference keeps, in label order (SR reference :116-117)
SPEAKERS = ('006', '041', '043', '044', '045', '046', '047', '048', '049', '105', '117', '118', '211', '212',
            '213', '214', '215', '260', '261', '420')


def get_file_names_and_labels(data_dir):
    """SR reference :113-137: the same whitelist-in-list-order rule as the digit corpus."""
    from ..voice_digit.extract_features_construct_dataset import _listed_classes
    return _listed_classes(data_dir, SPEAKERS)
