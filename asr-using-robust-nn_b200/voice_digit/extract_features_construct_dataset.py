"""Drop-in for ``Voice digit recogniton/extract_features_construct_dataset.py`` (hot-path functions).

Same names, argument order and return layout as the reference:

* ``extract_features(file_path, utterance_length)``            reference :24-39
* ``compute_mfcc_all_files(filenames)``                        reference :144-150
* ``get_file_names_and_labels(data_dir)``                      reference :118-140

plus the batched array-in variants they delegate to (``*_waveforms``).  The MFCC of all clips of
a call is ONE launch of the fused sm_100a kernel (``asr_mfcc_batch``) instead of one
``librosa.feature.mfcc`` call per file.  Lipschitz helpers, plotting helpers and the Keras model
code of the reference file are outside the hot path and are not mirrored.
"""
from __future__ import annotations

import os
from functools import lru_cache

import numpy as np
import torch

from .. import audio_io
from ..frontend import ClipBatch, MfccPlan
from ..params import MfccParams, REF_VDR

maxim = 0
# STANDARD_UTTERANCE_LENGTH is 101 when window length is 441, it is 44 when window length = 2048 (reference :17-18)
STANDARD_UTTERANCE_LENGTH = 44
PARAMS: MfccParams = REF_VDR      # librosa.feature.mfcc(y=raw_w, sr=sampling_rate): all defaults (:30)


@lru_cache(maxsize=16)
def _plan(params: MfccParams, device: int) -> MfccPlan:
    return MfccPlan(params, device)


def get_plan(params: MfccParams = None) -> MfccPlan:
    return _plan(PARAMS if params is None else params, torch.cuda.current_device())


def extract_features_waveforms(waves, utterance_length, params: MfccParams = None, out_dtype=torch.float32):
    """MFCC of a list of decoded waveforms (or of a packed ``ClipBatch`` already on the device) -> CUDA tensor
    (N, n_mfcc, utterance_length), frames truncated / zero-padded in the feature domain exactly like reference :33-37."""
    global maxim
    plan = get_plan(params)
    batch = waves if isinstance(waves, ClipBatch) else ClipBatch.from_arrays(waves)
    maxim = max(maxim, plan.num_frames(batch.max_length))
    out, status = plan.mfcc(batch, out_frames=utterance_length, out_dtype=out_dtype)
    bad = torch.nonzero(status).flatten()
    if bad.numel():
        i = int(bad[0])
        raise ValueError(f"clip {i} (length {int(batch.lengths_host[i])}) cannot be framed: "
                         f"status {int(status[i])} (librosa / np.pad would raise here)")
    return out


def extract_features(file_path, utterance_length):
    raw_w, sampling_rate = audio_io.load(file_path, sr=PARAMS.sr, mono=True)
    return extract_features_waveforms([raw_w], utterance_length)[0].cpu().numpy()


def compute_mfcc_all_waveforms(waves, utterance_length=None, params: MfccParams = None):
    """``compute_mfcc_all_files`` on decoded waveforms: float64 (N, n_mfcc*utterance_length), row-major
    flatten of each (n_mfcc, T) block (reference :145-149)."""
    L = STANDARD_UTTERANCE_LENGTH if utterance_length is None else utterance_length
    out = extract_features_waveforms(waves, L, params, out_dtype=torch.float64)
    return out.reshape(out.shape[0], -1).cpu().numpy()


def compute_mfcc_all_files(filenames):
    # decode on the host, upload the PCM once, resample to 22 050 Hz and extract on the device
    return compute_mfcc_all_waveforms(audio_io.load_batch(list(filenames), sr=PARAMS.sr))


# class folders of the Speech Commands corpus that the reference keeps, in label order (reference :120)
DIGITS = ('zero', 'one', 'two', 'three', 'four', 'five', 'six', 'seven', 'eight', 'nine')


def _listed_classes(data_dir, classes):
    """Reference :118-140: only the folders named in the fixed class list are taken, in the order of that list;
    label = position among the folders that are present (0..9 when all ten exist); every entry of a class folder is
    a file of that class, in sorted (glob) order."""
    present = set(os.listdir(data_dir))
    filenames, labels = [], []
    i = 0
    for name in classes:
        if name not in present:
            continue
        folder = os.path.join(data_dir, name)
        entries = sorted(os.listdir(folder))
        filenames += [os.path.join(folder, f) for f in entries]
        labels += [i] * len(entries)
        i += 1
    return np.array(filenames), np.array(labels, dtype=np.int32)


def get_file_names_and_labels(data_dir):
    return _listed_classes(data_dir, DIGITS)
