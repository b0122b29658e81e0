"""Array-in restatement of the reference's dataset-level loops (one clip per call, like the reference).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

The reference's functions take file paths and call ``librosa.load`` first; file
decode + resampling is host glue outside the hot path,
so these restatements start from the decoded waveform ``raw_w``.
"""
from __future__ import annotations

import numpy as np

from . import librosa_ref as lr
from . import noise_ref as nr


# ---- Voice digit recogniton/extract_features_construct_dataset.py:24-39 -------------------------
def extract_features(raw_w, utterance_length, p: lr.MfccParams = lr.REF_VDR):
    mfcc_features = lr.mfcc(raw_w, p)
    if mfcc_features.shape[1] > utterance_length:
        mfcc_features = mfcc_features[:, 0:utterance_length]
    else:
        mfcc_features = np.pad(mfcc_features, ((0, 0), (0, utterance_length - mfcc_features.shape[1])),
                               mode='constant', constant_values=0)
    return mfcc_features


# ---- Voice digit recogniton/extract_features_construct_dataset.py:144-150 -----------------------
def compute_mfcc_all_files(waves, utterance_length=44, p: lr.MfccParams = lr.REF_VDR):
    rows = p.n_mfcc * (1 + p.delta_orders)
    flat = np.zeros((len(waves), rows * utterance_length))
    for index in range(len(waves)):
        flat[index] = extract_features(waves[index], utterance_length, p).flatten()
    return flat


# ---- Speaker recognition/extract_features_construct_dataset.py:203-233 --------------------------
def sr_trim_split(raw_w, sampling_rate):
    """Lines :211-222: drop the first second and the tail, cut into 1-s windows."""
    window_length = 1 * sampling_rate
    audio_length = int(len(raw_w) / window_length)
    raw_w = raw_w[window_length:(audio_length - 1) * window_length]
    audio_length = int(len(raw_w) / window_length)
    return [raw_w[i * window_length:(i + 1) * window_length] for i in range(audio_length)]


def load_audio_dataset_and_labels(waves, labels, p: lr.MfccParams = lr.REF_SR):
    split_audio, local_labels = [], []
    for i, raw_w in enumerate(waves):
        for win in sr_trim_split(raw_w, p.sr):
            local_labels.append(labels[i])
            split_audio.append(win)
    feats = [lr.mfcc(np.array(w, dtype=float), p) for w in split_audio]
    if not feats:
        return np.zeros((0, 0)), np.array(local_labels)
    mfcc = np.array(feats, dtype=np.float64)
    return mfcc.reshape(mfcc.shape[0], mfcc.shape[1] * mfcc.shape[2]), np.array(local_labels)


# ---- Voice digit recogniton/attacks.py:248-294 (SNR) and :89-142 (sigma / mixture) --------------
def black_box_attack_on_audio_snr(raw_w, utterance_length, target_snr_db, z, p: lr.MfccParams = lr.REF_VDR):
    noisy = nr.add_white_noise_with_snr_z(raw_w, target_snr_db, z)
    return extract_features(noisy, utterance_length, p)


def black_box_attack_on_audio_dataset_snr(waves, target_snr_db, zs, utterance_length=44,
                                          p: lr.MfccParams = lr.REF_VDR):
    flat = np.zeros((len(waves), p.n_mfcc * (1 + p.delta_orders) * utterance_length))
    for i in range(len(waves)):
        flat[i] = black_box_attack_on_audio_snr(waves[i], utterance_length, target_snr_db, zs[i], p).flatten()
    return flat


def black_box_attack_on_audio(raw_w, utterance_length, sigma=0, p_peak=0, alpha=0, z=None, q=None, g=None,
                              p: lr.MfccParams = lr.REF_VDR):
    if sigma != 0:
        raw_w = nr.add_white_noise_z(raw_w, sigma, z)
    elif (p_peak != 0) and (alpha != 0):
        raw_w = nr.add_noise_z(raw_w, p_peak, alpha, q, g)
    return extract_features(raw_w, utterance_length, p)


# ---- Speaker recognition/attacks.py:254-295 ------------------------------------------------------
def sr_black_box_attack_on_audio_snr(waves, labels, target_snr_db, zs, p: lr.MfccParams = lr.REF_SR):
    """Noise on the WHOLE file first (:273), then trim/split (:274-284), then MFCC (:287-290)."""
    noisy = [nr.add_white_noise_with_snr_z(w, target_snr_db, z) for w, z in zip(waves, zs)]
    return load_audio_dataset_and_labels(noisy, np.copy(labels), p)
