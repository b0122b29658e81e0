"""On-disk format of the processed datasets and the dataset-build driver.

The Keras training and attack scripts of the reference read six ``.npy`` files per task
(``load_npy_dataset``, VDR/attacks.py:27-45 ; VDR/train_constraints.py:17-26):
``{train,dev,test}_data.npy`` float64 C-order ``(N, n_mfcc*T)`` and ``{train,dev,test}_label.npy``,
plus ``test_dataset_to_add_noise/test_{label,filenames}.npy`` (VDR/extract...py:219-220).  This module
writes exactly those files from the drop-in extractors, so the downstream scripts run unchanged, and
merges per-rank shards by clip index when the extraction ran on several GPUs.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Sequence

import numpy as np

SPLIT_NAMES = ("train", "dev", "test")


def split_70_20_10(filenames, labels):
    """The reference's slicing (VDR/extract...py:210-216): 70 % / next 20 % / LAST 10 %."""
    n = int(len(filenames))
    a, b, c = int(n * 0.7), int(n * 0.9), int(n * 0.1)
    files = (filenames[:a], filenames[a:b], filenames[n - c:] if c else filenames[:0])
    labs = (labels[:a], labels[a:b], labels[n - c:] if c else labels[:0])
    return files, labs


def as_rows(features) -> np.ndarray:
    """float64, C-order, 2-D: what ``np.save`` of the reference's ``mfcc_*`` arrays holds."""
    a = np.ascontiguousarray(np.asarray(features), dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("feature matrix must be (N, n_mfcc*T)")
    return a


def save_processed_dataset(save_dir: str, data: Sequence, labels: Sequence, noise_dir: Optional[str] = None,
                           test_filenames=None) -> None:
    """Write ``{train,dev,test}_{data,label}.npy`` (VDR/extract...py:227-232) and, if ``noise_dir`` is given,
    ``test_label.npy`` / ``test_filenames.npy`` there (:219-220)."""
    os.makedirs(save_dir, exist_ok=True)
    for name, x, y in zip(SPLIT_NAMES, data, labels):
        np.save(os.path.join(save_dir, f"{name}_data"), as_rows(x))
        np.save(os.path.join(save_dir, f"{name}_label"), np.asarray(y))
    if noise_dir is not None:
        os.makedirs(noise_dir, exist_ok=True)
        np.save(os.path.join(noise_dir, "test_label"), np.asarray(labels[2]))
        if test_filenames is not None:
            np.save(os.path.join(noise_dir, "test_filenames"), np.asarray(test_filenames))


def load_npy_dataset(path: str):
    """``load_npy_dataset`` of the reference (VDR/attacks.py:27-45): ``path`` is a prefix ending in a separator."""
    j = lambda n: np.load(os.path.join(path, n) if not path.endswith(("/", "\\\\")) else path + n, allow_pickle=True)
    return (j("train_data.npy"), j("train_label.npy"), j("dev_data.npy"), j("dev_label.npy"), j("test_data.npy"),
            j("test_label.npy"))


def merge_shards(shards: Sequence[np.ndarray]) -> np.ndarray:
    """Rows of rank 0, rank 1, ... concatenated: with ``sharding.shard_bounds`` (contiguous blocks) this is the
    single-GPU row order."""
    shards = [as_rows(s) for s in shards if len(s)]
    if not shards:
        return np.zeros((0, 0))
    return np.concatenate(shards, axis=0)


def build_dataset(filenames, labels, compute: Callable, save_dir: str, noise_dir: Optional[str] = None,
                  shuffle_seed: Optional[int] = None):
    """The reference's ``__main__`` (VDR/extract...py:199-232): shuffle, split 70/20/10, extract, save.

    ``compute(filenames) -> (N, D)`` is a drop-in ``compute_mfcc_all_files``.  The reference shuffles with an
    unseeded ``sklearn.utils.shuffle``; pass ``shuffle_seed`` for a reproducible run, ``None`` keeps the order."""
    filenames, labels = np.asarray(filenames), np.asarray(labels)
    if shuffle_seed is not None:
        perm = np.random.RandomState(shuffle_seed).permutation(len(filenames))
        filenames, labels = filenames[perm], labels[perm]
    files, labs = split_70_20_10(filenames, labels)
    data = [compute(list(f)) if len(f) else np.zeros((0, 0)) for f in files]
    save_processed_dataset(save_dir, data, labs, noise_dir, files[2])
    return data, labs
