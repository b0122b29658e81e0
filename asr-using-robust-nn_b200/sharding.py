"""Clip sharding over the GPUs of one box.

Clips are independent units: rank r owns one contiguous block of the global clip index range, so
concatenating the ranks' output rows in rank order reproduces the single-GPU row order (the
``.npy`` rows of ``compute_mfcc_all_files``, VDR/extract_features_construct_dataset.py:145-149).
No data-path exchange is needed for the MFCC or the noise mix - the seeded normal stream is indexed
by GLOBAL sample position (``asr_randn_f64(seed, first_index, ...)``), so a clip's noise does not
depend on the number of ranks.  The only collective of the path is the all-reduce of the
standardisation accumulators (``frontend.Standardizer``).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np


def shard_bounds(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[lo, hi) of rank's contiguous block; sizes differ by at most one, earlier ranks get the extra item."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} for world size {world_size}")
    if n_items < 0:
        raise ValueError("negative item count")
    q, r = divmod(n_items, world_size)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_list(items: Sequence, rank: int, world_size: int):
    lo, hi = shard_bounds(len(items), rank, world_size)
    return items[lo:hi]


def first_sample_index(lengths: Sequence[int], rank: int, world_size: int, align: int = 8) -> int:
    """Global position of the shard's first sample in the packed (8-element aligned) audio layout: the
    `first_index` to hand to ``frontend.randn`` so the noise stream is independent of the sharding."""
    lengths = np.asarray(lengths, dtype=np.int64)
    lo, _ = shard_bounds(len(lengths), rank, world_size)
    padded = (lengths[:lo] + align - 1) // align * align
    return int(padded.sum())
