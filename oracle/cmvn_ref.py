"""Restatement of ``standardize_dataset`` (dataset-level standardisation, "CMVN").

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

Follows ``Voice digit recogniton/attacks.py:48-69`` (twin:
``Speaker recognition/attacks.py:56-77``; inline copies in the four training
scripts, e.g. ``Voice digit recogniton/train_constraints.py:28-35``).  The
reference calls ``sklearn.preprocessing.StandardScaler`` - sklearn IS in the
image, so this oracle calls the very same class rather than restating it.
"""
from __future__ import annotations

import numpy as np
from sklearn.preprocessing import StandardScaler


def standardize_dataset(train_data, val_data, test_data):
    all_data = np.concatenate((train_data, val_data, test_data), axis=0)
    scaler1 = StandardScaler()
    all_data = scaler1.fit_transform(all_data)
    n0 = train_data.shape[0]
    n1 = val_data.shape[0]
    return all_data[:n0], all_data[n0:n0 + n1], all_data[n0 + n1:]


def column_stats(all_data):
    """(mean_, var_, scale_) of ``StandardScaler().fit(all_data)`` (float64)."""
    s = StandardScaler().fit(np.asarray(all_data, dtype=np.float64))
    return s.mean_, s.var_, s.scale_
