"""Restatement of the reference's additive-noise code (pure numpy, behaviour verbatim).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

Each function follows the reference text executed under THIS image's numpy
(2.3.x, NEP-50 scalar promotion: every scalar in the sigma chain of
``add_white_noise_with_snr`` is float32 when the audio is float32).

The ``*_z`` variants take the standard-normal stream ``z`` as an argument
instead of drawing it from numpy's global RNG.  They are bit-identical to the
seeded originals because ``np.random.normal(0, s, n) == float(s) *
np.random.standard_normal(n)`` bit for bit (checked in tests/test_oracle.py).
"""
from __future__ import annotations

import numpy as np


# ---- Voice digit recogniton/attacks.py:73-86 ; Speaker recognition/attacks.py:81-94 ----------
def add_white_noise(array, sigma):
    noise = np.random.normal(0, sigma, np.array(array).shape[0])
    return array + noise


def add_white_noise_z(array, sigma, z):
    """Same as ``add_white_noise`` with ``noise = float(sigma) * z``."""
    noise = float(sigma) * np.asarray(z, dtype=np.float64)
    return array + noise


# ---- Voice digit recogniton/attacks.py:145-183 ; Speaker recognition/attacks.py:149-189 ------
def mixtgauss(N, p, sigma0, sigma1):
    q = np.random.normal(0, 1, N)
    u = np.abs(q) < p
    return (sigma0 * (1 - u) + sigma1 * u) * np.random.normal(0, 1, N)


def add_noise(x, p, alpha):
    N = x.shape[0]
    noise = mixtgauss(N, p, alpha, 10 * alpha)
    return x + noise


def add_noise_z(x, p, alpha, q, g):
    """``add_noise`` with the two normal streams given: ``q`` (selector) then ``g`` (carrier)."""
    q = np.asarray(q, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64)
    sigma0 = alpha
    sigma1 = 10 * alpha
    u = np.abs(q) < p
    noise = (sigma0 * (1 - u) + sigma1 * u) * g
    return x + noise


# ---- Voice digit recogniton/attacks.py:222-245 ; Speaker recognition/attacks.py:228-251 ------
def snr_sigma(audio, target_snr_db):
    """The sigma chain of ``add_white_noise_with_snr`` (lines :233-238 and the ``np.sqrt`` in :241)."""
    sample = np.asanyarray(audio)
    signal_avg_watts = np.mean(sample ** 2)
    signal_avg_db = 10 * np.log10(signal_avg_watts)
    noise_avg_db = signal_avg_db - target_snr_db
    noise_avg_watts = 10 ** (noise_avg_db / 10)
    return np.sqrt(noise_avg_watts)


def snr_sigma_from_power(signal_avg_watts, target_snr_db):
    """Same chain starting from ``P = np.mean(sample**2)`` (a numpy scalar of the audio's dtype)."""
    signal_avg_db = 10 * np.log10(signal_avg_watts)
    noise_avg_db = signal_avg_db - target_snr_db
    noise_avg_watts = 10 ** (noise_avg_db / 10)
    return np.sqrt(noise_avg_watts)


def add_white_noise_with_snr(audio, target_snr_db):
    sample = np.asanyarray(audio)
    sigma = snr_sigma(sample, target_snr_db)
    noise_volts = 1 * np.random.normal(0, sigma, len(sample))
    return sample + noise_volts


def add_white_noise_with_snr_z(audio, target_snr_db, z):
    sample = np.asanyarray(audio)
    sigma = snr_sigma(sample, target_snr_db)
    noise_volts = 1 * (float(sigma) * np.asarray(z, dtype=np.float64))
    return sample + noise_volts


# ---- babble noise: BASELINE.json configs[1] ("white/babble"); NO reference implementation - PARITY UNPINNED ----
# Recipe of SURVEY.md 8(d): the noise of clip i is the sum of 6 other clips of the batch, indices (i + k*97) mod B,
# scaled by the sigma law of add_white_noise_with_snr so that noise power = P / 10^(snr/10).
def babble_stream(clips, i, stride=97, talkers=6):
    """float64 sum of the `talkers` other clips at every sample of clip i (a shorter talker contributes nothing past its end)."""
    n = len(clips[i])
    b = np.zeros(n, dtype=np.float64)
    for k in range(1, talkers + 1):
        c = np.asarray(clips[(i + k * stride) % len(clips)], dtype=np.float64)
        m = min(n, len(c))
        b[:m] += c[:m]
    return b


def add_babble_with_snr(clips, i, target_snr_db, stride=97, talkers=6, gain=None):
    """clip i + gain * babble, gain = sigma(P_i, snr) / sqrt(mean(babble^2)); `gain` may be given (bit-exact mix test)."""
    sample = np.asanyarray(clips[i])
    b = babble_stream(clips, i, stride, talkers)
    if gain is None:
        sigma = snr_sigma(sample, target_snr_db)
        pb = np.mean(b ** 2)
        gain = float(sigma) / np.sqrt(pb) if pb > 0 else 0.0
    return sample + float(gain) * b


# ---- numpy's float32 pairwise summation, restated (used to explain / pin the device order) ---
def pairwise_sum_f32(a: np.ndarray) -> np.float32:
    """Python replica of numpy's ``pairwise_sum`` for a contiguous float32 vector.

    n < 8: serial from 0; n <= 128: 8 strided accumulators combined as
    ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then the tail serially; else split at
    n2 = n/2 - (n/2 % 8) and recurse.  Slow (pure Python) - small cases only.
    """
    a = np.asarray(a, dtype=np.float32)
    n = a.shape[0]
    f = np.float32
    if n < 8:
        res = f(0.0)
        for i in range(n):
            res = f(res + a[i])
        return res
    if n <= 128:
        r = [f(a[j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = f(r[j] + a[i + j])
            i += 8
        res = f(f(f(r[0] + r[1]) + f(r[2] + r[3])) + f(f(r[4] + r[5]) + f(r[6] + r[7])))
        while i < n:
            res = f(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return f(pairwise_sum_f32(a[:n2]) + pairwise_sum_f32(a[n2:]))


def mean_power_f32(audio_f32: np.ndarray) -> np.float32:
    """``np.mean(sample**2)`` for float32 audio via the replica above (cross-check of the order)."""
    a = np.asarray(audio_f32, dtype=np.float32)
    s = pairwise_sum_f32(a * a)
    return np.float32(s / np.float32(a.shape[0]))
