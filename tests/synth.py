"""Seeded synthetic audio shared by tests, golden generation and bench.py."""
import numpy as np


def synth_clips(n_clips, length, sr, seed, lengths=None):
    """List of int16 clips: 0.1*N(0,1)*hann envelope + 0.05*sin(2 pi f0 n / sr), f0 ~ U(100, 4000)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_clips):
        L = int(length if lengths is None else lengths[i])
        n = np.arange(L)
        env = np.hanning(L) if L > 1 else np.ones(L)
        f0 = rng.uniform(100.0, 4000.0)
        x = 0.1 * rng.standard_normal(L) * env + 0.05 * np.sin(2 * np.pi * f0 * n / sr)
        out.append(np.clip(np.round(x * 32767.0), -32768, 32767).astype(np.int16))
    return out


def to_f32(clips_i16):
    return [(c.astype(np.float32) / np.float32(32768.0)) for c in clips_i16]
