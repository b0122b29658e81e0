"""numpy/scipy restatement of ``librosa.feature.mfcc`` (librosa 0.9 semantics).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  PARITY UNPINNED: librosa is
not in the image; this file restates its published algorithm and follows the
dtype flow librosa has for float32 / float64 input.  What IS pinned to librosa's
own published known answers: the Slaney mel scale (``hz_to_mel``, ``mel_to_hz``,
the 40 values of ``mel_frequencies(n_mels=40)``) and the corner values of
``librosa.filters.mel(sr=22050, n_fft=2048)`` - the docstring examples of librosa
0.9, ``tests/test_oracle.py::test_oracle_reproduces_librosa_docstring_examples``.

Reference call sites this restates the callee of:
  * ``Voice digit recogniton/extract_features_construct_dataset.py:30``
      ``librosa.feature.mfcc(y=raw_w, sr=sampling_rate)``            (all defaults)
  * ``Speaker recognition/extract_features_construct_dataset.py:227-228``
      ``librosa.feature.mfcc(y=float64 window, sr, win_length=441, n_fft=441, hop_length=220)``
  * ``Voice digit recogniton/attacks.py:114,267``; ``Speaker recognition/attacks.py:140-141,289-290``

Every keyword BASELINE.json's configs vary (window, n_fft, win_length, hop,
n_mels, n_mfcc, lifter) is a ``librosa.feature.mfcc`` keyword and goes through
the same code.  Pre-emphasis (``librosa.effects.preemphasis``) and deltas
(``librosa.feature.delta``) are never called by the reference; they are
restated here from librosa's documented behaviour for the BASELINE configs
that ask for them.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
import numpy as np
import scipy.fftpack
import scipy.signal


@dataclass(frozen=True)
class MfccParams:
    """Every ``librosa.feature.mfcc`` / ``melspectrogram`` / ``stft`` keyword the path uses."""
    sr: int = 22050
    n_fft: int = 2048
    win_length: int = 0          # 0 -> n_fft (librosa: win_length=None)
    hop_length: int = 512
    window: str = "hann"         # scipy.signal.get_window name ("hann", "hamming")
    center: bool = True
    pad_mode: str = "reflect"    # librosa 0.9 stft default; >=0.10 uses "constant"
    fftfreq_mode: str = "linspace"   # librosa 0.9 fft_frequencies; >=0.10 "rfftfreq"
    n_mels: int = 128
    fmin: float = 0.0
    fmax: float = 0.0            # 0 -> sr/2
    n_mfcc: int = 20
    top_db: float = 80.0         # <0 -> None
    amin: float = 1e-10
    lifter: float = 0.0
    preemph: float = 0.0         # 0 -> no pre-emphasis
    delta_orders: int = 0        # 0, 1 (append delta) or 2 (append delta, delta-delta)
    delta_width: int = 9

    def replace(self, **kw):
        d = asdict(self)
        d.update(kw)
        return MfccParams(**d)


# ---- presets: the reference's two real parameter sets and BASELINE.json's configs ------------
REF_VDR = MfccParams()                                   # VDR/extract...py:30 (all librosa defaults)
REF_SR = MfccParams(n_fft=441, win_length=441, hop_length=220)   # SR/extract...py:227-228
C1 = MfccParams(sr=16000, n_fft=512, win_length=400, hop_length=160, window="hamming",
                n_mels=26, n_mfcc=13, lifter=22.0)
C3 = MfccParams(sr=16000, n_fft=512, win_length=400, hop_length=160, window="hamming",
                n_mels=40, n_mfcc=20, lifter=22.0, delta_orders=2)
C5 = MfccParams(sr=16000, n_fft=1024, win_length=1024, hop_length=160, window="hamming",
                n_mels=80, n_mfcc=40, lifter=22.0, delta_orders=1)
PRESETS = {"ref_vdr": REF_VDR, "ref_sr": REF_SR, "c1": C1, "c2": C1, "c3": C3, "c4": C1, "c5": C5}


# ---- librosa.core.convert ----------------------------------------------------------------------
def hz_to_mel(f):
    """Slaney mel scale (librosa ``hz_to_mel(htk=False)``)."""
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        m = f >= min_log_hz
        mels[m] = min_log_mel + np.log(f[m] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if m.ndim:
        g = m >= min_log_mel
        freqs[g] = min_log_hz * np.exp(logstep * (m[g] - min_log_mel))
    elif m >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (m - min_log_mel))
    return freqs


def fft_frequencies(sr, n_fft, mode="linspace"):
    if mode == "linspace":       # librosa <= 0.9
        return np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)
    if mode == "rfftfreq":       # librosa >= 0.10
        return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    raise ValueError(mode)


def mel_filterbank(p: MfccParams) -> np.ndarray:
    """``librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm='slaney')`` -> float32."""
    fmax = p.fmax if p.fmax > 0 else float(p.sr) / 2
    n_bins = int(1 + p.n_fft // 2)
    weights = np.zeros((p.n_mels, n_bins), dtype=np.float32)
    fftfreqs = fft_frequencies(p.sr, p.n_fft, p.fftfreq_mode)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(p.fmin), hz_to_mel(fmax), p.n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(p.n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:p.n_mels + 2] - mel_f[:p.n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


# ---- librosa.core.spectrum ---------------------------------------------------------------------
def fft_window(p: MfccParams) -> np.ndarray:
    """``get_window(window, win_length, fftbins=True)`` centre-padded to n_fft (float64)."""
    win_length = p.win_length if p.win_length > 0 else p.n_fft
    w = scipy.signal.get_window(p.window, win_length, fftbins=True)
    lpad = (p.n_fft - win_length) // 2
    return np.pad(w, (lpad, p.n_fft - win_length - lpad))


def num_frames(p: MfccParams, length: int) -> int:
    if p.center:
        length = length + 2 * (p.n_fft // 2)
    if length < p.n_fft:
        return 0
    return 1 + (length - p.n_fft) // p.hop_length


def preemphasis(y: np.ndarray, coef: float) -> np.ndarray:
    """``librosa.effects.preemphasis(y, coef=coef)``: ``lfilter([1,-coef],[1],y, zi=2*y[0]-y[1])``.

    The filter STATE (not the previous sample) is initialised with the linear
    extrapolation ``2*y[0]-y[1]``, so ``out[0] = y[0] + (2*y[0]-y[1])``; computed
    in ``y.dtype``.  Not on the reference's path (parity unpinned).
    """
    y = np.asarray(y)
    dt = y.dtype if y.dtype.kind == "f" else np.dtype(np.float32)
    y = y.astype(dt, copy=False)
    b = np.asarray([1.0, -coef], dtype=dt)
    a = np.asarray([1.0], dtype=dt)
    zi = np.atleast_1d(2 * y[0:1] - y[1:2]).astype(dt)
    out, _ = scipy.signal.lfilter(b, a, y, zi=zi)
    return out.astype(dt, copy=False)


def power_spectrogram(y: np.ndarray, p: MfccParams) -> np.ndarray:
    """``np.abs(librosa.stft(y, ...))**2`` with librosa's dtype flow -> (1+n_fft//2, T)."""
    y = np.asarray(y)
    if y.dtype.kind != "f":
        raise TypeError("audio must be floating point (librosa.util.valid_audio)")
    w = fft_window(p).reshape(-1, 1)
    if p.center:
        pad = p.n_fft // 2
        if p.pad_mode == "reflect" and y.shape[0] <= pad:
            raise ValueError("reflect padding needs len(y) > n_fft//2")
        y = np.pad(y, pad, mode=p.pad_mode)
    T = num_frames(p.replace(center=False), y.shape[0])
    if T <= 0:
        raise ValueError("input too short for one frame")
    idx = np.arange(p.n_fft)[:, None] + p.hop_length * np.arange(T)[None, :]
    frames = y[idx]                                   # (n_fft, T), y.dtype
    cdtype = np.complex64 if y.dtype == np.float32 else np.complex128
    D = np.fft.rfft(w * frames, axis=0).astype(cdtype)   # FFT in double, stored in r2c(y.dtype)
    return np.abs(D) ** 2


def power_to_db(S: np.ndarray, amin: float, top_db: float) -> np.ndarray:
    log_spec = 10.0 * np.log10(np.maximum(amin, S))
    log_spec -= 10.0 * np.log10(np.maximum(amin, 1.0))
    if top_db is not None and top_db >= 0:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def delta(data: np.ndarray, width: int, order: int) -> np.ndarray:
    """``librosa.feature.delta(data, width, order, axis=-1, mode='interp')``."""
    if data.shape[-1] < width:
        raise ValueError("delta needs at least `width` frames")
    return scipy.signal.savgol_filter(data, width, deriv=order, polyorder=order, axis=-1, mode="interp")


def mfcc(y: np.ndarray, p: MfccParams) -> np.ndarray:
    """``librosa.feature.mfcc(y, sr, n_mfcc, dct_type=2, norm='ortho', lifter, **mel kwargs)``.

    Returns (n_mfcc * (1 + delta_orders), T) in float32 for float32 input and
    float64 for float64 input (librosa's dtype flow).
    """
    y = np.asarray(y)
    if p.preemph != 0.0:
        y = preemphasis(y, p.preemph)
    S = power_spectrogram(y, p)
    M = mel_filterbank(p) @ S                      # float32 @ float32 -> float32 ; @ float64 -> float64
    L = power_to_db(M, p.amin, p.top_db)
    C = scipy.fftpack.dct(L, axis=-2, type=2, norm="ortho")[:p.n_mfcc, :]
    if p.lifter > 0:
        LI = np.sin(np.pi * np.arange(1, 1 + p.n_mfcc, dtype=C.dtype) / p.lifter)
        C = C * (1 + (p.lifter / 2) * LI)[:, None]
        C = C.astype(L.dtype, copy=False)
    feats = [C]
    for order in range(1, p.delta_orders + 1):
        feats.append(delta(C, p.delta_width, order).astype(C.dtype, copy=False))
    return np.concatenate(feats, axis=0) if len(feats) > 1 else C


def log_mel(y: np.ndarray, p: MfccParams) -> np.ndarray:
    """Intermediate (n_mels, T) dB matrix, exposed for stage-level parity tests."""
    y = np.asarray(y)
    if p.preemph != 0.0:
        y = preemphasis(y, p.preemph)
    return power_to_db(mel_filterbank(p) @ power_spectrogram(y, p), p.amin, p.top_db)
