"""MFCC parameter sets: every ``librosa.feature.mfcc`` keyword the reference fixes or BASELINE.json varies.

``REF_VDR`` / ``REF_SR`` are the reference's real parameters
(``Voice digit recogniton/extract_features_construct_dataset.py:30`` - librosa
defaults; ``Speaker recognition/extract_features_construct_dataset.py:227-228``
- ``win_length=441, n_fft=441, hop_length=220``); ``C1``..``C5`` are the
BASELINE.json benchmark configurations.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict

from ._lib import MfccParamsC

_WINDOWS = {"hann": 0, "hamming": 1}
_PAD_MODES = {"reflect": 0, "constant": 1}
_FFTFREQ = {"linspace": 0, "rfftfreq": 1}


@dataclass(frozen=True)
class MfccParams:
    sr: int = 22050
    n_fft: int = 2048
    win_length: int = 0          # 0 -> n_fft (librosa win_length=None)
    hop_length: int = 512
    window: str = "hann"
    center: bool = True
    pad_mode: str = "reflect"    # librosa 0.9 (>= 0.10: "constant")
    fftfreq_mode: str = "linspace"   # librosa 0.9 (>= 0.10: "rfftfreq"; differs for odd n_fft)
    n_mels: int = 128
    fmin: float = 0.0
    fmax: float = 0.0            # 0 -> sr/2
    n_mfcc: int = 20
    top_db: float = 80.0         # < 0 -> None
    amin: float = 1e-10
    lifter: float = 0.0
    preemph: float = 0.0
    delta_orders: int = 0
    delta_width: int = 9

    def replace(self, **kw) -> "MfccParams":
        d = asdict(self)
        d.update(kw)
        return MfccParams(**d)

    @property
    def feature_rows(self) -> int:
        return self.n_mfcc * (1 + self.delta_orders)

    def num_frames(self, length: int) -> int:
        pad = self.n_fft // 2 if self.center else 0
        if self.pad_mode == "reflect" and pad > 0 and length <= pad:
            return 0
        padded = length + 2 * pad
        return 0 if padded < self.n_fft else 1 + (padded - self.n_fft) // self.hop_length

    def to_c(self) -> MfccParamsC:
        return MfccParamsC(
            sr=self.sr, n_fft=self.n_fft, win_length=self.win_length, hop_length=self.hop_length,
            window=_WINDOWS[self.window], center=int(self.center), pad_mode=_PAD_MODES[self.pad_mode],
            fftfreq_mode=_FFTFREQ[self.fftfreq_mode], n_mels=self.n_mels, n_mfcc=self.n_mfcc,
            fmin=self.fmin, fmax=self.fmax, top_db=self.top_db, amin=self.amin, lifter=self.lifter,
            preemph=self.preemph, delta_orders=self.delta_orders, delta_width=self.delta_width)


REF_VDR = MfccParams()
REF_SR = MfccParams(n_fft=441, win_length=441, hop_length=220)
C1 = MfccParams(sr=16000, n_fft=512, win_length=400, hop_length=160, window="hamming",
                n_mels=26, n_mfcc=13, lifter=22.0)
C3 = MfccParams(sr=16000, n_fft=512, win_length=400, hop_length=160, window="hamming",
                n_mels=40, n_mfcc=20, lifter=22.0, delta_orders=2)
C5 = MfccParams(sr=16000, n_fft=1024, win_length=1024, hop_length=160, window="hamming",
                n_mels=80, n_mfcc=40, lifter=22.0, delta_orders=1)
PRESETS = {"ref_vdr": REF_VDR, "ref_sr": REF_SR, "c1": C1, "c2": C1, "c3": C3, "c4": C1, "c5": C5}
