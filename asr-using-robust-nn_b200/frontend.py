"""Batched, array-in host API over the C-ABI (device memory and streams through torch).

This is the layer the reference-named entry points in ``voice_digit/`` and
``speaker/`` delegate to after decoding audio: waveforms + lengths in, the
reference's ``(N, n_mfcc*T)`` feature layout out.  torch is plumbing only
(allocations, streams, ``torch.distributed``); all arithmetic runs in
``libasr_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, AsrError, NoiseC
from .params import MfccParams

_DT = {torch.int16: _lib.ASR_I16, torch.float32: _lib.ASR_F32, torch.float64: _lib.ASR_F64}
_NP2T = {np.dtype(np.int16): torch.int16, np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}


_MATRIX_LAYOUTS: dict = {}


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise AsrError("no CUDA device: the asr_b200 path has no CPU fallback")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


@dataclass
class ClipBatch:
    """Packed clips on the device: ``audio[offsets[b] : offsets[b]+lengths[b]]`` is clip b."""
    audio: torch.Tensor        # 1-D int16 / float32 / float64, device
    offsets: torch.Tensor      # int64 [B], device
    lengths: torch.Tensor      # int32 [B], device
    max_length: int
    offsets_host: np.ndarray   # int64 [B]
    lengths_host: np.ndarray   # int32 [B]

    @property
    def n_clips(self) -> int:
        return int(self.lengths_host.shape[0])

    @property
    def dtype_code(self) -> int:
        return _DT[self.audio.dtype]

    @staticmethod
    def layout(lengths: Sequence[int], align: int = 8):
        """Offsets for packing clips back to back, each start aligned to `align` elements."""
        lengths = np.asarray(lengths, dtype=np.int64)
        padded = (lengths + align - 1) // align * align
        offsets = np.zeros(len(lengths), dtype=np.int64)
        if len(lengths) > 1:
            offsets[1:] = np.cumsum(padded[:-1])
        total = int(padded.sum()) if len(lengths) else 0
        return offsets, max(total, align)

    @staticmethod
    def from_arrays(clips: Sequence[np.ndarray], device="cuda", dtype=None) -> "ClipBatch":
        """Pack a list of 1-D numpy waveforms (int16 / float32 / float64, all the same dtype)."""
        _require_cuda()
        clips = [np.ascontiguousarray(c) for c in clips]
        if dtype is None:
            dtype = clips[0].dtype if clips else np.dtype(np.float32)
        dtype = np.dtype(dtype)
        if dtype not in _NP2T:
            raise TypeError(f"unsupported audio dtype {dtype}; use int16, float32 or float64")
        lengths = np.array([c.shape[0] for c in clips], dtype=np.int32)
        offsets, total = ClipBatch.layout(lengths)
        host = torch.zeros(total, dtype=_NP2T[dtype]).pin_memory()
        hv = host.numpy()
        for c, o in zip(clips, offsets):
            if c.ndim != 1:
                raise ValueError("clips must be 1-D (mono) waveforms")
            hv[o:o + c.shape[0]] = c.astype(dtype, copy=False)
        return ClipBatch(host.to(device, non_blocking=True), torch.from_numpy(offsets).to(device),
                         torch.from_numpy(lengths).to(device), int(lengths.max()) if len(lengths) else 0,
                         offsets, lengths)

    @staticmethod
    def from_matrix(x: torch.Tensor) -> "ClipBatch":
        """Equal-length clips already on the device as a contiguous (B, L) tensor (no copy)."""
        if x.dim() != 2 or not x.is_contiguous() or x.dtype not in _DT:
            raise ValueError("expected a contiguous (B, L) int16/float32/float64 tensor")
        B, L = x.shape
        key = (B, L, x.device)
        hit = _MATRIX_LAYOUTS.get(key)
        if hit is None:
            offsets = np.arange(B, dtype=np.int64) * L
            lengths = np.full(B, L, dtype=np.int32)
            hit = (torch.from_numpy(offsets).to(x.device), torch.from_numpy(lengths).to(x.device), offsets, lengths)
            if len(_MATRIX_LAYOUTS) > 64:
                _MATRIX_LAYOUTS.clear()
            _MATRIX_LAYOUTS[key] = hit
        return ClipBatch(x.reshape(-1), hit[0], hit[1], L, hit[2], hit[3])

    def like(self, data: torch.Tensor) -> "ClipBatch":
        """Same layout, different payload (e.g. the float64 noisy signal of a mix)."""
        return ClipBatch(data, self.offsets, self.lengths, self.max_length, self.offsets_host, self.lengths_host)

    def unpack(self, data: Optional[torch.Tensor] = None):
        """List of per-clip numpy arrays (of `data`, packed like the audio; default the audio itself)."""
        h = (self.audio if data is None else data).cpu().numpy()
        return [h[o:o + n].copy() for o, n in zip(self.offsets_host, self.lengths_host)]


@dataclass
class Noise:
    """Additive noise fused in front of the MFCC (see ``asr_noise`` in include/asr_b200.h)."""
    mode: int
    z: Optional[torch.Tensor] = None        # white: z ; mixture: selector q   (float64, packed like the audio)
    z2: Optional[torch.Tensor] = None       # mixture: carrier g
    sigma: Optional[torch.Tensor] = None    # white: float64 [B]
    p: float = 0.0
    sigma0: float = 0.0
    sigma1: float = 0.0

    @staticmethod
    def white(z: torch.Tensor, sigma: torch.Tensor) -> "Noise":
        return Noise(_lib.ASR_NOISE_WHITE, z=z, sigma=sigma)

    @staticmethod
    def mixture(q: torch.Tensor, g: torch.Tensor, p: float, alpha: float) -> "Noise":
        # add_noise: sigma0 = alpha ; sigma1 = 10 * alpha   (VDR/attacks.py:176-178)
        return Noise(_lib.ASR_NOISE_MIXTURE, z=q, z2=g, p=float(p), sigma0=alpha, sigma1=10 * alpha)

    @staticmethod
    def rows_white(z: torch.Tensor, sigma: float) -> "Noise":
        """Feature-domain white noise on an (N, D) matrix: x + sigma*z (add_white_noise_on_dataset, VDR/attacks.py:186-201)."""
        return Noise(_lib.ASR_NOISE_WHITE, z=z, sigma0=float(sigma))

    @staticmethod
    def rows_mixture(q: torch.Tensor, g: torch.Tensor, p: float, alpha: float) -> "Noise":
        """Feature-domain mixture noise (add_noise_mixture_on_dataset, VDR/attacks.py:204-219)."""
        return Noise(_lib.ASR_NOISE_MIXTURE, z=q, z2=g, p=float(p), sigma0=alpha, sigma1=10 * alpha)

    def to_c(self) -> NoiseC:
        for t in (self.z, self.z2, self.sigma):
            if t is not None and (t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous()):
                raise TypeError("noise streams / sigma must be contiguous float64 CUDA tensors")
        return NoiseC(mode=self.mode, reserved=0, z_dev=_ptr(self.z), z2_dev=_ptr(self.z2), sigma_dev=_ptr(self.sigma),
                      p=self.p, sigma0=self.sigma0, sigma1=self.sigma1)


class MfccPlan:
    """Immutable tables for one parameter set (wraps ``asr_plan``)."""


    def __init__(self, params: MfccParams, device: Optional[int] = None, path: str = "auto"):
        """`path`: "auto", "clip" (one CTA per clip), "frames" (block-pipelined n_fft = 512 path), "tiles" (the same with TMA staging) or "tc" (tensor-core path for int16 audio); see asr_path."""
        _require_cuda()
        self.params = params
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            pc = params.to_c()
            check(lib.asr_plan_create(C.byref(pc), C.byref(self._h)), "asr_plan_create")
        check(lib.asr_plan_set_path(self._h, {"auto": 0, "clip": 1, "frames": 2, "tiles": 3, "tc": 4}[path]), "asr_plan_set_path")
        self._ws = {}       # stream -> scratch tensor (the n_fft = 512 path keeps log-mel rows there between its launches)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            lib.asr_plan_destroy(h)
            self._h = None

    @property
    def feature_rows(self) -> int:
        return lib.asr_plan_feature_rows(self._h)

    def set_stage_probe(self, staged: Optional[torch.Tensor]) -> None:
        """Parity probe (asr_plan_set_stage_probe): a float32 CUDA tensor shaped like the batch's audio receives the
        frame samples the TILES path stages (audio + fused noise, before the window); ``None`` switches it off."""
        if staged is not None and (staged.dtype != torch.float32 or not staged.is_cuda or not staged.is_contiguous()):
            raise ValueError("staged must be a contiguous float32 CUDA tensor")
        self._probe = staged
        check(lib.asr_plan_set_stage_probe(self._h, 0 if staged is None else staged.data_ptr()), "asr_plan_set_stage_probe")

    def launches(self, noisy: bool = False) -> int:
        """Kernels one `mfcc` call launches (1, or 3 on the block-pipelined n_fft = 512 path)."""
        return lib.asr_plan_launches(self._h, int(noisy))

    def path_used(self, dtype=np.int16, noisy: bool = False) -> str:
        """Kernel path a call with aligned arrays of `dtype` takes: "clip", "frames" or "tiles" (see asr_path)."""
        code = {np.dtype(np.int16): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}[np.dtype(dtype)]
        return {1: "clip", 2: "frames", 3: "tiles", 4: "tc"}[lib.asr_plan_path_used(self._h, code, 1 if noisy else 0)]

    def _workspace(self, n_clips: int, max_length: int, device) -> Optional[torch.Tensor]:
        need = lib.asr_mfcc_workspace_bytes(self._h, n_clips, max_length)
        if need == 0:
            return None
        key = torch.cuda.current_stream().cuda_stream
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    @property
    def uses_fft(self) -> bool:
        return bool(lib.asr_plan_uses_fft(self._h))

    def num_frames(self, length: int) -> int:
        return lib.asr_plan_num_frames(self._h, int(length))

    def tables(self) -> dict:
        p = self.params
        n_bins = 1 + p.n_fft // 2
        window = np.zeros(p.n_fft, np.float32)
        mel = np.zeros((p.n_mels, n_bins), np.float32)
        dct = np.zeros((p.n_mfcc, p.n_mels), np.float32)
        taps = np.zeros((max(p.delta_orders, 1), p.delta_width), np.float32)
        check(lib.asr_plan_get_tables(self._h, window.ctypes.data, mel.ctypes.data, dct.ctypes.data, taps.ctypes.data),
              "asr_plan_get_tables")
        return {"window": window, "mel": mel, "dct": dct, "delta_taps": taps[:p.delta_orders]}

    def _launch(self, fn, what, batch: ClipBatch, out_frames, noise, out, out_dtype, status, rows):
        B = batch.n_clips
        dev = batch.audio.device
        if out is None:
            out = torch.empty((B, rows, out_frames), dtype=out_dtype, device=dev)
        elif out.shape != (B, rows, out_frames) or not out.is_contiguous() or not out.is_cuda:
            raise ValueError(f"out must be a contiguous CUDA tensor of shape {(B, rows, out_frames)}")
        if status is None:
            status = torch.empty(B, dtype=torch.int32, device=dev)
        nz = noise.to_c() if noise is not None else None
        with torch.cuda.device(dev):
            args = [self._h, batch.audio.data_ptr(), batch.dtype_code, batch.offsets.data_ptr(), batch.lengths.data_ptr(),
                    B, batch.max_length, C.byref(nz) if nz is not None else None, out.data_ptr()]
            if fn is lib.asr_mfcc_batch:
                args.append(_DT[out.dtype])
            ws = self._workspace(B, batch.max_length, dev)
            args += [out_frames, status.data_ptr(), _ptr(ws), 0 if ws is None else ws.numel(), _stream()]
            check(fn(*args), what)
        return out, status

    def mfcc(self, batch: ClipBatch, out_frames: Optional[int] = None, noise: Optional[Noise] = None,
             out: Optional[torch.Tensor] = None, out_dtype=torch.float32, status: Optional[torch.Tensor] = None):
        """Fused MFCC of every clip -> ``(B, rows, out_frames)`` (+ per-clip int32 status), asynchronous."""
        if out_frames is None:
            out_frames = max(1, self.num_frames(batch.max_length))
        if out is not None:
            out_dtype = out.dtype
        return self._launch(lib.asr_mfcc_batch, "asr_mfcc_batch", batch, int(out_frames), noise, out, out_dtype, status,
                            self.feature_rows)

    def logmel(self, batch: ClipBatch, out_frames: Optional[int] = None, noise: Optional[Noise] = None):
        """Stage probe: the clamped log-mel matrix ``(B, n_mels, out_frames)`` in float32."""
        if out_frames is None:
            out_frames = max(1, self.num_frames(batch.max_length))
        return self._launch(lib.asr_logmel_batch, "asr_logmel_batch", batch, int(out_frames), noise, None, torch.float32,
                            None, self.params.n_mels)

    def mfcc_host(self, audio: np.ndarray, offsets: np.ndarray, lengths: np.ndarray, out_frames: int,
                  snr_db: Optional[float] = None, seed: int = 0, out: Optional[np.ndarray] = None,
                  out_dtype=np.float64):
        """``asr_mfcc_batch_host``: host buffers in, reference-layout ``(N, rows*out_frames)`` host matrix out."""
        audio = np.ascontiguousarray(audio)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        n = lengths.shape[0]
        if out is None:
            out = np.empty((n, self.feature_rows * out_frames), dtype=out_dtype)
        status = np.zeros(n, dtype=np.int32)
        code = {np.dtype(np.int16): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}[audio.dtype]
        ocode = {np.dtype(np.float32): 1, np.dtype(np.float64): 2}[out.dtype]
        with torch.cuda.device(self.device):
            check(lib.asr_mfcc_batch_host(self._h, audio.ctypes.data, code, offsets.ctypes.data, lengths.ctypes.data, n,
                                          0 if snr_db is None else 1, 0.0 if snr_db is None else float(snr_db),
                                          int(seed), out.ctypes.data, ocode, int(out_frames), status.ctypes.data),
                  "asr_mfcc_batch_host")
        return out, status


# ---- noise path --------------------------------------------------------------------------------------
def clip_power(batch: ClipBatch, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``np.mean(sample**2)`` per clip in float32, numpy's pairwise order (bit-exact)."""
    if out is None:
        out = torch.empty(batch.n_clips, dtype=torch.float32, device=batch.audio.device)
    elif out.dtype != torch.float32 or out.numel() < batch.n_clips or not out.is_contiguous() or not out.is_cuda:
        raise ValueError("out must be a contiguous float32 CUDA tensor of at least n_clips elements")
    with torch.cuda.device(batch.audio.device):
        check(lib.asr_clip_power(batch.audio.data_ptr(), batch.dtype_code, batch.offsets.data_ptr(),
                                 batch.lengths.data_ptr(), batch.n_clips, out.data_ptr(), _stream()), "asr_clip_power")
    return out


def copy_mapped(dst: torch.Tensor, src: torch.Tensor) -> None:
    """``dst.copy_(src)`` for a SMALL transfer between a pinned host tensor and a CUDA tensor (either direction) as a kernel
    on the current stream instead of a copy-engine operation, so that it does not queue behind a bulk upload / download."""
    if dst.dtype != src.dtype or dst.numel() != src.numel() or not dst.is_contiguous() or not src.is_contiguous():
        raise ValueError("copy_mapped: contiguous tensors of one dtype and size")
    for t in (dst, src):
        if not (t.is_cuda or t.is_pinned()):
            raise ValueError("copy_mapped: tensors must be CUDA or pinned host tensors")
    dev = dst.device if dst.is_cuda else src.device
    with torch.cuda.device(dev):
        check(lib.asr_copy_mapped(src.data_ptr(), dst.data_ptr(), src.numel() * src.element_size(), _stream()), "asr_copy_mapped")


def snr_sigma_host_scalar(power: np.ndarray, target_snr_db) -> np.ndarray:
    """The reference's own scalar lines (VDR/attacks.py:235-241) run clip by clip on ``P`` - the definition of the
    bit-exact sigma (numpy scalar semantics keep every step in float32 for float32 audio).  Slow: a Python loop."""
    out = np.empty(power.shape[0], dtype=np.float64)
    for i in range(power.shape[0]):
        signal_avg_watts = power[i]                       # np.float32 scalar
        signal_avg_db = 10 * np.log10(signal_avg_watts)
        noise_avg_db = signal_avg_db - target_snr_db
        noise_avg_watts = 10 ** (noise_avg_db / 10)
        out[i] = float(np.sqrt(noise_avg_watts))
    return out


def snr_sigma_host(power: np.ndarray, target_snr_db, out: Optional[np.ndarray] = None) -> np.ndarray:
    """Bit-exact sigma for a whole batch in microseconds: the same chain as :func:`snr_sigma_host_scalar`.

    ``np.log10`` is evaluated by numpy itself, vectorised - the ufunc loop a float32 scalar goes through is the same
    SIMD kernel, so the values are those of the scalar chain on THIS host (numpy does not call libm's log10f on
    AVX-512 machines).  The scalar ``10 ** x`` of the reference is libm's powf, which numpy's *vectorised* power is
    not: that step, the float32 multiply / subtract / divide and the square root run in ``asr_snr_sigma_host``
    (C, same libm).  ``tests/test_host_logic.py`` asserts equality with the scalar chain over a million powers."""
    P = np.ascontiguousarray(power, dtype=np.float32)
    if isinstance(target_snr_db, np.floating) and target_snr_db.dtype.itemsize > 4:
        return snr_sigma_host_scalar(P, target_snr_db)     # a float64 numpy scalar promotes the chain: keep the literal text
    n = P.shape[0]
    if out is None:
        out = np.empty(n, dtype=np.float64)
    with np.errstate(divide="ignore"):
        lg = np.log10(P)
    check(lib.asr_snr_sigma_host(P.ctypes.data, lg.ctypes.data, float(np.float32(target_snr_db)), out.ctypes.data, n),
          "asr_snr_sigma_host")
    return out


def snr_sigma_device(power: torch.Tensor, target_snr_db: float) -> torch.Tensor:
    """Same chain evaluated on the device (float64 evaluation rounded to float32 at each step)."""
    out = torch.empty(power.shape[0], dtype=torch.float64, device=power.device)
    with torch.cuda.device(power.device):
        check(lib.asr_snr_sigma(power.data_ptr(), float(target_snr_db), out.data_ptr(), power.shape[0], _stream()),
              "asr_snr_sigma")
    return out


def mix_white(batch: ClipBatch, z: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
    """``float64(x) + sigma[b]*z`` packed like the audio (two roundings, bit-exact with numpy)."""
    out = torch.zeros(batch.audio.shape[0], dtype=torch.float64, device=batch.audio.device)
    with torch.cuda.device(batch.audio.device):
        check(lib.asr_mix_white(batch.audio.data_ptr(), batch.dtype_code, batch.offsets.data_ptr(),
                                batch.lengths.data_ptr(), batch.n_clips, z.data_ptr(), sigma.data_ptr(), out.data_ptr(),
                                _stream()), "asr_mix_white")
    return out


BABBLE_STRIDE, BABBLE_TALKERS = 97, 6      # six other clips of the batch, indices (i + k*97) mod B


_BABBLE_WS: dict = {}


def babble_stream(batch: ClipBatch, stride: int = BABBLE_STRIDE, talkers: int = BABBLE_TALKERS,
                  out: Optional[torch.Tensor] = None, power: Optional[torch.Tensor] = None):
    """Babble stream of a batch (float64, packed like the audio) and its mean power per clip (float64 [B])."""
    dev = batch.audio.device
    if out is None:
        out = torch.zeros(batch.audio.shape[0], dtype=torch.float64, device=dev)
    if power is None:
        power = torch.empty(batch.n_clips, dtype=torch.float64, device=dev)
    need = lib.asr_babble_workspace_bytes(batch.n_clips, batch.max_length)
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    ws = _BABBLE_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, dtype=torch.uint8, device=dev)       # zeroed once; the kernel leaves its counters at zero
        _BABBLE_WS[key] = ws
    with torch.cuda.device(dev):
        check(lib.asr_babble_stream(batch.audio.data_ptr(), batch.dtype_code, batch.offsets.data_ptr(), batch.lengths.data_ptr(),
                                    batch.n_clips, batch.max_length, int(stride), int(talkers), out.data_ptr(), power.data_ptr(),
                                    ws.data_ptr(), ws.numel(), _stream()), "asr_babble_stream")
    return out, power


def babble_gain_host(sigma: np.ndarray, babble_power: np.ndarray) -> np.ndarray:
    """gain[b] = sigma[b] / sqrt(Pb[b]) in float64 (0 where the babble is silent): noise RMS = the white-noise sigma."""
    pb = np.asarray(babble_power, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        g = np.asarray(sigma, dtype=np.float64) / np.sqrt(pb)
    return np.where(pb > 0, g, 0.0)


def mix_mixture(batch: ClipBatch, q: torch.Tensor, g: torch.Tensor, p: float, alpha: float) -> torch.Tensor:
    out = torch.zeros(batch.audio.shape[0], dtype=torch.float64, device=batch.audio.device)
    with torch.cuda.device(batch.audio.device):
        check(lib.asr_mix_mixture(batch.audio.data_ptr(), batch.dtype_code, batch.offsets.data_ptr(),
                                  batch.lengths.data_ptr(), batch.n_clips, q.data_ptr(), g.data_ptr(), float(p),
                                  alpha, 10 * alpha, out.data_ptr(), _stream()), "asr_mix_mixture")
    return out


def mix_rows_white(x: torch.Tensor, z: torch.Tensor, sigma: float) -> torch.Tensor:
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.asr_mix_rows_white(x.data_ptr(), x.numel(), z.data_ptr(), float(sigma), out.data_ptr(), _stream()),
              "asr_mix_rows_white")
    return out


def mix_rows_mixture(x: torch.Tensor, q: torch.Tensor, g: torch.Tensor, p: float, alpha: float) -> torch.Tensor:
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.asr_mix_rows_mixture(x.data_ptr(), x.numel(), q.data_ptr(), g.data_ptr(), float(p), alpha, 10 * alpha,
                                       out.data_ptr(), _stream()), "asr_mix_rows_mixture")
    return out


def randn(seed: int, first_index: int, n: int, device="cuda", out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Seeded float64 standard-normal stream; element i depends only on (seed, first_index + i)."""
    _require_cuda()
    if out is None:
        out = torch.empty(n, dtype=torch.float64, device=device)
    elif out.dtype != torch.float64 or out.numel() != n or not out.is_contiguous() or not out.is_cuda:
        raise ValueError("out must be a contiguous float64 CUDA tensor of n elements")
    with torch.cuda.device(out.device):
        check(lib.asr_randn_f64(int(seed), int(first_index), int(n), out.data_ptr(), _stream()), "asr_randn_f64")
    return out


# ---- audio ingest: resampling on the device ---------------------------------------------------------------
class Resampler:
    """``scipy.signal.resample_poly(x, up, down)`` for a packed batch on the device (the resampling step of
    ``librosa.load``; see ``asr_resample_design`` in include/asr_b200.h for what is restated)."""

    def __init__(self, orig_sr: int, target_sr: int, kaiser_beta: float = 5.0, device="cuda"):
        _require_cuda()
        u, d, n, pre = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        check(lib.asr_resample_design(int(target_sr), int(orig_sr), float(kaiser_beta), None, 0, C.byref(u), C.byref(d),
                                      C.byref(n), C.byref(pre)), "asr_resample_design")
        self.up, self.down, self.n_taps, self.n_pre_remove = u.value, d.value, n.value, pre.value
        self.taps_host = np.zeros(max(self.n_taps, 1), dtype=np.float32)
        if self.n_taps:
            check(lib.asr_resample_design(int(target_sr), int(orig_sr), float(kaiser_beta), self.taps_host.ctypes.data,
                                          self.n_taps, C.byref(u), C.byref(d), C.byref(n), C.byref(pre)), "asr_resample_design")
        self.taps = torch.from_numpy(self.taps_host).to(device)

    def out_length(self, n_in: int) -> int:
        return int(lib.asr_resample_out_len(int(n_in), self.up, self.down))

    def __call__(self, batch: ClipBatch) -> ClipBatch:
        """int16 / float32 clips -> float32 clips at the target rate (a new packed batch, same clip order)."""
        if batch.audio.dtype not in (torch.int16, torch.float32):
            raise TypeError("resampling takes int16 or float32 audio")
        dev = batch.audio.device
        if self.up == 1 and self.down == 1:
            audio = batch.audio if batch.audio.dtype == torch.float32 else batch.audio.to(torch.float32) / 32768.0
            return batch.like(audio)
        out_len = np.array([self.out_length(n) for n in batch.lengths_host], dtype=np.int32)
        offsets, total = ClipBatch.layout(out_len)
        out = torch.zeros(total, dtype=torch.float32, device=dev)
        off_d = torch.from_numpy(offsets).to(dev)
        len_d = torch.from_numpy(out_len).to(dev)
        with torch.cuda.device(dev):
            check(lib.asr_resample_batch(batch.audio.data_ptr(), batch.dtype_code, batch.offsets.data_ptr(),
                                         batch.lengths.data_ptr(), batch.n_clips, batch.max_length, self.up, self.down,
                                         self.taps.data_ptr(), self.n_taps, self.n_pre_remove, out.data_ptr(),
                                         off_d.data_ptr(), _stream()), "asr_resample_batch")
        return ClipBatch(out, off_d, len_d, int(out_len.max()) if len(out_len) else 0, offsets, out_len)


# ---- standardisation ---------------------------------------------------------------------------------
class Standardizer:
    """``StandardScaler().fit_transform`` over row blocks that may live on several GPUs
    (``standardize_dataset``, VDR/attacks.py:48-69).

    sklearn's two passes run on this rank's rows (pass 1: column sums; pass 2: sums of ``x - m`` and ``(x - m)^2`` about
    the LOCAL mean), the rank's message ``[n, S, C, Q]`` (3 D + 1 float64) is exchanged ONCE (one all-gather over
    NCCL / NVLink, the only inter-GPU exchange of the path) and every rank merges the messages in rank order, moving the
    centred sums to the global mean (Chan's update) - with one rank the merge is the identity.  On a single GPU the
    whole ``fit`` + ``transform`` of one matrix is three launches (`fit_transform`): the last reduction of each pass
    happens inside the next launch.

    ``noise=`` (``Noise.rows_white`` / ``Noise.rows_mixture``) makes every pass read ``x + noise`` instead of ``x``: the
    MFCC-domain attacks of VDR/attacks.py:186-219 followed by ``standardize_dataset`` (:433-491) without writing the
    noisy matrix.
    """

    def __init__(self, n_cols: int, device="cuda", group=None, distributed: bool = False):
        self.n_cols = int(n_cols)
        self.device = torch.device(device)
        self.group = group
        self.distributed = distributed
        D = self.n_cols
        # persistent float64 vectors (stable addresses: the fit can be captured in CUDA graphs)
        self.msg = torch.zeros(3 * D + 1, dtype=torch.float64, device=self.device)     # [n, S (D), C (D), Q (D)] of this rank
        self.msgs = self.msg.view(1, -1)                                               # all ranks' messages, rank order
        self.mean = torch.zeros(D, dtype=torch.float64, device=self.device)
        self.var = torch.zeros(D, dtype=torch.float64, device=self.device)
        self.scale = torch.ones(D, dtype=torch.float64, device=self.device)
        self.n_dev = torch.zeros(1, dtype=torch.float64, device=self.device)           # rows over all ranks (device copy)
        self._n_total: Optional[int] = None
        self._ws: Optional[torch.Tensor] = None
        self._slabs = (0, 0)
        self._n_local = 0
        # transport of the one exchange: "p2p" (one-shot all-gather over peer memory, asr_cmvn_exchange_p2p: no NCCL call,
        # capturable in a CUDA graph) when the ranks' symmetric regions can be set up, else "nccl" (torch.distributed all-gather)
        self.exchange_transport = "none"
        self._p2p = None
        if distributed:
            self.exchange_transport = "nccl"
            if os.environ.get("ASR_B200_P2P_EXCHANGE", "1") != "0":
                self._setup_p2p()

    @property
    def n_total(self) -> int:
        if self._n_total is None:
            self._n_total = int(round(float(self.n_dev.item())))
        return self._n_total

    def _world(self) -> int:
        if not self.distributed:
            return 1
        import torch.distributed as dist
        return dist.get_world_size(self.group)

    def _workspace(self) -> torch.Tensor:
        if self._ws is None:
            self._ws = torch.empty(lib.asr_cmvn_workspace_bytes(self.n_cols), dtype=torch.uint8, device=self.device)
        return self._ws

    @staticmethod
    def _mat(x: torch.Tensor):
        if x.dim() != 2 or x.stride(1) != 1 or x.dtype not in (torch.float32, torch.float64):
            raise ValueError("row blocks must be 2-D float32/float64 with unit column stride")
        return x.data_ptr(), _DT[x.dtype], x.shape[0], x.shape[1], x.stride(0)

    @staticmethod
    def _nz(noise):
        return C.byref(noise.to_c()) if noise is not None else None

    # ---- the launch groups of a fit (the sharded step is: local_stats, local_message | exchange | merge) ----
    def local_stats(self, blocks: Sequence[torch.Tensor], noises: Optional[Sequence] = None) -> None:
        """Both passes over this rank's row blocks: slab partials in the workspace (2 launches per block)."""
        D = self.n_cols
        ws = self._workspace()
        noises = list(noises) if noises is not None else [None] * len(blocks)
        rows = sum(int(x.shape[0]) for x in blocks)
        with torch.cuda.device(self.device):
            counts = [0, 0]
            for pas in (1, 2):
                for x, nz in zip(blocks, noises):
                    p, dt, r, c, ld = self._mat(x)
                    if c != D:
                        raise ValueError(f"row block has {c} columns, expected {D}")
                    if r == 0:
                        continue
                    keep = nz.to_c() if nz is not None else None
                    n = lib.asr_cmvn_partial_sums(p, dt, r, c, ld, C.byref(keep) if keep is not None else None, pas, counts[0],
                                                  max(rows, 1), counts[pas - 1], ws.data_ptr(), ws.numel(), _stream())
                    if n < 0:
                        check(n, "asr_cmvn_partial_sums")
                    counts[pas - 1] += n
        self._slabs = (counts[0], counts[1])
        self._n_local = rows
        self._n_total = None

    def local_message(self) -> None:
        ws = self._workspace()
        with torch.cuda.device(self.device):
            check(lib.asr_cmvn_local_message(ws.data_ptr(), ws.numel(), self._slabs[0], self._slabs[1], self._n_local, self.n_cols,
                                             self.msg.data_ptr(), _stream()), "asr_cmvn_local_message")

    def _setup_p2p(self) -> None:
        """Symmetric regions of all ranks mapped into this process (torch's symmetric memory is the plumbing: allocation,
        handle exchange, peer mapping); any failure leaves the NCCL transport in place."""
        import torch.distributed as dist
        if dist.get_backend(self.group) != "nccl" or self.device.type != "cuda":
            return
        try:
            import torch.distributed._symmetric_memory as symm
            grp = self.group if self.group is not None else dist.group.WORLD
            w, r = dist.get_world_size(grp), dist.get_rank(grp)
            nbytes = int(lib.asr_cmvn_p2p_region_bytes(w, self.n_cols))
            region = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
            region.zero_()
            torch.cuda.synchronize(self.device)
            hdl = symm.rendezvous(region, grp)
            state = torch.zeros(w, dtype=torch.int32, device=self.device)
            msgs = torch.zeros((w, self.msg.numel()), dtype=torch.float64, device=self.device)
            ptrs = torch.tensor([int(a) for a in hdl.buffer_ptrs], dtype=torch.int64, device=self.device)
            torch.cuda.synchronize(self.device)
            dist.barrier(grp)                                   # every region is zeroed and mapped before anybody writes
            self._p2p = {"region": region, "hdl": hdl, "state": state, "ptrs": ptrs, "rank": r, "world": w}
            self.msgs = msgs
            self.exchange_transport = "p2p"
        except Exception as e:                                  # noqa: BLE001
            self._p2p = None
            self._p2p_error = str(e).splitlines()[0][:200] if str(e) else type(e).__name__
        # every rank must use the same transport: agree (a rank whose set-up failed takes everybody to NCCL)
        try:
            ok = torch.tensor([1 if self._p2p is not None else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                self._p2p = None
        except Exception:                                       # noqa: BLE001
            self._p2p = None
        if self._p2p is None:
            self.exchange_transport = "nccl"
            self.msgs = self.msg.view(1, -1)

    def exchange(self) -> None:
        """The one collective of the path: every rank's message to every rank."""
        if not self.distributed:
            self.msgs = self.msg.view(1, -1)
            return
        if self._p2p is not None:
            q = self._p2p
            with torch.cuda.device(self.device):
                check(lib.asr_cmvn_exchange_p2p(self.msg.data_ptr(), self.n_cols, q["rank"], q["world"], q["ptrs"].data_ptr(),
                                                q["state"].data_ptr(), self.msgs.data_ptr(), _stream()), "asr_cmvn_exchange_p2p")
            return
        import torch.distributed as dist
        w = self._world()
        if self.msgs.shape[0] != w or self.msgs.data_ptr() == self.msg.data_ptr():
            self.msgs = torch.zeros((w, self.msg.numel()), dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(self.msgs.view(-1), self.msg, group=self.group)

    def merge(self) -> None:
        with torch.cuda.device(self.device):
            check(lib.asr_cmvn_merge(self.msgs.data_ptr(), self.msgs.shape[0], self.n_cols, self.mean.data_ptr(), self.var.data_ptr(),
                                     self.scale.data_ptr(), self.n_dev.data_ptr(), _stream()), "asr_cmvn_merge")
        self._n_total = None

    def fit(self, blocks: Sequence[torch.Tensor], n_total: Optional[int] = None, noises: Optional[Sequence] = None) -> "Standardizer":
        """Statistics of the rows of all ranks.  (`n_total` is accepted for compatibility: the merge derives it.)"""
        self.local_stats(blocks, noises)
        self.local_message()
        self.exchange()
        self.merge()
        if n_total is not None:
            self._n_total = int(n_total)
        return self

    def transform(self, x: torch.Tensor, out_dtype=torch.float64, out: Optional[torch.Tensor] = None, noise=None) -> torch.Tensor:
        return self._apply(x, out_dtype, out, noise, fused=False)

    def fit_transform(self, x: torch.Tensor, out_dtype=torch.float64, out: Optional[torch.Tensor] = None, noise=None) -> torch.Tensor:
        """Single rank, one matrix: statistics and standardised rows in three launches."""
        if self.distributed:
            self.fit([x], noises=[noise])
            return self.transform(x, out_dtype, out, noise)
        self.local_stats([x], [noise])
        res = self._apply(x, out_dtype, out, noise, fused=True)
        self._n_total = int(x.shape[0])
        return res

    def _apply(self, x, out_dtype, out, noise, fused):
        p, dt, r, c, ld = self._mat(x)
        if out is None:
            out = torch.empty((r, c), dtype=out_dtype, device=x.device)
        elif out.shape != (r, c) or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous {(r, c)} tensor")
        out_dtype = out.dtype
        keep = noise.to_c() if noise is not None else None
        ws = self._workspace() if fused else None
        with torch.cuda.device(x.device):
            check(lib.asr_cmvn_apply2(p, dt, r, c, ld, C.byref(keep) if keep is not None else None,
                                      ws.data_ptr() if fused else None, ws.numel() if fused else 0, self._slabs[0], self._slabs[1],
                                      max(self._n_local, 1), self.mean.data_ptr(), self.var.data_ptr(), self.scale.data_ptr(),
                                      out.data_ptr(), _DT[out_dtype], _stream()), "asr_cmvn_apply2")
        return out
