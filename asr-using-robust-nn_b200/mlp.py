"""the classifier's forward pass on the GPU, so that the accuracy-vs-SNR sweep of
``attacks.py`` (VDR/attacks.py:401-422) closes on the device: noisy MFCC (fused launch) -> standardise with the
training statistics -> ``model.predict`` -> accuracy.

The network is the reference's ``get_model`` (VDR/train_constraints.py:63-88): 880 -> 1024 -> 512 -> 256 -> 128 -> 64
(ReLU, BatchNormalization after each) -> 10 (softmax).  At inference every BatchNormalization is an affine map; it is
folded into the FOLLOWING Dense layer once, at construction (in float64, rounded to float32), so a forward pass is six
float32 affine maps (+ ReLU) and a softmax: ONE launch of this repository's ``mlp_forward_kernel`` through the C-ABI entry
``asr_mlp_forward`` (``csrc/mlp_kernel.cu``: 16 rows per CTA through all layers, activations in shared memory, weights
streamed from L2, float32 FMA accumulation like ``model.predict``); logits, probabilities and decisions come out of that
launch.  ``logits_library`` keeps the cuBLAS form (``torch.addmm``, TF32 off) as an independent check for the tests.  The
reference's trained ``.h5`` files are not in its tree, so weights come from the caller (``from_keras_weights``) or are
synthetic.
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence

import ctypes as C

import numpy as np
import torch

from ._lib import lib, check

BN_EPS = 1e-3      # keras.layers.BatchNormalization default epsilon


class DenseStack:
    """Dense(relu) -> BatchNorm ... -> Dense(softmax), BatchNorm folded forward."""

    def __init__(self, layers: Sequence[dict], device="cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the asr_b200 path has no CPU fallback")
        self.device = torch.device(device)
        ws, bs = [], []
        scale = shift = None                               # affine map pending from the previous BatchNormalization
        for i, ly in enumerate(layers):
            k = np.asarray(ly["kernel"], dtype=np.float64)
            b = np.asarray(ly["bias"], dtype=np.float64)
            if scale is not None:                          # (s*h + t) @ K + b = h @ (s[:,None]*K) + (t @ K + b)
                b = b + shift @ k
                k = scale[:, None] * k
            ws.append(torch.from_numpy(k.astype(np.float32)).to(self.device))
            bs.append(torch.from_numpy(b.astype(np.float32)).to(self.device))
            if "gamma" in ly:
                inv = np.asarray(ly["gamma"], np.float64) / np.sqrt(np.asarray(ly["moving_var"], np.float64) + BN_EPS)
                scale, shift = inv, np.asarray(ly["beta"], np.float64) - np.asarray(ly["moving_mean"], np.float64) * inv
            else:
                scale = shift = None
        self.weights, self.biases = [w.contiguous() for w in ws], [b.contiguous() for b in bs]
        self.n_in, self.n_out = ws[0].shape[0], ws[-1].shape[1]
        n = len(ws)
        if n > 8 or max(w.shape[1] for w in ws) > 1024:
            raise ValueError("asr_mlp_forward takes 1..8 layers with outputs up to 1024 wide")
        self._dims = (C.c_int32 * (n + 1))(*([ws[0].shape[0]] + [w.shape[1] for w in ws]))
        self._wptr = (C.c_void_p * n)(*[w.data_ptr() for w in self.weights])
        self._bptr = (C.c_void_p * n)(*[b.data_ptr() for b in self.biases])

    @staticmethod
    def from_keras_weights(arrays: Iterable[np.ndarray], device="cuda") -> "DenseStack":
        """``model.get_weights()`` order of the reference model: per hidden layer kernel, bias, gamma, beta,
        moving_mean, moving_variance; then the output layer's kernel, bias."""
        a = [np.asarray(x) for x in arrays]
        layers, i = [], 0
        while i < len(a):
            ly = {"kernel": a[i], "bias": a[i + 1]}
            i += 2
            if i + 3 < len(a) and a[i].ndim == 1 and a[i].shape[0] == ly["kernel"].shape[1] and a[i + 3].ndim == 1 \
                    and (i + 4 >= len(a) or a[i + 4].ndim == 2):
                ly.update(gamma=a[i], beta=a[i + 1], moving_mean=a[i + 2], moving_var=a[i + 3])
                i += 4
            layers.append(ly)
        return DenseStack(layers, device=device)

    def _forward(self, x: torch.Tensor, want_logits: bool, want_probs: bool, want_argmax: bool):
        """One ``asr_mlp_forward`` launch on the current stream; returns (logits, probs, argmax), None where not asked."""
        if x.dim() != 2 or x.shape[1] != self.n_in:
            raise ValueError(f"expected (N, {self.n_in}) features")
        h = x.to(self.device, torch.float32)
        if h.stride(1) != 1:
            h = h.contiguous()
        n = h.shape[0]
        lg = torch.empty((n, self.n_out), dtype=torch.float32, device=self.device) if want_logits else None
        pr = torch.empty((n, self.n_out), dtype=torch.float32, device=self.device) if want_probs else None
        am = torch.empty(n, dtype=torch.int32, device=self.device) if want_argmax else None
        with torch.cuda.device(self.device):
            check(lib.asr_mlp_forward(h.data_ptr(), n, h.stride(0) if n > 0 else self.n_in, len(self.weights), self._dims, self._wptr,
                                      self._bptr, lg.data_ptr() if want_logits else None, pr.data_ptr() if want_probs else None,
                                      am.data_ptr() if want_argmax else None, torch.cuda.current_stream(self.device).cuda_stream),
                  "asr_mlp_forward")
        return lg, pr, am

    def logits(self, x: torch.Tensor) -> torch.Tensor:
        return self._forward(x, True, False, False)[0]

    def logits_library(self, x: torch.Tensor) -> torch.Tensor:
        """The same forward pass through cuBLAS (``torch.addmm``, TF32 off): an independent implementation for the tests."""
        if x.dim() != 2 or x.shape[1] != self.n_in:
            raise ValueError(f"expected (N, {self.n_in}) features")
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False      # model.predict is float32
        try:
            h = x.to(self.device, torch.float32)
            for i, (w, b) in enumerate(zip(self.weights, self.biases)):
                h = torch.addmm(b, h, w)
                if i < len(self.weights) - 1:
                    h = torch.relu_(h)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        return h

    def predict(self, x: torch.Tensor) -> torch.Tensor:
        """``model.predict``: float32 softmax probabilities (N, n_classes), on the device."""
        return self._forward(x, False, True, False)[1]

    def decide(self, x: torch.Tensor) -> torch.Tensor:
        """``np.argmax(model.predict(x), axis=1)`` (VDR/attacks.py:412), int32 on the device."""
        return self._forward(x, False, False, True)[2]

    def accuracy(self, x: torch.Tensor, labels: torch.Tensor) -> float:
        """VDR/attacks.py:412-414; `labels` are class indices or one-hot rows."""
        lab = labels.to(self.device)
        if lab.dim() == 2:
            lab = lab.argmax(dim=1)
        return float((self.decide(x).long() == lab.long()).float().mean().item())


def accuracy_vs_snr(models: Sequence[DenseStack], batch, labels: torch.Tensor, snrs: Sequence[float], plan,
                    standardizer, seed: int = 0, out_frames: Optional[int] = None):
    """The SNR branch of the black-box sweep (VDR/attacks.py:401-422) on the device: for every SNR the test audio is
    mixed with seeded white noise at that SNR inside the MFCC launch, the (N, n_mfcc*T) rows are standardised with the
    statistics `standardizer` was fitted with (train + dev + test rows, VDR/attacks.py:48-69), every model predicts.
    Returns ``{snr: [accuracy per model]}``."""
    from .frontend import Noise, clip_power, snr_sigma_host, randn
    out = {}
    power = clip_power(batch).cpu().numpy()               # bit-exact P; the sigma chain is the reference's own (host)
    n = batch.audio.shape[0]
    for i, snr in enumerate(snrs):
        z = randn(seed + i, 0, n, device=batch.audio.device)
        noise = Noise.white(z, torch.from_numpy(snr_sigma_host(power, snr)).to(z.device))
        feats, _ = plan.mfcc(batch, out_frames=out_frames, noise=noise)
        rows = standardizer.transform(feats.reshape(feats.shape[0], -1), out_dtype=torch.float32)
        out[snr] = [m.accuracy(rows, labels) for m in models]
    return out
