"""the classifier's forward pass on the GPU, so that the accuracy-vs-SNR sweep of
``attacks.py`` (VDR/attacks.py:401-422) closes on the device: noisy MFCC (fused launch) -> standardise with the
training statistics -> ``model.predict`` -> accuracy.

The network is the reference's ``get_model`` (VDR/train_constraints.py:63-88): 880 -> 1024 -> 512 -> 256 -> 128 -> 64
(ReLU, BatchNormalization after each) -> 10 (softmax).  At inference every BatchNormalization is an affine map; it is
folded into the FOLLOWING Dense layer once, at construction (in float64, rounded to float32), so a forward pass is six
plain float32 GEMMs with bias (+ ReLU) - library GEMMs (cuBLAS through ``torch.addmm``, TF32 off), which is what the
build rules prescribe for plain GEMMs; the softmax / argmax / accuracy stay on the device.  The reference's trained
``.h5`` files are not in its tree, so weights come from the caller (``from_keras_weights``) or are synthetic.
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence

import numpy as np
import torch

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
        self.weights, self.biases = ws, bs
        self.n_in, self.n_out = ws[0].shape[0], ws[-1].shape[1]

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

    def logits(self, x: torch.Tensor) -> torch.Tensor:
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
        return torch.softmax(self.logits(x), dim=1)

    def accuracy(self, x: torch.Tensor, labels: torch.Tensor) -> float:
        """VDR/attacks.py:412-414; `labels` are class indices or one-hot rows."""
        lab = labels.to(self.device)
        if lab.dim() == 2:
            lab = lab.argmax(dim=1)
        return float((self.logits(x).argmax(dim=1) == lab).float().mean().item())


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
