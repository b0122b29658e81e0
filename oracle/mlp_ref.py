"""CPU restatement of the reference's classifier forward pass (TEST INFRASTRUCTURE - only tests/ may import it).

``get_model`` (VDR/train_constraints.py:63-88, SR twin train_constraints.py): Input(880) -> [Dense(relu) ->
BatchNormalization -> Dropout] x 3 -> [Dense(relu) -> BatchNormalization] x 2 -> Dense(10, softmax).  At inference
(``model.predict``, VDR/attacks.py:409-410) Dropout is the identity and Keras BatchNormalization is
``gamma * (x - moving_mean) / sqrt(moving_var + epsilon) + beta`` with epsilon = 1e-3, everything in float32.
The trained ``.h5`` weights are not in the reference tree (LFS blobs), so parity is on synthetic weights only.
"""
from __future__ import annotations

import numpy as np

LAYER_SIZES = (880, 1024, 512, 256, 128, 64, 10)     # VDR/train_constraints.py:66-85
BN_EPS = 1e-3                                          # keras.layers.BatchNormalization default


def random_weights(seed: int = 0, sizes=LAYER_SIZES):
    """Synthetic weights of the reference architecture: non-negative kernels (``kernel_constraint=NonNeg()``),
    BatchNormalization statistics after every hidden layer."""
    rng = np.random.default_rng(seed)
    layers = []
    for i in range(len(sizes) - 1):
        fan_in, fan_out = sizes[i], sizes[i + 1]
        layer = {
            "kernel": np.abs(rng.standard_normal((fan_in, fan_out))).astype(np.float32) * np.float32(1.0 / fan_in),
            "bias": (0.1 * rng.standard_normal(fan_out)).astype(np.float32),
        }
        if i < len(sizes) - 2:
            layer.update(gamma=rng.uniform(0.5, 1.5, fan_out).astype(np.float32),
                         beta=(0.1 * rng.standard_normal(fan_out)).astype(np.float32),
                         moving_mean=(0.2 * rng.standard_normal(fan_out)).astype(np.float32),
                         moving_var=rng.uniform(0.5, 2.0, fan_out).astype(np.float32))
        layers.append(layer)
    return layers


def predict(x: np.ndarray, layers) -> np.ndarray:
    """``model.predict(x)``: float32 softmax probabilities (N, n_classes)."""
    h = np.asarray(x, dtype=np.float32)
    for i, ly in enumerate(layers):
        h = h @ ly["kernel"] + ly["bias"]
        if i < len(layers) - 1:
            h = np.maximum(h, np.float32(0))
            h = ly["gamma"] * (h - ly["moving_mean"]) / np.sqrt(ly["moving_var"] + np.float32(BN_EPS)) + ly["beta"]
    h = h - h.max(axis=1, keepdims=True)
    e = np.exp(h)
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


def accuracy(pred: np.ndarray, labels_onehot: np.ndarray) -> float:
    """VDR/attacks.py:412-414: ``np.sum(argmax(pred) == argmax(labels)) / len(labels)``."""
    return float(np.sum(np.argmax(pred, axis=1) == np.argmax(labels_onehot, axis=1)) / len(labels_onehot))
