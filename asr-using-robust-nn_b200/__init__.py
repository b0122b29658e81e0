"""B200-native (sm_100a) MFCC front end, additive-noise mixing and dataset standardisation.

Drop-in for the data-parallel hot path of fmazilu/ASR-using-robust-NN:

* ``asr_b200.voice_digit.extract_features_construct_dataset`` / ``.attacks`` mirror
  ``Voice digit recogniton/extract_features_construct_dataset.py`` / ``attacks.py``;
* ``asr_b200.speaker.*`` mirror ``Speaker recognition/*``;
* ``asr_b200.frontend`` is the batched array-in API they delegate to;
* all arithmetic of the hot path runs in ``libasr_b200.so`` (hand-written CUDA, C-ABI in ``include/asr_b200.h``);
* ``asr_b200.mlp`` is the classifier's forward pass for the accuracy-vs-SNR sweep (BatchNorm
  folded, one fused launch of ``asr_mlp_forward``).
"""
from ._lib import AsrError, LIB_PATH  # noqa: F401
from .params import MfccParams, REF_VDR, REF_SR, C1, C3, C5, PRESETS  # noqa: F401
from .frontend import (ClipBatch, Noise, MfccPlan, Standardizer, Resampler, clip_power, snr_sigma_host, snr_sigma_host_scalar,  # noqa: F401
                       snr_sigma_device, mix_white, mix_mixture, babble_stream, babble_gain_host, mix_rows_white, mix_rows_mixture, randn)

from .mlp import DenseStack, accuracy_vs_snr  # noqa: F401

__version__ = "0.1.0"
