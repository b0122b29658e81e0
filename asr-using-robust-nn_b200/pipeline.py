"""The ``test_dataset_to_add_noise`` path as one device pipeline (what ``bench.py`` times).

Per batch of clips (this rank's shard): bit-exact clip power -> SNR sigma -> white-noise mix fused
into the MFCC launch -> dataset standardisation of the feature rows (two-pass column statistics,
all-reduced over NCCL when sharded) -> standardised ``(B, n_mfcc*T)`` rows.  Mirrors the loop body
of ``Voice digit recogniton/attacks.py:402-407`` (``black_box_attack_on_audio_dataset_snr`` then
``standardize_dataset``).
"""
from __future__ import annotations

from typing import Optional

import torch

from .frontend import ClipBatch, MfccPlan, Noise, Standardizer, clip_power, snr_sigma_device, randn
from .params import MfccParams

# kernels of libasr_b200 launched by one `run_device` step (power, sigma, mfcc, 2x colsum(partial+final),
# mean, finalize, apply); the e2e step adds the randn launch
LAUNCHES_PER_STEP = 10
LAUNCHES_PER_STEP_CLEAN = 8


class NoisyFeaturePipeline:
    def __init__(self, params: MfccParams, out_frames: int, device=None, distributed: bool = False, group=None,
                 world_size: int = 1):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.plan = MfccPlan(params, self.device.index)
        self.out_frames = int(out_frames)
        self.rows = self.plan.feature_rows
        self.D = self.rows * self.out_frames
        self.world_size = world_size
        self.std = Standardizer(self.D, device=self.device, group=group, distributed=distributed)
        self._feats = None
        self.ev_mfcc = None          # optional (start, end) CUDA events around the MFCC launch

    def _feat_buffer(self, B: int) -> torch.Tensor:
        if self._feats is None or self._feats.shape[0] != B:
            self._feats = torch.empty((B, self.rows, self.out_frames), dtype=torch.float32, device=self.device)
        return self._feats

    def run_device(self, batch: ClipBatch, z: Optional[torch.Tensor], snr_db: Optional[float],
                   standardize: bool = True, out_dtype=torch.float32) -> torch.Tensor:
        """Inputs resident in HBM; everything is asynchronous on the current stream."""
        noise = None
        if snr_db is not None:
            sigma = snr_sigma_device(clip_power(batch), snr_db)
            noise = Noise.white(z, sigma)
        feats = self._feat_buffer(batch.n_clips)
        if self.ev_mfcc is not None:
            self.ev_mfcc[0].record()
        self.plan.mfcc(batch, out_frames=self.out_frames, noise=noise, out=feats)
        if self.ev_mfcc is not None:
            self.ev_mfcc[1].record()
        flat = feats.view(batch.n_clips, self.D)
        if not standardize:
            return flat
        self.std.fit([flat], n_total=batch.n_clips * self.world_size)
        return self.std.transform(flat, out_dtype=out_dtype)

    def run_host(self, audio_host: torch.Tensor, snr_db: Optional[float], seed: int, out_host: torch.Tensor,
                 first_index: int = 0) -> torch.Tensor:
        """End to end from PINNED host memory: (B, L) int16/float32 host tensor in, standardised
        float32 (B, D) rows written to the pinned `out_host`.  The noise stream is generated on the
        device from `seed` (element index = first_index + position in this shard)."""
        dev_audio = audio_host.to(self.device, non_blocking=True)
        batch = ClipBatch.from_matrix(dev_audio)
        z = randn(seed, first_index, dev_audio.numel(), device=self.device) if snr_db is not None else None
        out = self.run_device(batch, z, snr_db)
        out_host.copy_(out, non_blocking=True)
        return out_host
