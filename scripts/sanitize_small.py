"""Tiny end-to-end run of every kernel for compute-sanitizer (memcheck): ragged clips, both MFCC paths, noise, power,
standardisation, resampler.  Sizes are small on purpose."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200 as A  # noqa: E402
from synth import synth_clips  # noqa: E402

lengths = [16000, 300, 5000, 257, 16000, 100, 9999, 16000, 480, 7000]
clips = synth_clips(len(lengths), 0, 16000, 1, lengths=lengths)
batch = A.ClipBatch.from_arrays(clips)
z = A.randn(3, 0, batch.audio.shape[0])
sig = A.snr_sigma_device(A.clip_power(batch), 10.0)
for preset, kw in (("c1", {}), ("c3", {}), ("ref_sr", {}), ("c5", {})):
    for path in ("clip", "frames"):
        plan = A.MfccPlan(A.PRESETS[preset], path=path)
        for noise in (None, A.Noise.white(z, sig), A.Noise.mixture(z, z, 0.01, 0.004)):
            out, st = plan.mfcc(batch, noise=noise)
            lm, _ = plan.logmel(batch)
        torch.cuda.synchronize()
        print(preset, path, "ok", float(out.abs().max()))
f32 = A.ClipBatch.from_arrays([c.astype(np.float32) / 32768 for c in clips])
A.MfccPlan(A.C1, path="frames").mfcc(f32, noise=A.Noise.white(z, sig))
mixed = A.mix_white(batch, z, sig)
rs = A.Resampler(16000, 22050)(batch)
feats = torch.randn(50, 131, device="cuda")
st = A.Standardizer(131).fit([feats])
st.transform(feats)
torch.cuda.synchronize()
print("all ok")
