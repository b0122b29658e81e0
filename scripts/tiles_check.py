"""Quick GPU comparison of the tiles path against the per-clip kernel (development aid)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200 as A  # noqa: E402
from synth import synth_clips  # noqa: E402


def run(preset, clips, path, noise=None):
    plan = A.MfccPlan(A.PRESETS[preset], path=path)
    batch = A.ClipBatch.from_arrays(clips)
    nz = None
    if noise:
        z = A.randn(7, 0, batch.audio.shape[0])
        sig = A.snr_sigma_device(A.clip_power(batch), 10.0)
        nz = A.Noise.white(z, sig)
    out, st = plan.mfcc(batch, noise=nz)
    torch.cuda.synchronize()
    return out.cpu().numpy(), st.cpu().numpy()


cases = [
    ("c1 4 clips", "c1", synth_clips(4, 16000, 16000, 1)),
    ("c1 300 clips", "c1", synth_clips(300, 16000, 16000, 2)),
    ("c1 ragged", "c1", synth_clips(10, 0, 16000, 3, lengths=[16000, 300, 5000, 257, 16000, 100, 9999, 16000, 480, 7000])),
    ("c1 many tiny", "c1", synth_clips(200, 0, 16000, 4, lengths=[300 + 37 * (i % 11) for i in range(200)])),
    ("c3 ragged", "c3", synth_clips(40, 0, 16000, 5, lengths=[16000 + 997 * i for i in range(40)])),
]
for name, preset, clips in cases:
    f32 = [c.astype(np.float32) / np.float32(32768.0) for c in clips]
    variants = [("int16", clips, False), ("int16", clips, True), ("float32", f32, False), ("float32", f32, True),
                ("float64", [c.astype(np.float64) for c in f32], False)]
    for dname, cl, noise in variants:
        t0 = time.time()
        ref, st0 = run(preset, cl, "clip", noise)
        got, st1 = run(preset, cl, "tiles", noise)
        err = float(np.abs(got - ref).max())
        print(f"{name:14s} noise={int(noise)} dtype={dname:8s} max|tiles-clip|={err:.3e} status_eq={bool((st0 == st1).all())} "
              f"ref_absmax={float(np.abs(ref).max()):.1f} {time.time() - t0:.1f}s", flush=True)
print("done")
