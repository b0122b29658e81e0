"""Per-phase cycle breadcrumbs of the tile kernel (CTA 0: main warp 0, descriptor warp, helper warp 1).
Run with ASR_B200_DBG_SKIP=128 (plus skip bits / ASR_B200_TILE_VAR for experiments)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200 as A
from asr_b200._lib import lib
from synth import synth_clips

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
noisy = len(sys.argv) > 2 and sys.argv[2] == "noisy"
base = np.stack(synth_clips(256, 16000, 16000, 3))
host = np.concatenate([np.roll(base, 37 * r, axis=1) for r in range((B + 255) // 256)], axis=0)[:B]
batch = A.ClipBatch.from_matrix(torch.from_numpy(np.ascontiguousarray(host)).cuda())
noise = None
if noisy:
    z = A.randn(1, 0, B * 16000)
    noise = A.Noise.white(z, torch.from_numpy(A.snr_sigma_host(A.clip_power(batch).cpu().numpy(), 10)).cuda())
plan = A.MfccPlan(A.C1, path="tiles")
for _ in range(3):
    out, st = plan.mfcc(batch, noise=noise)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
out, st = plan.mfcc(batch, noise=noise)
b.record()
torch.cuda.synchronize()
w = [lib.asr_plan_debug_word(plan._h, i) * 16 for i in range(16)]
print(f"B={B} noisy={noisy} call {a.elapsed_time(b):.3f} ms  (cycles of CTA 0, whole launch)")
print("  main warp 0 :", dict(zip(["combine", "fft", "bar_main", "mel", "end_barrier"], w[0:5])), "sum", sum(w[0:5]))
print("  descr warp  :", dict(zip(["assemble", "tma_wait", "convert", "bar_help", "issue", "end_barrier"], w[5:11])), "sum", sum(w[5:11]))
print("  helper warp1:", dict(zip(["tma_wait", "convert", "bar_help", "issue", "end_barrier"], w[11:16])), "sum", sum(w[11:16]))
