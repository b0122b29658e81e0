"""Bisect helper for the tensor-core path: equal-length batches of growing size, clean and noisy, TC against TILES."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200 as A
from synth import synth_clips

base = np.stack(synth_clips(256, 16000, 16000, 3))
sizes = [int(a) for a in sys.argv[1:]] or [64, 512, 2048, 8192]
for B in sizes:
    host = np.concatenate([np.roll(base, 37 * r, axis=1) for r in range((B + 255) // 256)], axis=0)[:B]
    audio = torch.from_numpy(np.ascontiguousarray(host)).cuda()
    batch = A.ClipBatch.from_matrix(audio)
    for noisy in (False, True):
        noise = None
        if noisy:
            z = A.randn(1, 0, B * 16000)
            noise = A.Noise.white(z, torch.from_numpy(A.snr_sigma_host(A.clip_power(batch).cpu().numpy(), 10)).cuda())
        ptc, ptl = A.MfccPlan(A.C1, path="tc"), A.MfccPlan(A.C1, path="tiles")
        ref, _ = ptl.mfcc(batch, noise=noise)
        torch.cuda.synchronize()
        reps = int(os.environ.get("TC_REPS", "3"))
        t0 = time.perf_counter()
        try:
            for rep in range(reps):                 # back to back, no synchronisation in between
                out, st = ptc.mfcc(batch, noise=noise)
            torch.cuda.synchronize()
        except Exception as e:                                   # noqa: BLE001
            from asr_b200._lib import lib
            print("FAILED after", time.perf_counter() - t0, "s at rep", rep, ":", str(e).splitlines()[0][:80],
                  "breadcrumb", [lib.asr_plan_debug_word(ptc._h, i) for i in range(4)], flush=True)
            raise SystemExit(1)
        dt = (time.perf_counter() - t0) / reps
        if int(os.environ.get("ASR_B200_DBG_SKIP", "0")) & 128:
            from asr_b200._lib import lib
            names = ["wait_hl", "wait_afree", "copy", "fence", "bar", "mma_issue", "wait_dfull", "epilogue", "mel", "combine"]
            cyc = [lib.asr_plan_debug_word(ptc._h, 6 + i) * 16 for i in range(10)]
            print("  CTA0 thread0 cycles per phase (last launch):", {n: c for n, c in zip(names, cyc)}, "total", sum(cyc), flush=True)
        print(f"B={B} noisy={noisy} ok maxdiff={float((out - ref).abs().max()):.2e} last call {dt * 1e3:.3f} ms", flush=True)
