"""Phase breakdown of the tensor-core dense DFT kernel on the reference's speaker-corpus shape (1148 float64 windows)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200 as A
from asr_b200._lib import lib
from synth import synth_clips
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1148
clips = [c.astype(np.float64) / 32768.0 for c in synth_clips(B, 22050, 22050, 3)]
batch = A.ClipBatch.from_arrays(clips)
plan = A.MfccPlan(A.REF_SR)
for _ in range(3):
    out, st = plan.mfcc(batch)
torch.cuda.synchronize()
names = ["rows", "gather", "wait_free", "convert", "sync", "issue", "wait_d", "power", "mel", "combine"]
print({n: lib.asr_plan_debug_word(plan._h, 6 + i) * 16 for i, n in enumerate(names)})
