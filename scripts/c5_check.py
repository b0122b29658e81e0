import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import asr_b200 as A
from oracle import librosa_ref as lr
from synth import synth_clips, to_f32
clips = synth_clips(3, 160000, 16000, 55, lengths=[160000, 160000, 47111])
for variant in ("mfcc", "logmel", "no_delta"):
    P = A.C5 if variant != "no_delta" else A.C5.replace(delta_orders=0)
    Pr = lr.C5 if variant != "no_delta" else lr.C5.replace(delta_orders=0)
    plan = A.MfccPlan(P, path="clip")
    batch = A.ClipBatch.from_arrays(clips)
    out, st = (plan.logmel(batch, out_frames=1001) if variant == "logmel" else plan.mfcc(batch, out_frames=1001))
    torch.cuda.synchronize()
    for i, c in enumerate(clips):
        x = to_f32([c])[0]
        ref = lr.log_mel(x, Pr) if variant == "logmel" else lr.mfcc(x, Pr)
        t = ref.shape[1]
        err = np.abs(out[i, :, :t].cpu().numpy() - ref)
        r, f = np.unravel_index(err.argmax(), err.shape)
        print(variant, i, "max err %.3e at row %d frame %d (ref %.3f) ; rows with err>2e-3: %s ; frames: %s" % (err.max(), r, f, ref[r, f], np.unique(np.where(err > 2e-3)[0])[:10], np.unique(np.where(err > 2e-3)[1])[:12]))
