"""TC path inside the pipeline (clean), with and without CUDA graphs."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200 as A
from asr_b200.pipeline import NoisyFeaturePipeline
from synth import synth_clips
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
base = np.stack(synth_clips(256, 16000, 16000, 3))
host = np.concatenate([np.roll(base, 37 * r, axis=1) for r in range((B + 255) // 256)], axis=0)[:B]
audio = torch.from_numpy(np.ascontiguousarray(host)).cuda()
batch = A.ClipBatch.from_matrix(audio)
for graphs in (False, True):
    pipe = NoisyFeaturePipeline(A.C1, 101, use_graphs=graphs, path="tc")
    for i in range(8):
        out = pipe.run_device(batch, None, None)
    torch.cuda.synchronize()
    print("graphs", graphs, "ok", float(out.abs().max()), flush=True)
