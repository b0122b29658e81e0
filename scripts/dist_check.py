"""Two-rank check (torchrun): the step with the all-reduces captured inside its CUDA graph gives the same rows as
the three-graph form, and the statistics match a single-process computation over both shards."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200 as A  # noqa: E402
from asr_b200.pipeline import NoisyFeaturePipeline  # noqa: E402
from synth import synth_clips  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
B, L = 512, 16000
clips = np.stack(synth_clips(B, L, 16000, 100 + rank))
batch = A.ClipBatch.from_matrix(torch.from_numpy(clips).to(dev))
z = A.randn(7 + rank, rank * B * L, B * L, device=dev)
T = A.C1.num_frames(L)
outs = {}
for cap in (True, False):
    pipe = NoisyFeaturePipeline(A.C1, T, device=dev, distributed=True, world_size=world)
    pipe.capture_collectives = cap
    for _ in range(3):
        out = pipe.run_device(batch, z, 10.0)
    torch.cuda.synchronize()
    outs[cap] = out.clone()
    n_graphs = len(next(iter(pipe._cache.values())).graphs)
    if rank == 0:
        print(f"capture_collectives={cap}: graphs per step = {n_graphs}", flush=True)
same = torch.equal(outs[True], outs[False])
# global statistics of the standardised rows: mean 0, variance 1 over BOTH shards
s = torch.stack([outs[True].double().sum(0), (outs[True].double() ** 2).sum(0)])
dist.all_reduce(s)
mean = s[0] / (B * world)
var = s[1] / (B * world) - mean ** 2
ok = same and float(mean.abs().max()) < 1e-5 and float((var - 1).abs().max()) < 1e-3
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("identical rows:", same, "max|mean|", float(mean.abs().max()), "max|var-1|", float((var - 1).abs().max()))
    print("DIST CHECK", "OK" if int(flag.item()) == 1 else "FAILED", flush=True)
dist.destroy_process_group()
