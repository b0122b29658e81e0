"""N-rank check (torchrun, NCCL): the sharded step (clips by index, ONE all-gather of the standardisation messages) gives
the rows the single-GPU step gives on the union of the shards.  Rank 0 recomputes the whole batch alone and compares.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 scripts/dist_rows_check.py
Prints max |sharded - single| over all rows (float32 rows; statistics in float64) and the statistics' relative difference."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import asr_b200 as A  # noqa: E402
from asr_b200 import sharding  # noqa: E402
from asr_b200.pipeline import NoisyFeaturePipeline  # noqa: E402
from synth import synth_clips  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
N, L = 1000 * world + 37, 16000                      # uneven shards on purpose
T = A.C1.num_frames(L)
clips_all = np.stack(synth_clips(N, L, 16000, 4242))
lo, hi = sharding.shard_bounds(N, rank, world)
for noisy in (False, True):
    batch = A.ClipBatch.from_matrix(torch.from_numpy(clips_all[lo:hi]).to(dev))
    z = A.randn(77, lo * L, (hi - lo) * L, device=dev) if noisy else None      # the stream is indexed by global sample position
    pipe = NoisyFeaturePipeline(A.C1, T, device=dev, distributed=True, world_size=world)
    for _ in range(2):
        out = pipe.run_device(batch, z, 10.0 if noisy else None)
    torch.cuda.synchronize()
    rows = [torch.empty((sharding.shard_bounds(N, r, world)[1] - sharding.shard_bounds(N, r, world)[0], out.shape[1]),
                        dtype=out.dtype, device=dev) for r in range(world)]
    # gather the sharded rows on every rank (rows of unequal counts: broadcast shard by shard)
    for r in range(world):
        if r == rank:
            rows[r].copy_(out)
        dist.broadcast(rows[r], src=r)
    if rank == 0:
        whole = A.ClipBatch.from_matrix(torch.from_numpy(clips_all).to(dev))
        zw = A.randn(77, 0, N * L, device=dev) if noisy else None
        single = NoisyFeaturePipeline(A.C1, T, device=dev, distributed=False)
        ref = single.run_device(whole, zw, 10.0 if noisy else None)
        torch.cuda.synchronize()
        got = torch.cat(rows, 0)
        d = (got.double() - ref.double()).abs().max().item()
        m1, v1 = single.std.mean.double(), single.std.var.double()
        m2, v2 = pipe.std.mean.double(), pipe.std.var.double()
        rm = ((m1 - m2).abs() / (m1.abs() + 1e-300)).max().item()
        rv = ((v1 - v2).abs() / (v1.abs() + 1e-300)).max().item()
        print(f"world={world} noisy={noisy} clips={N} exchange={pipe.std.exchange_transport} graphs/step={len(next(iter(pipe._cache.values())).graphs)}: max |sharded - single| over rows = {d:.3e}; "
              f"mean rel diff {rm:.2e}, var rel diff {rv:.2e}; rows bit-equal: {bool(torch.equal(got, ref))}", flush=True)
        assert d <= 1e-5 and rv <= 1e-9, "sharded rows differ from the single-GPU rows"
    dist.barrier()
dist.destroy_process_group()
