#!/usr/bin/env python
"""BASELINE.json configs[3]: a synthetic corpus of C1-shaped clips sharded over the GPUs of one box, dataset-level
CMVN statistics all-reduced once over NCCL, then applied.

  python scripts/corpus_c4.py [--clips 1000000] [--batch 8192]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/corpus_c4.py ...

Clips are generated on the device batch by batch (seeded normal stream indexed by global sample position, scaled to
int16), so no host memory or PCIe traffic is involved; generation is outside the timed sections.  Prints one JSON line.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=8192)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import asr_b200 as A
    from asr_b200 import sharding
    from asr_b200.pipeline import NoisyFeaturePipeline
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = 16000
    lo, hi = sharding.shard_bounds(args.clips, rank, world)
    n_local = hi - lo
    pipe = NoisyFeaturePipeline(A.C1, A.C1.num_frames(L), device=dev, distributed=world > 1, world_size=world, use_graphs=False)
    audio = torch.empty((args.batch, L), dtype=torch.int16, device=dev)
    zbuf = torch.empty(args.batch * L, dtype=torch.float64, device=dev)
    ev = []

    def batches():
        for b0 in range(lo, hi, args.batch):
            nb = min(args.batch, hi - b0)
            A.randn(1000, b0 * L, nb * L, device=dev, out=zbuf[:nb * L])       # global sample index: shard-independent corpus
            audio[:nb].copy_((zbuf[:nb * L].view(nb, L) * 3276.7).round_().clamp_(-32768, 32767))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            yield A.ClipBatch.from_matrix(audio[:nb]), None, None
            b.record()
            ev.append((a, b))

    for timed in (False, True):
        ev.clear()
        t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda.synchronize()
        t0.record()
        feats = torch.empty((n_local, pipe.rows, pipe.out_frames), dtype=torch.float32, device=dev)
        done = 0
        for batch, _, _ in batches():
            pipe._group1(batch, None, None, feats[done:done + batch.n_clips])
            done += batch.n_clips
        flat = feats.view(n_local, pipe.D)
        t1.record()
        pipe.std.fit([flat])
        out = pipe.std.transform(flat)
        t2.record()
        torch.cuda.synchronize()
    ms_mfcc = sum(a.elapsed_time(b) for a, b in ev)
    ms_cmvn = t1.elapsed_time(t2)
    tt = torch.tensor([ms_mfcc, ms_cmvn], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(tt[0] + tt[1])
        print(json.dumps({"workload": "c4: synthetic corpus, C1 front end, one dataset-level CMVN all-reduce", "clips": args.clips,
                          "n_gpus": world, "batch": args.batch, "ms_mfcc_max_over_ranks": float(tt[0]),
                          "ms_cmvn_fit_apply_max_over_ranks": float(tt[1]), "clips_per_s": args.clips / (ms * 1e-3),
                          "scaling": "strong", "mean_abs_col_mean": float(out.double().mean(0).abs().mean()),
                          "note": "clip generation on the device is outside the timed sections"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
