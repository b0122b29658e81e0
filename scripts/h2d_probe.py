"""Host->device copy bandwidth per rank with 1..N ranks copying at once (bare cudaMemcpyAsync from pinned memory).
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/h2d_probe.py
Every rank pins 256 MiB, all ranks start together, 20 copies each; prints one line per rank count (max / min / sum over ranks).
Names the ceiling the end-to-end arm of bench.py runs into at 8 GPUs."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
nbytes = 256 << 20
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
host.fill_(1)
back = torch.empty(nbytes // 6, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
for active in sorted({1, 2, 4, world} & set(range(1, world + 1))):
    for direction in ("h2d", "h2d+d2h"):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        gbs = 0.0
        if rank < active:
            s2 = torch.cuda.Stream()
            d.copy_(host, non_blocking=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                d.copy_(host, non_blocking=True)
                if direction != "h2d":
                    with torch.cuda.stream(s2):
                        back.copy_(d[: nbytes // 6], non_blocking=True)
            torch.cuda.synchronize()
            gbs = 20 * nbytes / (time.perf_counter() - t0) / 1e9
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        if world > 1:
            g = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(g, t)
            vals = [float(x.item()) for x in g][:active]
        else:
            vals = [gbs]
        if rank == 0:
            print(f"ranks copying {active} ({direction}): per-rank H2D GB/s min {min(vals):.1f} max {max(vals):.1f} sum {sum(vals):.1f}", flush=True)
if world > 1:
    dist.destroy_process_group()
