#!/usr/bin/env python
"""bench.py - MFCC clips/sec of the fused sm_100a pipeline (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c5] [--batch B]
  python bench.py --impl reference ...      # the CPU path (oracle port of the reference) on the host cores

A step = one pass of the hot path over one batch of synthetic clips per GPU:
  c2 (default, BASELINE configs[1]): bit-exact clip power -> SNR sigma -> white noise at SNR fused into
      the MFCC launch (C1 front end: 16 kHz, 1 s, 512/400/160 Hamming, 26 mel, 13 MFCC, lifter 22)
      -> dataset standardisation of the (B, 1313) rows (column stats all-reduced over NCCL when N > 1).
  c1 / c3 / c5: the clean MFCC front end of that config (+ standardisation).
Prints ONE JSON line (rank 0).  `value` = inputs resident in HBM; `e2e` = pinned host int16 in,
standardised float32 rows back in pinned host memory, copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SNRS = (0, 5, 10, 20)
METRIC = "MFCC clips/sec (1 s @16 kHz)"
UNIT = "clips/s"
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # 74.45, theoretical non-tensor FP32


# ---- workload definitions -----------------------------------------------------------------------------------
def workload(name):
    """(preset name, clip length, default per-GPU batch, noisy?)"""
    return {"c1": ("c1", 16000, 8192, False), "c2": ("c1", 16000, 8192, True),
            "c3": ("c3", 40000, 2048, False), "c5": ("c5", 160000, 512, False)}[name]


def algorithmic_work(p, L, noisy, in_bytes):
    """bytes and flops per clip."""
    T = p.num_frames(L)
    n_bins = p.n_fft // 2 + 1
    win = p.win_length or p.n_fft
    nnz = 2 * n_bins - 2                              # sparse mel: <= 2 filters per bin
    rows = p.n_mfcc * (1 + p.delta_orders)
    flops_frame = (win + 2.5 * p.n_fft * np.log2(p.n_fft) + 3 * n_bins + 2 * nnz + 3 * p.n_mels
                   + 2 * p.n_mels * p.n_mfcc + p.n_mfcc + p.delta_orders * 18 * p.n_mfcc)
    flops = T * flops_frame + (2 * L if p.preemph else 0) + (5 * L if noisy else 0)
    byts = in_bytes * L + 4 * rows * T + (8 * L if noisy else 0)
    return float(byts), float(flops), T


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(name, batch):
    """dram bytes per launch of the MFCC kernel from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t.get(f"{name}_b{batch}")
        return None if e is None else float(e["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU baseline (oracle port of the reference's per-clip Python loop) ---------------------------------------
def _cpu_worker(args):
    clips_i16, preset, snr, seed = args
    from threadpoolctl import threadpool_limits
    from oracle import librosa_ref as lr, noise_ref as nr
    p = lr.PRESETS[preset]
    np.random.seed(seed)
    rows = []
    with threadpool_limits(limits=1):
        for c in clips_i16:
            x = c.astype(np.float32) / np.float32(32768.0)
            if snr is not None:
                x = nr.add_white_noise_with_snr(x, snr)            # VDR/attacks.py:264
            rows.append(lr.mfcc(x, p).flatten())                   # VDR/attacks.py:267 ; flatten :293
    return np.stack(rows)


def cpu_pass(pool, cores, clips, preset, snr, standardize=True):
    """One pass of the reference CPU path over `clips` on `cores` processes; returns seconds."""
    from oracle import cmvn_ref as cr
    chunks = [clips[i::cores] for i in range(cores)]
    t0 = time.perf_counter()
    parts = pool.map(_cpu_worker, [(ch, preset, snr, 1000 + i) for i, ch in enumerate(chunks) if len(ch)])
    feats = np.concatenate(parts, axis=0).astype(np.float64)
    if standardize:
        cr.standardize_dataset(feats[:1], feats[1:2], feats[2:])   # StandardScaler over all rows (VDR/attacks.py:407)
    return time.perf_counter() - t0


def cpu_baseline(preset, L, sr, snr, budget_s, cores=None, passes=1):
    import multiprocessing as mp
    from synth import synth_clips
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        probe = synth_clips(max(cores * 16, 64), L, sr, 7)
        cpu_pass(pool, cores, probe, preset, snr)                   # warms the workers (imports)
        per_clip = cpu_pass(pool, cores, probe, preset, snr) / len(probe)
        n = int(min(8192, max(cores * 4, budget_s / max(per_clip, 1e-6))))
        clips = synth_clips(n, L, sr, 8)
        times = [cpu_pass(pool, cores, clips, preset, snr) for _ in range(passes)]
    return n, cores, times


# ---- main -------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c5"])
    ap.add_argument("--batch", type=int, default=0, help="clips per GPU per step")
    ap.add_argument("--path", default="auto", choices=["auto", "clip", "frames", "tiles", "tc"], help="kernel path of the MFCC launch (asr_path)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    W = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    K = max(args.steps, 1)
    preset, L, defB, noisy = workload(args.workload)
    B = args.batch or defB
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from oracle import librosa_ref as lr          # parameter presets + the CPU baseline only
    p_or = lr.PRESETS[preset]
    config = {"workload": f"{args.workload}: " + ("C1 front end + white noise at SNR 0/5/10/20 dB + dataset standardisation"
                                                   if noisy else f"{preset} MFCC front end + dataset standardisation"),
              "clip_samples": L, "sr": p_or.sr, "n_fft": p_or.n_fft, "win_length": p_or.win_length or p_or.n_fft,
              "hop_length": p_or.hop_length, "window": p_or.window, "n_mels": p_or.n_mels, "n_mfcc": p_or.n_mfcc,
              "delta_orders": p_or.delta_orders, "lifter": p_or.lifter, "clips_per_gpu_per_step": B,
              "audio_dtype": "int16", "sharding": f"clips by index over {world} rank(s)"}

    # ---------------- reference arm: the CPU path on the host cores (rank 0 only) ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        from synth import synth_clips
        import multiprocessing as mp
        cores = os.cpu_count() or 1
        snr = 10 if noisy else None
        with mp.get_context("fork").Pool(cores) as pool:
            probe = synth_clips(max(cores * 16, 64), L, p_or.sr, 7)
            cpu_pass(pool, cores, probe, preset, snr)                       # warms the workers (imports)
            per_clip = cpu_pass(pool, cores, probe, preset, snr) / len(probe)
            n = int(min(B, max(cores * 2, 2.0 / max(per_clip, 1e-6))))      # ~2 s of all-core work per step
            clips = synth_clips(n, L, p_or.sr, 8)
            for _ in range(W):
                cpu_pass(pool, cores, clips, preset, snr)
            t = [cpu_pass(pool, cores, clips, preset, snr) for _ in range(K)]
        total = float(sum(t))
        val = n * K / total
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (FFT in f64, as librosa)", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n} clips per step (bounded sample of the {B}-clip batch), numpy/scipy "
                                       "restatement of librosa 0.9 called one clip at a time, "
                                       f"{cores} processes, BLAS threads 1"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ---------------- B200 arm ----------------
    import torch
    import asr_b200 as A
    from asr_b200.pipeline import NoisyFeaturePipeline
    from synth import synth_clips
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the asr_b200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    params = A.PRESETS[preset]
    T = params.num_frames(L)
    pipe = NoisyFeaturePipeline(params, T, device=dev, distributed=world > 1, world_size=world, path=args.path)
    D = pipe.D

    # synthetic inputs: 256 distinct seeded clips tiled to the batch, rolled per row so no two rows are equal
    base = np.stack(synth_clips(256, L, params.sr, 20240 + rank))
    reps = (B + 255) // 256
    host = np.concatenate([np.roll(base, 37 * r, axis=1) for r in range(reps)], axis=0)[:B]
    audio_host = torch.from_numpy(np.ascontiguousarray(host)).pin_memory()
    audio_dev = audio_host.to(dev)
    batch = A.ClipBatch.from_matrix(audio_dev)
    z = A.randn(1234 + rank, rank * B * L, B * L, device=dev) if noisy else None
    out_host = torch.empty((B, D), dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()

    def step(i):
        # the next step works on the same resident batch: its power pass is enqueued in front of this step's MFCC launch
        return pipe.run_device(batch, z, SNRS[i % 4] if noisy else None, prefetch=batch if noisy else None)

    def sync_all():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None      # samples through warm-up and the timed region
    for i in range(4 if noisy else 1):                              # setup: capture the step's CUDA graphs (one per SNR)
        step(i)
    for i in range(W):
        step(i)
    sync_all()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(K):
        step(i)
    t_end.record()
    sync_all()
    ms_total = t_start.elapsed_time(t_end)
    if dist is not None:
        tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    # keep the GPU under the same load until nvidia-smi (100 ms period) has samples; same count on every rank
    n_extra = int(max(0.0, 1500.0 - ms_total) / max(ms_total / K, 1e-3))
    for i in range(n_extra):
        step(i)
    sync_all()
    clocks = sampler.stop() if sampler else None
    value = B * world * K / (ms_total * 1e-3)

    # ---- the dominant kernel alone: the same MFCC launch (same buffers, same noise descriptor) K times,
    #      a CUDA event pair around every launch on the launching stream ----
    feats_k = torch.empty((B, pipe.rows, pipe.out_frames), dtype=torch.float32, device=dev)
    noise_k = None
    if noisy:
        noise_k = A.Noise.white(z, torch.from_numpy(A.snr_sigma_host(A.clip_power(batch).cpu().numpy(), 10)).to(dev))
    Kk = max(5, min(K, 50))
    for _ in range(2):
        pipe.plan.mfcc(batch, out_frames=pipe.out_frames, noise=noise_k, out=feats_k)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(Kk)]
    for a_, b_ in ev:
        a_.record()
        pipe.plan.mfcc(batch, out_frames=pipe.out_frames, noise=noise_k, out=feats_k)
        b_.record()
    torch.cuda.synchronize()
    ms_mfcc = sum(a_.elapsed_time(b_) for a_, b_ in ev) / Kk

    # ---------------- e2e: pinned host in -> pinned host out, copies in the timed region ----------------
    e2e = None
    if not args.no_e2e:
        for i in range(8):          # captures the graphs of both buffer sets (the slot alternates with i, the SNR with i % 4)
            pipe.run_host(audio_host, SNRS[i % 4] if noisy else None, 99, out_host, first_index=rank * B * L)
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        Ke = max(3, min(K, 10))
        a.record()
        for i in range(Ke):
            pipe.run_host(audio_host, SNRS[i % 4] if noisy else None, 99, out_host, first_index=rank * B * L)
        b.record()
        sync_all()
        ms_e = a.elapsed_time(b)
        if dist is not None:
            tt = torch.tensor([ms_e], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_e = float(tt.item())
        e2e = {"value": B * world * Ke / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(audio_host.numel() * 2),
               "d2h_bytes_per_step": int(out_host.numel() * 4), "steps": Ke,
               "note": "per GPU bytes; pinned host int16 in, standardised float32 rows back in pinned host memory, every step; "
                       "noise stream generated on the device from the seed; upload / kernels / download of consecutive steps "
                       "overlap on three streams"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    byts, flops, _ = algorithmic_work(params, L, noisy, 2)
    peak, peak_src = measured_peaks()
    gbs = byts * B / (ms_mfcc * 1e-3) / 1e9
    tfl = flops * B / (ms_mfcc * 1e-3) / 1e12
    roofline = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                "traffic": ncu_traffic(args.workload, B), "kernel": {"tc": "asr_mfcc_batch = frame_prefix_kernel + tc512_kernel (dominant, tcgen05) + cepstra_t_kernel",
                                                                   "tiles": "asr_mfcc_batch = frame_prefix_kernel + tile512_kernel (dominant) + cepstra_t_kernel",
                                                                   "frames": "asr_mfcc_batch = frame_prefix_kernel + frames512_kernel (dominant) + cepstra_kernel",
                                                                   "clip": "asr_mfcc_batch = asr::mfcc_kernel"}[pipe.plan.path_used(np.int16, noisy)], "kernel_ms": ms_mfcc,
                "kernel_timing": f"mean of {Kk} asr_mfcc_batch calls (the step's MFCC launches) on the step's buffers, one CUDA event pair "
                                 "per launch, taken right after the timed region (the step itself replays a CUDA graph)",
                "kernel_share_of_step": ms_mfcc / (ms_total / K), "peak_source": peak_src,
                "algorithmic_bytes_per_clip": byts, "algorithmic_flops_per_clip": flops,
                "binding": "fp32 (non-tensor CUDA cores); the HBM fraction is reported because the schema asks for it",
                "fp32": {"achieved": tfl, "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tfl / FP32_PEAK_TFLOPS,
                         "peak_source": "theoretical 148 SM x 128 lanes x 2 x 1.965 GHz"}}
    cpu = None
    if not args.no_cpu_baseline:
        n, cores, times = cpu_baseline(preset, L, params.sr, 10 if noisy else None, budget_s=15.0)
        cpu = {"value": n / times[0], "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} clips of the same workload (numpy/scipy restatement of librosa 0.9 + the reference's "
                         f"noise code, one clip per call, {cores} processes, BLAS threads 1), {times[0]:.1f} s"}
    launches = pipe.launches_per_step(noisy) * K
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": dict(config, l2="inputs per step exceed L2 (audio "
                                                                 f"{audio_host.numel() * 2 >> 20} MiB" +
                                                                 (f" + noise {B * L * 8 >> 20} MiB" if noisy else "") + ")"),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
