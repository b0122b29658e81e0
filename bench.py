#!/usr/bin/env python
"""bench.py - MFCC clips/sec of the fused sm_100a pipeline (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--batch B] [--path auto|clip|frames|tiles|tc]
  python bench.py --impl reference ...      # the CPU path (oracle port of the reference) on the host cores

A step = one pass of the hot path over one batch of synthetic clips per GPU.  Workloads (BASELINE.json configs,
SURVEY.md 8(d)):
  c2 (default, configs[1]): bit-exact clip power -> the reference's sigma chain -> white noise at SNR 0/5/10/20 dB
      fused into the MFCC launch (C1 front end: 16 kHz, 1 s, 512/400/160 Hamming, 26 mel, 13 MFCC, lifter 22)
      -> dataset standardisation of the (B, 1313) rows (column statistics all-reduced over NCCL when N > 1).
  c2b   the same with babble noise (sum of six other clips of the batch) instead of white noise.
  c1    configs[0]: the clean C1 front end (+ standardisation); `--batch 1024` is the batch BASELINE states.
  c3    configs[2]: 1-4 s utterances (lengths U{16000..64000}, sorted into 0.5 s buckets), 40 mel, 20 MFCC + delta + delta-delta.
  c4    configs[3]: a corpus of --clips synthetic C1 clips sharded over the GPUs, statistics all-reduced ONCE per corpus.
  c5    configs[4]: 10 s clips, 1024-point FFT, 80 mel, 40 MFCC + delta.
  ref_vdr / ref_sr   the reference's own two parameter sets (librosa defaults at 22 050 Hz on float32 audio; n_fft = 441 on
      float64 one-second windows), at the reference's test-set sizes (2 366 clips / 1 148 windows).
Prints ONE JSON line (rank 0).  `value` = inputs resident in HBM; `e2e` = pinned host audio in, standardised float32 rows
back in pinned host memory, copies inside the timed region.
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
FP32_THEORY_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # 74.45: 148 SM x 128 lanes x 2 x 1.965 GHz

# name -> (preset, noise, default clips per GPU per step, audio dtype, what)
WORKLOADS = {
    "c1": ("c1", None, 8192, "int16", "C1 MFCC front end + dataset standardisation"),
    "c2": ("c1", "white", 8192, "int16", "C1 front end + white noise at SNR 0/5/10/20 dB + dataset standardisation"),
    "c2b": ("c1", "babble", 8192, "int16", "C1 front end + babble noise (six other clips) at SNR 0/5/10/20 dB + dataset standardisation"),
    "c3": ("c3", None, 1024, "int16", "C3 front end (1-4 s utterances in 0.5 s buckets, 40 mel, 20 MFCC + delta + delta-delta) + dataset standardisation"),
    "c4": ("c1", None, 8192, "int16", "corpus of C1 clips in batches, column statistics all-reduced once per corpus, then applied"),
    "c5": ("c5", None, 128, "int16", "C5 front end (10 s clips, 1024-point FFT, 80 mel, 40 MFCC + delta) + dataset standardisation"),
    "ref_vdr": ("ref_vdr", None, 2366, "float32", "the reference's digit-corpus parameters (librosa defaults, 22 050 Hz, pad/truncate to 44 frames) + dataset standardisation"),
    "ref_sr": ("ref_sr", None, 1148, "float64", "the reference's speaker-corpus parameters (n_fft = win = 441, hop 220, float64 one-second windows) + dataset standardisation"),
}


def clip_lengths(name, B, sr, seed):
    """Per-clip sample counts of one batch."""
    if name == "c3":
        rng = np.random.default_rng(20243 + seed)
        L = rng.integers(16000, 64001, size=B)
        bucket = (L + 7999) // 8000                               # 0.5 s buckets: clips of a bucket are neighbours
        return L[np.argsort(bucket, kind="stable")].astype(np.int64)
    fixed = {"c5": 160000, "ref_vdr": 22050, "ref_sr": 22050}.get(name, 16000)
    return np.full(B, fixed, dtype=np.int64)


def make_config(name, p, B, world, lengths, clips_total=None):
    preset, noise, _, dt, what = WORKLOADS[name]
    cfg = {"workload": f"{name}: {what}",
           "clip_samples": int(lengths[0]) if len(set(lengths.tolist())) == 1 else f"{int(lengths.min())}..{int(lengths.max())} (mean {float(lengths.mean()):.0f})",
           "sr": p.sr, "n_fft": p.n_fft, "win_length": p.win_length or p.n_fft, "hop_length": p.hop_length, "window": p.window,
           "n_mels": p.n_mels, "n_mfcc": p.n_mfcc, "delta_orders": p.delta_orders, "lifter": p.lifter,
           "clips_per_gpu_per_step": int(B), "audio_dtype": dt, "sharding": f"clips by index over {world} rank(s)"}
    if clips_total is not None:
        cfg["corpus_clips"] = int(clips_total)
    return cfg


def algorithmic_work(p, lengths, noise, in_bytes, out_frames):
    """mean bytes and flops per clip of the batch."""
    n_bins = p.n_fft // 2 + 1
    win = p.win_length or p.n_fft
    nnz = 2 * n_bins - 2                              # sparse mel: <= 2 filters per bin
    rows = p.n_mfcc * (1 + p.delta_orders)
    flops_frame = (win + 2.5 * p.n_fft * np.log2(p.n_fft) + 3 * n_bins + 2 * nnz + 3 * p.n_mels
                   + 2 * p.n_mels * p.n_mfcc + p.n_mfcc + p.delta_orders * 18 * p.n_mfcc)
    T = np.array([p.num_frames(int(L)) for L in np.unique(lengths)])
    Tm = float(np.mean([p.num_frames(int(L)) for L in lengths])) if len(T) > 1 else float(T[0])
    Lm = float(lengths.mean())
    flops = Tm * flops_frame + (2 * Lm if p.preemph else 0) + (5 * Lm if noise else 0)
    byts = in_bytes * Lm + 4 * rows * min(Tm, out_frames)
    if noise == "white":
        byts += 8 * Lm                                # the float64 standard-normal stream is an input
    elif noise == "babble":
        byts += 6 * in_bytes * Lm                     # six other clips of the batch are read per clip
    return float(byts), float(flops), Tm


def measured_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(name, batch):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t.get(f"{name}_b{batch}")
        return None if e is None or e.get("dram_bytes_per_launch") is None else float(e["dram_bytes_per_launch"])
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


# ---- synthetic clips ------------------------------------------------------------------------------------------------
def synth_batch(name, B, sr, seed):
    """List of int16 clips of the workload's lengths: 256 distinct seeded clips, rolled per row so no two rows are equal."""
    from synth import synth_clips
    lengths = clip_lengths(name, B, sr, seed)
    Lmax = int(lengths.max())
    base = synth_clips(min(256, B), Lmax, sr, 20240 + seed)
    return [np.roll(base[i % len(base)], 37 * (i // len(base)))[:int(n)] for i, n in enumerate(lengths)], lengths


def as_dtype(clips_i16, dt):
    if dt == "int16":
        return clips_i16
    f = [(c.astype(np.float32) / np.float32(32768.0)) for c in clips_i16]
    return f if dt == "float32" else [c.astype(np.float64) for c in f]


# ---- CPU baseline (oracle port of the reference's per-clip Python loop) ---------------------------------------
def _cpu_worker(args):
    clips, all_clips, idx, preset, noise, snr, seed, dt, out_frames = args
    from threadpoolctl import threadpool_limits
    from oracle import librosa_ref as lr, noise_ref as nr
    p = lr.PRESETS[preset]
    np.random.seed(seed)
    rows = []
    with threadpool_limits(limits=1):
        for k, c in zip(idx, clips):
            x = c.astype(np.float32) / np.float32(32768.0)
            if dt == "float64":
                x = x.astype(np.float64)
            if noise == "white":
                x = nr.add_white_noise_with_snr(x, snr)            # VDR/attacks.py:264
            elif noise == "babble":
                x = nr.add_babble_with_snr(all_clips, k, snr)
            m = lr.mfcc(x, p)                                      # VDR/attacks.py:267
            T = m.shape[1]
            if T < out_frames:                                     # pad / truncate in the feature domain (VDR/extract...py:33-37)
                m = np.pad(m, ((0, 0), (0, out_frames - T)))
            rows.append(m[:, :out_frames].flatten())               # flatten :149
    return np.stack(rows)


def cpu_pass(pool, cores, clips, preset, noise, snr, dt, out_frames, standardize=True):
    """One pass of the reference CPU path over `clips` on `cores` processes; returns seconds."""
    from oracle import cmvn_ref as cr
    f32 = [c.astype(np.float32) / np.float32(32768.0) for c in clips] if noise == "babble" else None
    idx = np.arange(len(clips))
    t0 = time.perf_counter()
    parts = pool.map(_cpu_worker, [(clips[i::cores], f32, idx[i::cores], preset, noise, snr, 1000 + i, dt, out_frames)
                                   for i in range(cores) if len(clips[i::cores])])
    feats = np.concatenate(parts, axis=0).astype(np.float64)
    if standardize and len(feats) > 2:
        cr.standardize_dataset(feats[:1], feats[1:2], feats[2:])   # StandardScaler over all rows (VDR/attacks.py:407)
    return time.perf_counter() - t0


def cpu_sample(name, preset, sr, noise, dt, out_frames, budget_s, cap):
    """Bounded sample of the workload on all host cores: (clips in the sample, cores, seconds)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(cores) as pool:
        probe, _ = synth_batch(name, max(cores * 4, 32), sr, 7)
        snr = 10 if noise else None
        cpu_pass(pool, cores, probe, preset, noise, snr, dt, out_frames)                      # warms the workers (imports)
        per_clip = cpu_pass(pool, cores, probe, preset, noise, snr, dt, out_frames) / len(probe)
        n = int(min(cap, max(cores * 2, budget_s / max(per_clip, 1e-6))))
        clips, _ = synth_batch(name, n, sr, 8)
        return pool, cores, clips, snr


# ---- main -------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="clips per GPU per step")
    ap.add_argument("--clips", type=int, default=1_000_000, help="c4: corpus size over all GPUs")
    ap.add_argument("--path", default="auto", choices=["auto", "clip", "frames", "tiles", "tc"], help="kernel path of the MFCC launch (asr_path)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    W = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    K = max(args.steps, 1)
    name = args.workload
    preset, noise, defB, dt, _ = WORKLOADS[name]
    B = args.batch or defB
    noisy = noise is not None
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # ---------------- reference arm: the CPU path on the host cores (rank 0 only; nothing of asr_b200 is loaded) ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import librosa_ref as lr
        p = lr.PRESETS[preset]
        lengths = clip_lengths(name, B, p.sr, 0)
        out_frames = 44 if name == "ref_vdr" else max(_frames(p, int(L)) for L in np.unique(lengths))
        config = make_config(name, _P(p), B, world, lengths, args.clips if name == "c4" else None)
        pool_ctx = cpu_sample(name, preset, p.sr, noise, dt, out_frames, budget_s=2.0, cap=B)
        pool, cores, clips, snr = pool_ctx
        import multiprocessing as mp
        with mp.get_context("fork").Pool(cores) as pool:
            for _ in range(W):
                cpu_pass(pool, cores, clips, preset, noise, snr, dt, out_frames)
            t = [cpu_pass(pool, cores, clips, preset, noise, snr, dt, out_frames) for _ in range(K)]
        n = len(clips)
        total = float(sum(t))
        val = n * K / total
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * total / K, "higher_is_better": True, "scaling": "strong" if name == "c4" else "weak", "vs_baseline": None,
            "dtype": "f32 (FFT in f64, as librosa)", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n} clips per step (bounded sample of the {B}-clip batch; SNR 10 dB where the B200 arm "
                                       "cycles 0/5/10/20), numpy/scipy restatement of librosa 0.9 called one clip at a time, "
                                       f"{cores} processes, BLAS threads 1"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ---------------- B200 arm ----------------
    import torch
    import asr_b200 as A
    from asr_b200.pipeline import NoisyFeaturePipeline
    from asr_b200._lib import lib, check
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
    if name == "c4":
        return bench_corpus(args, params, rank, world, dev, dist, W, K)
    clips_i16, lengths = synth_batch(name, B, params.sr, rank)
    out_frames = 44 if name == "ref_vdr" else max(params.num_frames(int(L)) for L in np.unique(lengths))
    config = make_config(name, params, B, world, lengths)
    pipe = NoisyFeaturePipeline(params, out_frames, device=dev, distributed=world > 1, world_size=world, path=args.path)
    D = pipe.D
    clips = as_dtype(clips_i16, dt)
    layout = A.ClipBatch.from_arrays(clips)                         # packed on the device (offsets aligned to 8 elements)
    batch = layout
    n_el = int(batch.audio.shape[0])
    host_np = np.zeros(n_el, dtype=clips[0].dtype)
    for c, o in zip(clips, layout.offsets_host):
        host_np[int(o):int(o) + len(c)] = c
    audio_host = torch.from_numpy(host_np).pin_memory()
    z = A.randn(1234 + rank, rank * n_el, n_el, device=dev) if noise == "white" else None
    out_host = torch.empty((B, D), dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    kind = noise or "white"

    def step(i):
        # later steps work on the same resident batch: one power pass per step is issued behind this step's launches (side stream)
        return pipe.run_device(batch, z, SNRS[i % 4] if noisy else None, prefetch=batch if noisy else None, noise_kind=kind)

    def sync_all():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None      # samples through warm-up and the timed region
    for i in range(4 if noisy else 1):                              # setup: capture the step's CUDA graphs
        step(i)
    if noisy and pipe._pow_stream is not None:
        # the power pass runs on a side stream beside the step's launches: keep it TWO steps ahead, so that the host's
        # sigma chain of a step never waits for a pass that shares the device with the previous step (every step still
        # consumes one pass and issues one)
        pipe.prefetch_power(batch, babble=kind == "babble")
    for i in range(W):
        step(i)
    sync_all()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(K):
        step(i)
    pipe.join()                                                     # the last step's side-stream power pass ends inside the timed region
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

    # ---- the dominant kernel alone: the same MFCC launch (same buffers, same noise descriptor) Kk times,
    #      a CUDA event pair around every launch on the launching stream ----
    feats_k = torch.empty((B, pipe.rows, pipe.out_frames), dtype=torch.float32, device=dev)
    noise_k = None
    if noisy:
        sig = A.snr_sigma_host(A.clip_power(batch).cpu().numpy(), 10)
        zk = z
        if noise == "babble":
            zk, pb = A.babble_stream(batch)
            sig = A.babble_gain_host(sig, pb.cpu().numpy())
        noise_k = A.Noise.white(zk, torch.from_numpy(sig).to(dev))
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

    # ---- measured FP32 FMA peak of this device ----
    fp32_peak, fp32_src = FP32_THEORY_TFLOPS, "theoretical 148 SM x 128 lanes x 2 x 1.965 GHz"
    if rank == 0:
        try:
            nb, it = 148 * 8, 4096
            sink = torch.empty(nb * 512, dtype=torch.float32, device=dev)
            best = 1e9
            for r in range(4):
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
                check(lib.asr_fp32_peak_probe(nb, it, sink.data_ptr(), torch.cuda.current_stream().cuda_stream), "asr_fp32_peak_probe")
                b_.record()
                torch.cuda.synchronize()
                if r:
                    best = min(best, a_.elapsed_time(b_))
            fp32_peak = nb * 512 * it * 64 * 2 / (best * 1e-3) / 1e12
            fp32_src = f"measured: FMA-chain probe (asr_fp32_peak_probe, {nb} x 512 threads, 8 chains, best of 3) in this run"
        except Exception as e:                                       # noqa: BLE001
            fp32_src += f" (probe failed: {e})"

    # ---------------- e2e: pinned host in -> pinned host out, copies in the timed region ----------------
    e2e = None
    if not args.no_e2e:
        equal = len(set(lengths.tolist())) == 1
        host_in = audio_host[:B * int(lengths[0])].view(B, int(lengths[0])) if equal and n_el == B * int(lengths[0]) else audio_host
        lay = None if host_in.dim() == 2 else layout

        def e2e_step(i):
            # every step uploads the batch from pinned host memory; the NEXT step's batch (here: the same host buffer) is announced,
            # so its upload is enqueued before this step waits for its own sigma and the copy engine never idles between steps
            pipe.run_host(host_in, SNRS[i % 4] if noisy else None, 99, out_host, first_index=rank * n_el, layout=lay, noise_kind=kind,
                          next_audio_host=host_in)

        for i in range(8):          # captures the graphs of both buffer sets (the slot alternates with i, the SNR with i % 4)
            e2e_step(i)
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        Ke = max(3, min(K, 10))
        a.record()
        for i in range(Ke):
            e2e_step(i)
        pipe.join()                  # the upload the last step issued ahead ends inside the timed region: Ke uploads, Ke steps, Ke downloads
        b.record()
        sync_all()
        ms_e = a.elapsed_time(b)
        if dist is not None:
            tt = torch.tensor([ms_e], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_e = float(tt.item())
        # the C-ABI host-buffer entry point INTEGRATION.md tells a maintainer to bind (asr_mfcc_batch_host: MFCC [+ SNR noise
        # with the device sigma chain], features back on the host, no standardisation): a blocking call, host clock
        capi = None
        if noise in (None, "white") and dt != "float64" or not noisy:
            try:
                offs = np.asarray(layout.offsets_host, dtype=np.int64)
                lens = np.asarray(lengths, dtype=np.int32)
                feats_host = torch.empty((B, pipe.rows * pipe.out_frames), dtype=torch.float32).pin_memory().numpy()
                snr_c = 10.0 if noisy else None
                for _ in range(2):
                    pipe.plan.mfcc_host(audio_host.numpy(), offs, lens, pipe.out_frames, snr_db=snr_c, seed=99, out=feats_host)
                t0 = time.perf_counter()
                for _ in range(Ke):
                    pipe.plan.mfcc_host(audio_host.numpy(), offs, lens, pipe.out_frames, snr_db=snr_c, seed=99, out=feats_host)
                capi = {"value": B * Ke / (time.perf_counter() - t0), "unit": UNIT, "scope": "this rank; asr_mfcc_batch_host, pinned host audio in, "
                        "float32 feature rows out to pinned host memory, blocking call timed on the host clock, chunks double-buffered inside the call"}
            except Exception as e:                                   # noqa: BLE001
                capi = {"error": str(e).splitlines()[0][:120]}
        e2e = {"value": B * world * Ke / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(audio_host.numel() * audio_host.element_size()),
               "d2h_bytes_per_step": int(out_host.numel() * 4), "steps": Ke, "capi_host_call": capi,
               "note": f"per GPU bytes; pinned host {dt} in, standardised float32 rows back in pinned host memory, every step; "
                       "noise generated on the device (white: from the seed; babble: from the batch); upload / kernels / download "
                       "of consecutive steps overlap on three streams; the next step's host buffer is announced to each call "
                       "(run_host(next_audio_host=...)), so its upload is enqueued before the call waits for its own sigma"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    in_bytes = {"int16": 2, "float32": 4, "float64": 8}[dt]
    byts, flops, _ = algorithmic_work(params, lengths, noise, in_bytes, out_frames)
    hbm_peak, hbm_src = measured_hbm()
    gbs = byts * B / (ms_mfcc * 1e-3) / 1e9
    tfl = flops * B / (ms_mfcc * 1e-3) / 1e12
    t_hbm, t_fp32 = byts * B / (hbm_peak * 1e9), flops * B / (fp32_peak * 1e12)      # the two bound times of one launch
    binding = "hbm" if t_hbm >= t_fp32 else "fp32"
    path = pipe.plan.path_used({"int16": np.int16, "float32": np.float32, "float64": np.float64}[dt], noisy)
    kernels = {"tc": ("asr_mfcc_batch = frame_prefix_kernel + tc512_kernel (dominant, tcgen05) + cepstra_t_kernel" if params.n_fft == 512 else
                      "asr_mfcc_batch = frame_prefix_kernel + tcdft_kernel (dominant, tcgen05 dense DFT) + cepstra_t_kernel"),
               "tiles": "asr_mfcc_batch = frame_prefix_kernel + tile512_kernel (dominant) + cepstra_t_kernel",
               "frames": "asr_mfcc_batch = frame_prefix_kernel + frames512_kernel (dominant) + cepstra_kernel",
               "clip": "asr_mfcc_batch = asr::mfcc_kernel"}
    hbm_part = {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "peak_source": hbm_src}
    fp32_part = {"achieved": tfl, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tfl / fp32_peak, "peak_source": fp32_src,
                 "theoretical_peak": FP32_THEORY_TFLOPS}
    top = hbm_part if binding == "hbm" else fp32_part
    roofline = {"bound": "hbm" if binding == "hbm" else "fp32 (non-tensor CUDA cores; the schema's 'tensor' does not apply)",
                "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"], "frac": top["frac"],
                "traffic": ncu_traffic(name, B), "kernel": kernels[path], "kernel_ms": ms_mfcc,
                "binding_rule": f"the longer of the two bound times of one launch: HBM {t_hbm * 1e3:.4f} ms, FP32 {t_fp32 * 1e3:.4f} ms",
                "kernel_timing": f"mean of {Kk} asr_mfcc_batch calls (the step's MFCC launches) on the step's buffers, one CUDA event pair "
                                 "per call, taken right after the timed region (the step itself replays a CUDA graph)",
                "kernel_share_of_step": ms_mfcc / (ms_total / K), "peak_source": top["peak_source"],
                "algorithmic_bytes_per_clip": byts, "algorithmic_flops_per_clip": flops, "hbm": hbm_part, "fp32": fp32_part}
    cpu = None
    if not args.no_cpu_baseline:
        pool, cores, cclips, snr = cpu_sample(name, preset, params.sr, noise, dt, out_frames, budget_s=15.0, cap=8192)
        import multiprocessing as mp
        with mp.get_context("fork").Pool(cores) as pool:
            cpu_pass(pool, cores, cclips[:cores * 2], preset, noise, snr, dt, out_frames)
            secs = cpu_pass(pool, cores, cclips, preset, noise, snr, dt, out_frames)
        cpu = {"value": len(cclips) / secs, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{len(cclips)} clips of the same workload (numpy/scipy restatement of librosa 0.9 + the reference's "
                         f"noise code at SNR 10 dB, one clip per call, {cores} processes, BLAS threads 1), {secs:.1f} s"}
    launches = pipe.launches_per_step(noisy) * K
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "l2": f"inputs per step exceed L2 or are rewritten between steps (audio {n_el * in_bytes >> 20} MiB"
                  + (f" + noise {n_el * 8 >> 20} MiB" if noisy else "") + f", features {B * D * 4 >> 20} MiB, log-mel workspace)",
            "sigma": "reference chain on the host (numpy log10 + libm powf), bit-exact; float64 two-rounding mix" if noisy else None,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


class _P:
    """Preset view with the fields make_config reads (the reference arm builds its config from the oracle's preset)."""
    def __init__(self, p):
        self.__dict__.update({k: getattr(p, k) for k in ("sr", "n_fft", "win_length", "hop_length", "window", "n_mels", "n_mfcc",
                                                         "delta_orders", "lifter")})


def _frames(p, L):
    pad = p.n_fft // 2 if p.center else 0
    return 0 if L + 2 * pad < p.n_fft else 1 + (L + 2 * pad - p.n_fft) // p.hop_length


def bench_corpus(args, params, rank, world, dev, dist, W, K):
    """configs[3]: this rank's shard of a --clips corpus in batches; statistics all-reduced once per corpus, then applied.
    Total work is fixed (strong scaling).  A step = one pass over the corpus."""
    import torch
    import asr_b200 as A
    from asr_b200 import sharding
    from asr_b200.pipeline import NoisyFeaturePipeline
    L, Bb = 16000, args.batch or 8192
    lo, hi = sharding.shard_bounds(args.clips, rank, world)
    n_local = hi - lo
    pipe = NoisyFeaturePipeline(params, params.num_frames(L), device=dev, distributed=world > 1, world_size=world,
                                use_graphs=False, path=args.path)
    # the corpus is generated on the device once (seeded normal stream indexed by global sample position -> int16): the
    # shard stays resident (1 M clips = 32 GB of int16 on one GPU) and every pass reads it from HBM
    audio = torch.empty((n_local, L), dtype=torch.int16, device=dev)
    zbuf = torch.empty(Bb * L, dtype=torch.float64, device=dev)
    for b0 in range(0, n_local, Bb):
        nb = min(Bb, n_local - b0)
        A.randn(1000, (lo + b0) * L, nb * L, device=dev, out=zbuf[:nb * L])
        audio[b0:b0 + nb].copy_((zbuf[:nb * L].view(nb, L) * 3276.7).round_().clamp_(-32768, 32767))
    del zbuf

    def batches():
        for b0 in range(0, n_local, Bb):
            yield A.ClipBatch.from_matrix(audio[b0:b0 + min(Bb, n_local - b0)]), None, None

    def sync_all():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    Kc = max(1, min(K, 3))
    sampler = ClockSampler(dev.index) if rank == 0 else None
    for _ in range(min(W, 1) or 1):
        out = pipe.run_corpus(batches(), n_local)
    sync_all()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(Kc):
        out = pipe.run_corpus(batches(), n_local)
    b.record()
    sync_all()
    ms = a.elapsed_time(b)
    if dist is not None:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    clocks = sampler.stop() if sampler else None
    chk = float(out[:8].double().abs().sum().item())
    if rank == 0:
        lengths = np.full(Bb, L, dtype=np.int64)
        config = make_config("c4", params, Bb, world, lengths, args.clips)
        value = args.clips * Kc / (ms * 1e-3)
        byts, flops, _ = algorithmic_work(params, lengths, None, 2, params.num_frames(L))
        hbm_peak, hbm_src = measured_hbm()
        gbs = byts * args.clips / world * Kc / (ms * 1e-3) / 1e9
        print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": Kc, "warmup": 1,
                          "ms_per_step": ms / Kc, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic (generated on the device, resident in HBM)", "config": config,
                          "l2": f"the shard ({n_local * L * 2 >> 20} MiB of int16) exceeds L2",
                          "clocks": clocks, "e2e": None, "gpu_launches": (pipe.plan.launches(False) * ((n_local + Bb - 1) // Bb) + 7) * Kc,
                          "roofline": {"bound": "fp32", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                       "traffic": None, "note": "whole corpus pass (MFCC batches + one statistics all-reduce + apply); "
                                                                "HBM fraction of the pass, the MFCC launches are the c1 kernel", "peak_source": hbm_src},
                          "cpu_baseline": None, "checksum_first_rows": chk}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
