"""Where does the end-to-end step (pinned host audio in -> standardised rows in pinned host memory) spend its time?
CUDA events around every upload (copy stream), every step's kernels (compute stream) and every download of
`NoisyFeaturePipeline.run_host` on the default workload (8192 one-second int16 clips, white noise): prints, per step, the
duration of the three and the idle gap of the copy engine between consecutive uploads, with and without the next batch
announced (`next_audio_host=`).  Usage: python scripts/e2e_timeline.py > gpurun_out/e2e_timeline.txt"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import asr_b200 as A
    from asr_b200.pipeline import NoisyFeaturePipeline
    from synth import synth_clips
    B, L, N = 8192, 16000, 12
    base = np.stack(synth_clips(256, L, 16000, 5))
    host = torch.from_numpy(np.tile(base, (B // 256, 1))).pin_memory()
    out_host = torch.empty((B, 13 * 101), dtype=torch.float32).pin_memory()
    snrs = (0, 5, 10, 20)
    # bare copy of the same buffer, nothing else running
    dev = torch.empty(host.shape, dtype=host.dtype, device="cuda")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(5):
        dev.copy_(host, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"bare upload of {host.numel() * 2 / 1e6:.0f} MB: {ms:.3f} ms = {host.numel() * 2 / ms / 1e6:.1f} GB/s")
    del dev

    for announce in (False, True):
        pipe = NoisyFeaturePipeline(A.C1, 101)
        ups, comps = [], []
        orig_upload = pipe._upload

        def upload(sl, audio_host, _orig=orig_upload, _pipe=pipe):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            _pipe._s_h2d.wait_event(sl["computed"])
            e0.record(_pipe._s_h2d)
            _orig(sl, audio_host)
            e1.record(_pipe._s_h2d)
            ups.append((e0, e1))

        pipe._upload = upload
        orig_run_device = pipe.run_device
        hostlog = []

        def run_device(*a_, _orig=orig_run_device, **kw):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()                                  # compute stream: behind this step's randn launch
            h0 = time.perf_counter()
            r = _orig(*a_, **kw)
            h1 = time.perf_counter()
            c1.record()
            comps.append((c0, c1))
            hostlog.append((h0, h1))
            return r

        pipe.run_device = run_device
        calls, downs = [], []

        def step(i):
            t_in = time.perf_counter()
            pipe.run_host(host, snrs[i % 4], 99, out_host, next_audio_host=host if announce else None)
            d1 = torch.cuda.Event(enable_timing=True)
            d1.record(pipe._s_d2h)
            downs.append(d1)
            calls.append((t_in, time.perf_counter()))

        for i in range(8):
            step(i)
        torch.cuda.synchronize()
        ups.clear(); comps.clear(); hostlog.clear(); calls.clear(); downs.clear()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(N):
            step(i)
        t1.record()
        torch.cuda.synchronize()
        total = t0.elapsed_time(t1) / N
        durs = [e0.elapsed_time(e1) for e0, e1 in ups]
        gaps = [ups[i][1].elapsed_time(ups[i + 1][0]) for i in range(len(ups) - 1)]
        print(f"announce={announce}: {total:.3f} ms per step = {B / total / 1e3:.3f} M clips/s; uploads {len(ups)}: "
              f"mean {np.mean(durs):.3f} ms (min {min(durs):.3f}, max {max(durs):.3f}); copy-engine idle between uploads: "
              f"mean {np.mean(gaps):.3f} ms (min {min(gaps):.3f}, max {max(gaps):.3f})")
        # timeline of a middle step, relative to the start of the first upload issued in the timed region (ms)
        ref = ups[0][0]
        for j in range(4, 8):
            u = ups[j] if j < len(ups) else None
            c = comps[j]
            print(f"  call {j}: upload[{j}] {ref.elapsed_time(u[0]):8.3f} .. {ref.elapsed_time(u[1]):8.3f} | kernels of call {j} (behind randn) "
                  f"{ref.elapsed_time(c[0]):8.3f} .. {ref.elapsed_time(c[1]):8.3f} | download done {ref.elapsed_time(downs[j]):8.3f} | host: call "
                  f"{(calls[j][0] - calls[4][0]) * 1e3:8.3f} .. {(calls[j][1] - calls[4][0]) * 1e3:8.3f}, inside run_device {(hostlog[j][1] - hostlog[j][0]) * 1e3:6.3f} ms")
        del pipe


if __name__ == "__main__":
    main()
