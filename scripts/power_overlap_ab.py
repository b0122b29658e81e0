"""Where should the power pass of the next batch run?  A/B of ASR_B200_POWER_STREAM on the default step (C2: 8192 one-second
int16 clips, white noise at SNR 0/5/10/20, standardisation), one process, one GPU:

  0  asr_clip_power on the caller's stream in front of the step's MFCC launch
  1  the same kernel on a side stream behind the step's launches (it gets SMs when the persistent MFCC kernel ends)

ASR_B200_POW_WARPS = 12 | 4 (read by the library at its first power launch, so one process per shape) selects the CTA shape
of the power kernel.  Prints ms per step (CUDA events around K steps, passes kept two steps ahead), the stand-alone time of
the power kernel and checks that the rows of both modes are bit-equal.
Usage: ASR_B200_POW_WARPS=4 python scripts/power_overlap_ab.py [K] >> gpurun_out/power_ab.txt
(profiles/r2_power_overlap_ab.txt also keeps the run of the rejected "background" launch - a 2-warp, 64-register CTA shaped
to sit beside the persistent MFCC kernel: modes 2 / 3 there.)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import asr_b200 as A
    from asr_b200.pipeline import NoisyFeaturePipeline
    from synth import synth_clips
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    B, L = 8192, 16000
    base = np.stack(synth_clips(256, L, 16000, 5))
    audio = torch.from_numpy(np.tile(base, (B // 256, 1))).cuda()
    audio += torch.randint(-3, 4, audio.shape, dtype=torch.int16, device="cuda")     # every clip its own power
    batch = A.ClipBatch.from_matrix(audio)
    z = A.randn(7, 0, B * L)
    snrs = (0, 5, 10, 20)

    def timed(fn, n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    out = torch.empty(B, dtype=torch.float32, device="cuda")
    A.clip_power(batch, out=out)
    print(f"ASR_B200_POW_WARPS={os.environ.get('ASR_B200_POW_WARPS', 'default')}: stand-alone asr_clip_power "
          f"{timed(lambda i: A.clip_power(batch, out=out), 50):.4f} ms")

    rows = {}
    for clean in (False,):
        for mode in ("0", "1", "0", "1"):
            os.environ["ASR_B200_POWER_STREAM"] = mode
            pipe = NoisyFeaturePipeline(A.C1, 101)

            def step(i):
                return pipe.run_device(batch, z, snrs[i % 4], prefetch=batch)

            for i in range(4):
                step(i)
            if pipe._pow_stream is not None:
                pipe.prefetch_power(batch)
            for i in range(8):
                step(i)

            def region(n):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                for i in range(n):
                    step(i)
                pipe.join()
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b) / n

            ms = [region(K) for _ in range(3)]
            got = step(2).clone()
            torch.cuda.synchronize()
            if "ref" not in rows:
                rows["ref"] = got
            same = torch.equal(rows["ref"], got)
            print(f"ASR_B200_POWER_STREAM={mode}: {min(ms):.4f} ms per step (runs: {', '.join(f'{m:.4f}' for m in ms)}), "
                  f"{B / min(ms) / 1e3:.2f} M clips/s, rows bit-equal to mode 0: {same}")
            assert same
            del pipe


if __name__ == "__main__":
    main()
