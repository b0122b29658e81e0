"""asr_mfcc_batch_host (the C-ABI host-buffer call INTEGRATION.md binds) on 8192 one-second int16 clips from pinned memory,
white noise at SNR 10 (device sigma chain), float32 rows back to pinned memory: ms per blocking call on the host clock.
ASR_B200_HOST_CHUNK_MIB (read once per process) sets the chunk size.  Usage: ASR_B200_HOST_CHUNK_MIB=32 python scripts/host_call_time.py"""
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
    from synth import synth_clips
    B, L = 8192, 16000
    base = np.stack(synth_clips(256, L, 16000, 5))
    audio = torch.from_numpy(np.tile(base, (B // 256, 1)).reshape(-1)).pin_memory().numpy()
    offs = np.arange(B, dtype=np.int64) * L
    lens = np.full(B, L, dtype=np.int32)
    plan = A.MfccPlan(A.C1)
    out = torch.empty((B, 13 * 101), dtype=torch.float32).pin_memory().numpy()
    for snr in (None, 10.0):
        for _ in range(3):
            plan.mfcc_host(audio, offs, lens, 101, snr_db=snr, seed=99, out=out)
        t0 = time.perf_counter()
        n = 10
        for _ in range(n):
            plan.mfcc_host(audio, offs, lens, 101, snr_db=snr, seed=99, out=out)
        ms = (time.perf_counter() - t0) / n * 1e3
        print(f"ASR_B200_HOST_CHUNK_MIB={os.environ.get('ASR_B200_HOST_CHUNK_MIB', 'default')} snr={snr}: {ms:.3f} ms per call = {B / ms / 1e3:.3f} M clips/s")


if __name__ == "__main__":
    main()
