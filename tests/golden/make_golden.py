"""Generates tests/golden/*.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

PARITY UNPINNED: the reference ships no golden vectors and librosa is not installable here, so
these fixtures pin the ORACLE (oracle/librosa_ref.py etc.), not the reference.  Inputs are seeded
synthetic clips (tests/synth.py); outputs are float32/float64 exactly as the oracle returns them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import librosa_ref as lr, noise_ref as nr, cmvn_ref as cr, pipeline_ref as pr  # noqa: E402
from synth import synth_clips, to_f32  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {  # name -> (preset, n_clips, length, seed)
    "ref_vdr": ("ref_vdr", 3, 22050, 20240),
    "ref_sr": ("ref_sr", 3, 22050, 20241),
    "c1": ("c1", 4, 16000, 20242),
    "c3": ("c3", 3, 24000, 20243),
    "c5": ("c5", 1, 40000, 20245),
}


def main():
    for name, (preset, n, L, seed) in CASES.items():
        p = lr.PRESETS[preset]
        clips = synth_clips(n, L, p.sr, seed)
        xs = to_f32(clips)
        if preset == "ref_sr":
            xs = [x.astype(np.float64) for x in xs]      # the reference passes float64 windows (SR/extract...py:227)
        feats = np.stack([lr.mfcc(x, p) for x in xs])
        np.savez_compressed(os.path.join(HERE, f"mfcc_{name}.npz"), audio_i16=np.stack(clips), mfcc=feats)
        print(name, feats.shape, feats.dtype, float(np.abs(feats).max()))
    # noise: seeded numpy global RNG, exactly as the reference consumes it
    clips = to_f32(synth_clips(2, 4133, 16000, 20250))
    out = {}
    for snr in (0, 5, 10, 20):
        np.random.seed(1234 + snr)
        out[f"snr{snr}"] = np.stack([nr.add_white_noise_with_snr(c, snr) for c in clips])
    np.random.seed(77)
    out["white"] = np.stack([nr.add_white_noise(c, 0.01) for c in clips])
    np.random.seed(78)
    out["mixture"] = np.stack([nr.add_noise(c, 0.01, 0.002) for c in clips])
    out["power"] = np.array([np.mean(c ** 2) for c in clips], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "noise.npz"), audio_f32=np.stack(clips), **out)
    # standardisation
    rng = np.random.default_rng(20260)
    a, b, c = rng.standard_normal((37, 24)) * 3 + 1, rng.standard_normal((11, 24)), rng.standard_normal((5, 24)) - 2
    a[:, 5] = 0.0; b[:, 5] = 0.0; c[:, 5] = 0.0          # a constant column (zero-padded frames): scale -> 1
    sa, sb, sc = cr.standardize_dataset(a, b, c)
    np.savez_compressed(os.path.join(HERE, "cmvn.npz"), a=a, b=b, c=c, sa=sa, sb=sb, sc=sc)
    print("noise + cmvn written")


if __name__ == "__main__":
    main()
