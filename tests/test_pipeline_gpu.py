"""The C2 step as the bench runs it: graph-replayed device pipeline and the three-stream host pipeline."""
import numpy as np
import pytest
import torch

from synth import synth_clips, to_f32

pytestmark = pytest.mark.gpu


def _reference_step(clips_i16, snr, z_host):
    from oracle import librosa_ref as lr, noise_ref as nr, cmvn_ref as cr
    rows = []
    for c, z in zip(to_f32(clips_i16), z_host):
        x = nr.add_white_noise_with_snr_z(c, snr, z) if snr is not None else c
        rows.append(lr.mfcc(np.asarray(x), lr.C1).astype(np.float64).flatten())
    rows = np.stack(rows)
    a, b, c = cr.standardize_dataset(rows[:1], rows[1:2], rows[2:])
    return np.concatenate([a, b, c])


@pytest.mark.parametrize("use_graphs", [False, True])
def test_device_step_matches_oracle(use_graphs):
    import asr_b200 as A
    from asr_b200.pipeline import NoisyFeaturePipeline
    B, L = 24, 16000
    clips = synth_clips(B, L, 16000, 501)
    audio = torch.from_numpy(np.stack(clips)).cuda()
    batch = A.ClipBatch.from_matrix(audio)
    z = A.randn(7, 0, B * L)
    zh = z.cpu().numpy().reshape(B, L)
    pipe = NoisyFeaturePipeline(A.C1, 101, use_graphs=use_graphs)
    for snr in (10, 0, 10):                      # the third call replays the graph captured by the first
        out = pipe.run_device(batch, z, snr).cpu().numpy()
        ref = _reference_step(clips, snr, zh)
        # standardised units: the per-column scale divides the 3e-3 feature tolerance by a std of order 1..10
        assert np.abs(out - ref).max() < 2e-2
        assert np.abs(out - ref).mean() < 2e-4
    out = pipe.run_device(batch, None, None).cpu().numpy()
    ref = _reference_step(clips, None, zh)
    assert np.abs(out - ref).max() < 2e-2


def test_host_pipeline_slots_do_not_race():
    """Consecutive run_host calls with different inputs and SNRs: each result equals the eager device step."""
    import asr_b200 as A
    from asr_b200.pipeline import NoisyFeaturePipeline
    B, L = 64, 16000
    pipe = NoisyFeaturePipeline(A.C1, 101)
    eager = NoisyFeaturePipeline(A.C1, 101, use_graphs=False)
    hosts = [torch.from_numpy(np.stack(synth_clips(B, L, 16000, 600 + i))).pin_memory() for i in range(3)]
    outs = [torch.empty((B, pipe.D), dtype=torch.float32).pin_memory() for _ in range(6)]
    snrs = [0, 5, 10, 20, None, 5]
    for i in range(6):
        pipe.run_host(hosts[i % 3], snrs[i], 99, outs[i], first_index=1000 * i)
    torch.cuda.current_stream().synchronize()
    for i in range(6):
        dev = hosts[i % 3].cuda()
        z = A.randn(99, 1000 * i, B * L) if snrs[i] is not None else None
        want = eager.run_device(A.ClipBatch.from_matrix(dev), z, snrs[i]).cpu()
        assert torch.equal(outs[i], want), i


@pytest.mark.parametrize("kind", ["white", "babble"])
def test_host_pipeline_with_announced_next_batch(kind):
    """run_host(next_audio_host=...): the next call's upload and power pass are issued ahead; announced, unannounced and
    wrongly announced calls in one sequence, clean and noisy - every result equals the eager device step."""
    import asr_b200 as A
    from asr_b200.pipeline import NoisyFeaturePipeline
    B, L = 64, 16000
    pipe = NoisyFeaturePipeline(A.C1, 101)
    eager = NoisyFeaturePipeline(A.C1, 101, use_graphs=False)
    hosts = [torch.from_numpy(np.stack(synth_clips(B, L, 16000, 700 + i))).pin_memory() for i in range(4)]
    seq = [0, 1, 2, 3, 3, 0, 2, 1, 1]
    snrs = [0, 5, 10, None, 20, 5, 0, 10, 5]
    announce = [1, 2, 3, 3, None, 1, 1, 1, None]            # call 5 announces batch 1 but batch 2 follows (a stale announcement)
    outs = [torch.empty((B, pipe.D), dtype=torch.float32).pin_memory() for _ in seq]
    for i, (h, snr, a) in enumerate(zip(seq, snrs, announce)):
        pipe.run_host(hosts[h], snr, 99, outs[i], first_index=1000 * i, noise_kind=kind,
                      next_audio_host=None if a is None else hosts[a])
    torch.cuda.current_stream().synchronize()
    for i, (h, snr) in enumerate(zip(seq, snrs)):
        dev = hosts[h].cuda()
        z = A.randn(99, 1000 * i, B * L) if (snr is not None and kind == "white") else None
        want = eager.run_device(A.ClipBatch.from_matrix(dev), z, snr, noise_kind=kind).cpu()
        assert torch.equal(outs[i], want), (i, h, snr)


def test_corpus_in_batches_equals_one_shot():
    """configs[3] shape: the corpus arrives in batches, statistics are taken once over all rows."""
    import asr_b200 as A
    from asr_b200.pipeline import NoisyFeaturePipeline
    B, L = 40, 16000
    clips = synth_clips(B, L, 16000, 700)
    audio = torch.from_numpy(np.stack(clips)).cuda()
    pipe = NoisyFeaturePipeline(A.C1, 101, use_graphs=False)
    one = pipe.run_device(A.ClipBatch.from_matrix(audio), None, None).clone()
    parts = [(A.ClipBatch.from_matrix(audio[i:i + 16].contiguous()), None, None) for i in range(0, B, 16)]
    many = pipe.run_corpus(iter(parts), B)
    assert torch.equal(one, many)


def test_benchmarked_step_uses_the_reference_sigma_chain():
    """The pipeline bench.py times (graph replay, prefetch of the next step's power) mixes with the reference's own
    sigma: its rows equal an eager step fed with the scalar host chain, bit for bit, at every SNR."""
    import asr_b200 as A
    from asr_b200.pipeline import NoisyFeaturePipeline
    B, L = 96, 16000
    clips = synth_clips(B, L, 16000, 808)
    audio = torch.from_numpy(np.stack(clips)).cuda()
    batch = A.ClipBatch.from_matrix(audio)
    z = A.randn(3, 0, B * L)
    pipe = NoisyFeaturePipeline(A.C1, 101)
    assert pipe.sigma_mode == "host"
    plan = A.MfccPlan(A.C1)
    std = A.Standardizer(pipe.D)
    P = A.clip_power(batch).cpu().numpy()
    for snr in (0, 5, 10, 20, 0, 5):
        got = pipe.run_device(batch, z, snr, prefetch=batch).clone()
        sigma = torch.from_numpy(A.snr_sigma_host_scalar(P, snr)).cuda()
        feats, _ = plan.mfcc(batch, out_frames=101, noise=A.Noise.white(z, sigma))
        flat = feats.reshape(B, -1)
        std.fit([flat])
        want = std.transform(flat, out_dtype=torch.float32)
        assert torch.equal(got, want), snr
    dev = NoisyFeaturePipeline(A.C1, 101, sigma_mode="device")
    out = dev.run_device(batch, z, 10)
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("mode", ["0", "1"])
def test_power_prefetch_modes_give_the_same_rows(mode, monkeypatch):
    """ASR_B200_POWER_STREAM: the prefetched power pass in front of the step (0) or on a side stream behind the step's
    launches (1, the default) - two alternating batches, passes issued two steps ahead, white and babble noise: every
    step's rows equal the rows of a pipeline that never prefetches."""
    import asr_b200 as A
    from asr_b200.pipeline import NoisyFeaturePipeline
    B, L = 64, 16000
    batches, zs = [], []
    for k in range(2):
        audio = torch.from_numpy(np.stack(synth_clips(B, L, 16000, 400 + k))).cuda()
        batches.append(A.ClipBatch.from_matrix(audio))
        zs.append(A.randn(5 + k, 0, B * L))
    monkeypatch.setenv("ASR_B200_POWER_STREAM", "0")
    plain = NoisyFeaturePipeline(A.C1, 101)
    monkeypatch.setenv("ASR_B200_POWER_STREAM", mode)
    pipe = NoisyFeaturePipeline(A.C1, 101)
    assert pipe._pow_mode == int(mode) and (pipe._pow_stream is None) == (mode == "0")
    for kind in ("white", "babble"):
        babble = kind == "babble"
        pipe.prefetch_power(batches[0], babble=babble)
        pipe.prefetch_power(batches[1], babble=babble)
        for i, snr in enumerate((0, 5, 10, 20, 10, 5, 0)):
            b, z = batches[i % 2], zs[i % 2]
            got = pipe.run_device(b, None if babble else z, snr, prefetch=b, noise_kind=kind).clone()
            want = plain.run_device(b, None if babble else z, snr, noise_kind=kind)
            assert torch.equal(got, want), (kind, i, snr)
        pipe.join()
    torch.cuda.synchronize()


def test_babble_step_matches_oracle():
    """The C2 step with babble instead of white noise: features of the oracle's babble mix, standardised."""
    import asr_b200 as A
    from asr_b200.pipeline import NoisyFeaturePipeline
    from oracle import librosa_ref as lr, noise_ref as nr, cmvn_ref as cr
    B, L = 128, 16000
    clips = synth_clips(B, L, 16000, 909)
    audio = torch.from_numpy(np.stack(clips)).cuda()
    batch = A.ClipBatch.from_matrix(audio)
    pipe = NoisyFeaturePipeline(A.C1, 101)
    xs = to_f32(clips)
    for snr in (10, 0):
        out = pipe.run_device(batch, None, snr, noise_kind="babble", prefetch=batch).cpu().numpy()
        rows = np.stack([lr.mfcc(nr.add_babble_with_snr(xs, i, snr), lr.C1).astype(np.float64).flatten() for i in range(B)])
        a, b, c = cr.standardize_dataset(rows[:1], rows[1:2], rows[2:])
        ref = np.concatenate([a, b, c])
        assert np.abs(out - ref).max() < 2e-2 and np.abs(out - ref).mean() < 2e-4
