"""The reference-named entry points: layout, dtype and values of what the Keras scripts consume."""
import os

import numpy as np
import pytest
import scipy.io.wavfile

from synth import synth_clips, to_f32

pytestmark = pytest.mark.gpu


def test_vdr_compute_mfcc_all_files(tmp_path):
    from asr_b200.voice_digit import extract_features_construct_dataset as efcd
    from oracle import pipeline_ref as pr
    clips = synth_clips(5, 0, 22050, 51, lengths=[22050, 22050, 15000, 22050, 30000])
    paths = []
    for i, c in enumerate(clips):
        d = tmp_path / ("zero" if i % 2 == 0 else "one")
        d.mkdir(exist_ok=True)
        p = str(d / f"clip{i}.wav")
        scipy.io.wavfile.write(p, 22050, c)
        paths.append(p)
    flat = efcd.compute_mfcc_all_files(paths)
    assert flat.shape == (5, 880) and flat.dtype == np.float64 and flat.flags["C_CONTIGUOUS"]
    ref = pr.compute_mfcc_all_files(to_f32(clips))
    assert np.abs(flat - ref).max() <= 3e-3
    one = efcd.extract_features(paths[2], 44)
    assert one.shape == (20, 44) and one.dtype == np.float32
    assert np.abs(one.flatten() - ref[2]).max() <= 3e-3
    assert (one[:, 1 + 15000 // 512:] == 0).all()
    names, labels = efcd.get_file_names_and_labels(str(tmp_path))
    assert len(names) == 5 and labels.dtype == np.int32 and set(labels.tolist()) == {0, 1}


def test_sr_load_audio_dataset_and_labels():
    from asr_b200.speaker import extract_features_construct_dataset as efcd
    from asr_b200.speaker import attacks
    from oracle import pipeline_ref as pr
    lengths = [22050 * 5 + 321, 22050 * 2, 22050 * 4, 22050 * 3 + 7]
    waves = to_f32(synth_clips(4, 0, 22050, 52, lengths=lengths))
    labels = np.array([3, 1, 4, 1], dtype=np.int32)
    mfcc, lab = efcd.load_waveforms_and_labels(waves, labels)
    ref, rlab = pr.load_audio_dataset_and_labels(waves, labels)
    assert mfcc.shape == ref.shape == (6, 2020) and mfcc.dtype == np.float64
    assert lab.tolist() == rlab.tolist() == [3, 3, 3, 4, 4, 1]
    assert np.abs(mfcc - ref).max() <= 6e-3
    for snr in (20, 5):
        np.random.seed(70 + snr)
        got, glab = attacks.black_box_attack_on_waveforms_snr(waves, labels, snr)
        np.random.seed(70 + snr)
        zs = [np.random.standard_normal(len(w)) for w in waves]
        ref, rlab = pr.sr_black_box_attack_on_audio_snr(waves, labels, snr, zs)
        assert got.shape == ref.shape and glab.tolist() == rlab.tolist()
        assert np.abs(got - ref).max() <= 6e-3


def test_host_buffer_abi_entry_point():
    """asr_mfcc_batch_host: host int16 in, float64 (N, rows*T) out - what a ctypes stub on the reference side calls."""
    import asr_b200 as A
    from oracle import librosa_ref as lr
    clips = synth_clips(300, 0, 16000, 53, lengths=np.random.default_rng(1).integers(8000, 16001, 300))
    lengths = np.array([len(c) for c in clips], dtype=np.int32)
    offsets, total = A.ClipBatch.layout(lengths)
    audio = np.zeros(total, dtype=np.int16)
    for c, o in zip(clips, offsets):
        audio[o:o + len(c)] = c
    plan = A.MfccPlan(A.C1)
    out, status = plan.mfcc_host(audio, offsets, lengths, out_frames=101)
    assert out.shape == (300, 13 * 101) and out.dtype == np.float64 and (status == 0).all()
    for i in (0, 150, 299):
        ref = lr.mfcc(to_f32([clips[i]])[0], lr.C1)
        T = ref.shape[1]
        got = out[i].reshape(13, 101)
        assert np.abs(got[:, :T] - ref).max() <= 3e-3 and (got[:, T:] == 0).all()
    # SNR mode: the device-generated seeded stream is reproducible and changes the features
    a, _ = plan.mfcc_host(audio, offsets, lengths, out_frames=101, snr_db=10, seed=7)
    b, _ = plan.mfcc_host(audio, offsets, lengths, out_frames=101, snr_db=10, seed=7)
    assert np.array_equal(a, b) and np.abs(a - out).max() > 1


@pytest.mark.gpu
def test_capi_host_call_persistent_slots_chunks_and_growth():
    """asr_mfcc_batch_host keeps its streams / buffers in the plan: a small call, then a larger one that needs several chunks
    (> 4096 clips) and bigger buffers, then the small one again - rows must equal the device-path rows of the same clips."""
    import torch
    import asr_b200 as A
    base = synth_clips(64, 4000, 16000, 11)
    plan = A.MfccPlan(A.C1)

    def run(n):
        clips = [np.roll(base[i % 64], 13 * (i // 64))[: 3000 + 7 * (i % 100)] for i in range(n)]
        lengths = np.array([len(c) for c in clips], dtype=np.int32)
        offsets, total = A.ClipBatch.layout(lengths)
        audio = np.zeros(total, dtype=np.int16)
        for c, o in zip(clips, offsets):
            audio[o:o + len(c)] = c
        out, status = plan.mfcc_host(audio, offsets, lengths, out_frames=30, out_dtype=np.float32)
        dev, st = plan.mfcc(A.ClipBatch.from_arrays(clips), out_frames=30)
        torch.cuda.synchronize()
        assert (status == 0).all() and int(st.max()) == 0
        assert np.array_equal(out, dev.cpu().numpy().reshape(n, -1))

    run(50)
    run(9000)        # three chunks (4096 + 4096 + 808): both slots are reused, pinned descriptor buffers re-filled
    run(50)
