"""CPU oracle for the MFCC / noise-mix / standardisation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline - never as something the CUDA path routes
through.

PARITY UNPINNED.  The reference (fmazilu/ASR-using-robust-NN) holds no tests,
no golden vectors and no feature matrices (its ``*_data.npy`` are missing LFS
blobs), and the arithmetic of its hot path lives in a third-party package,
``librosa`` (version unpinned by the reference; era evidence points at
0.8.1/0.9.x), which is not installed in the build image and cannot be fetched.
This package therefore RESTATES librosa 0.9 semantics in numpy/scipy
(`librosa_ref`), restates the reference's own pure-numpy noise and
standardisation code verbatim in behaviour (`noise_ref`, `cmvn_ref`), and is
cross-checked against an independent implementation that IS in the image
(``torchaudio.transforms.MFCC`` with librosa-compatible settings, fp32 FFT) in
``tests/test_oracle.py``.  Golden vectors under ``tests/golden/`` are generated
from this oracle by ``tests/golden/make_golden.py`` - they pin the oracle
against regressions, not against the reference.
"""
from . import librosa_ref, noise_ref, cmvn_ref, pipeline_ref  # noqa: F401
