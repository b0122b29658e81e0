"""The tile kernel's staging protocol under its device-side checks: the -DASR_TILE_DEBUG build (make debug ->
libasr_b200_dbg.so) traps on a raw-buffer / sample-buffer overrun, on a mismatch between the bytes announced to the
mbarrier and the bytes the bulk copies deliver, and on a descriptor outside its CTA's frame range.  compute-sanitizer is
closed on the GPU pool, so this is the memory-safety evidence for the TMA / mbarrier path: the tile tests (ragged clips,
all dtypes, noise, edge frames, many CTAs) are run once more in a subprocess against the checking build."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DBG = os.path.join(ROOT, "asr-using-robust-nn_b200", "libasr_b200_dbg.so")


def test_tile_tests_pass_under_the_checking_build():
    if not os.path.exists(DBG):
        pytest.skip("libasr_b200_dbg.so not built (make -C asr-using-robust-nn_b200 debug)")
    env = dict(os.environ, ASR_B200_LIB=DBG)
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", os.path.join(ROOT, "tests", "test_tiles_gpu.py"),
                        os.path.join(ROOT, "tests", "test_pipeline_gpu.py")],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, tail
    assert "check failed" not in r.stdout + r.stderr, tail
    assert " passed" in r.stdout, tail
