"""The C-ABI library loads and exports every symbol ``include/asr_b200.h`` declares; the ctypes binding
covers the same set; without a CUDA device the entry points fail loudly (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "asr_b200.h")
LIB = os.path.join(ROOT, "asr-using-robust-nn_b200", "libasr_b200.so")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_]*\s*\*?\s*(asr_[a-z0-9_]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import subprocess
        subprocess.run(["make", "-C", os.path.dirname(LIB), "-j4"], check=True)
    return ctypes.CDLL(LIB)


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("asr_plan_create", "asr_plan_destroy", "asr_mfcc_batch", "asr_mfcc_batch_host", "asr_clip_power",
                 "asr_snr_sigma", "asr_snr_sigma_host", "asr_babble_stream", "asr_babble_workspace_bytes", "asr_plan_set_stage_probe", "asr_tc_selftest", "asr_cmvn_workspace_bytes", "asr_cmvn_partial_sums", "asr_cmvn_local_message", "asr_cmvn_merge", "asr_cmvn_apply2", "asr_plan_debug_word", "asr_fp32_peak_probe", "asr_mix_white", "asr_mix_mixture", "asr_randn_f64", "asr_cmvn_colsum",
                 "asr_cmvn_colsum_centered", "asr_cmvn_finalize", "asr_cmvn_apply", "asr_last_error", "asr_version"):
        assert must in names
    assert len(names) >= 24


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_checking_build_is_not_stale():
    """libasr_b200_dbg.so (make debug, loaded by tests/test_tile_debug_gpu.py) exports the same surface as the product library."""
    dbg = os.path.join(os.path.dirname(LIB), "libasr_b200_dbg.so")
    if not os.path.exists(dbg):
        import subprocess
        subprocess.run(["make", "-C", os.path.dirname(LIB), "-j4", "debug"], check=True)
    d = ctypes.CDLL(dbg)
    missing = [n for n in declared_functions() if not hasattr(d, n)]
    assert not missing, f"stale checking build (make -C asr-using-robust-nn_b200 debug): {missing}"


def test_ctypes_binding_matches_header():
    from asr_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_version_and_struct_layout(lib):
    from asr_b200 import _lib
    lib.asr_version.restype = ctypes.c_int
    assert lib.asr_version() == 100
    assert ctypes.sizeof(_lib.MfccParamsC) == 18 * 4          # 12 int32 + 6 float, no padding
    assert ctypes.sizeof(_lib.NoiseC) == 8 + 3 * 8 + 3 * 8


def test_no_cpu_fallback(lib):
    """Without a GPU the plan cannot be created: ASR_ERR_CUDA and a message, never a silent CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from asr_b200 import _lib, params
    h = ctypes.c_void_p()
    pc = params.C1.to_c()
    rc = _lib.lib.asr_plan_create(ctypes.byref(pc), ctypes.byref(h))
    assert rc == -2 and not h.value
    assert b"CUDA" in _lib.lib.asr_last_error()
    import asr_b200 as A
    with pytest.raises(A.AsrError):
        A.MfccPlan(A.C1)
    with pytest.raises(A.AsrError):
        A.ClipBatch.from_arrays([__import__("numpy").zeros(100, "int16")])


def test_invalid_parameters_are_rejected_before_any_cuda_call(lib):
    from asr_b200 import _lib, params
    h = ctypes.c_void_p()
    for bad in (params.C1.replace(n_fft=4), params.C1.replace(n_mfcc=40), params.C1.replace(hop_length=0),
                params.C1.replace(delta_orders=3), params.C1.replace(win_length=1024)):
        pc = bad.to_c()
        assert _lib.lib.asr_plan_create(ctypes.byref(pc), ctypes.byref(h)) == -1
        assert _lib.lib.asr_last_error().startswith(b"asr_plan_create")


def test_copy_mapped_rejects_bad_arguments_before_any_cuda_call(lib):
    """asr_copy_mapped: argument checks come first (no device needed); a zero-byte transfer is a no-op."""
    import torch
    from asr_b200 import _lib
    from asr_b200.frontend import copy_mapped
    f = _lib.lib.asr_copy_mapped
    assert f(None, None, 0, None) == 0
    assert f(None, None, 8, None) == -1 and b"null" in _lib.lib.asr_last_error()
    buf = (ctypes.c_char * 64)()
    base = ctypes.addressof(buf)
    assert f(base + 1, base + 8, 8, None) == -1            # misaligned pointer
    assert f(base, base + 8, 6, None) == -1                # byte count not a multiple of 4
    with pytest.raises(ValueError):
        copy_mapped(torch.zeros(4), torch.zeros(4))         # neither CUDA nor pinned
    with pytest.raises(ValueError):
        copy_mapped(torch.zeros(4), torch.zeros(4, dtype=torch.float64))


def test_c_abi_from_plain_c(tmp_path):
    """include/asr_b200.h is valid strict C99 (-Wall -Wextra -Werror) and the library links and runs from a plain C program
    (examples/mfcc_from_c.c: the reference's compute_mfcc_all_files through asr_mfcc_batch_host).  With a CUDA device it
    prints features (exit 0; `profiles/r2_mfcc_from_c.txt` keeps a B200 run, equal to the oracle to 1.1e-3); without one the
    library reports ASR_ERR_CUDA - no CPU fallback - and the program exits 3."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "mfcc_from_c")
    libdir = os.path.dirname(LIB)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "mfcc_from_c.c"), "-L", libdir, "-lasr_b200", f"-Wl,-rpath,{libdir}", "-lm", "-o", exe],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode in (0, 3), (r.returncode, r.stdout[-1000:], r.stderr[-1000:])
    if r.returncode == 3:
        assert "no CPU fallback" in r.stdout
    else:
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("clip ")]
        assert len(lines) == 4 and all("status 0, 44 frames" in ln for ln in lines)
