"""ctypes binding of ``libasr_b200.so`` (the C-ABI declared in ``include/asr_b200.h``).

The library is built in-tree by ``make -C asr-using-robust-nn_b200`` (or
``__graft_entry__.build()``).  There is no CPU fallback: if the shared library is
missing this module raises at import, and every compute call raises
``AsrError`` when CUDA is unusable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ASR_B200_LIB: another build of the same library (the -DASR_TILE_DEBUG build `make debug` produces); there is still no fallback
LIB_PATH = os.environ.get("ASR_B200_LIB") or os.path.join(_HERE, "libasr_b200.so")

ASR_I16, ASR_F32, ASR_F64 = 0, 1, 2
ASR_NOISE_NONE, ASR_NOISE_WHITE, ASR_NOISE_MIXTURE = 0, 1, 2
ASR_CLIP_OK, ASR_CLIP_TOO_SHORT, ASR_CLIP_TOO_FEW_FRAMES = 0, 1, 2
ASR_PATH_AUTO, ASR_PATH_CLIP, ASR_PATH_FRAMES, ASR_PATH_TILES, ASR_PATH_TC = 0, 1, 2, 3, 4


class AsrError(RuntimeError):
    pass


class MfccParamsC(C.Structure):
    _fields_ = [
        ("sr", C.c_int32), ("n_fft", C.c_int32), ("win_length", C.c_int32), ("hop_length", C.c_int32),
        ("window", C.c_int32), ("center", C.c_int32), ("pad_mode", C.c_int32), ("fftfreq_mode", C.c_int32),
        ("n_mels", C.c_int32), ("n_mfcc", C.c_int32),
        ("fmin", C.c_float), ("fmax", C.c_float), ("top_db", C.c_float), ("amin", C.c_float),
        ("lifter", C.c_float), ("preemph", C.c_float),
        ("delta_orders", C.c_int32), ("delta_width", C.c_int32),
    ]


class NoiseC(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("reserved", C.c_int32),
        ("z_dev", C.c_void_p), ("z2_dev", C.c_void_p), ("sigma_dev", C.c_void_p),
        ("p", C.c_double), ("sigma0", C.c_double), ("sigma1", C.c_double),
    ]


# name -> (restype, argtypes); every symbol include/asr_b200.h declares
_vp, _i32, _i64, _u64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double
SIGNATURES = {
    "asr_version": (C.c_int, []),
    "asr_last_error": (C.c_char_p, []),
    "asr_device_count": (C.c_int, []),
    "asr_plan_create": (C.c_int, [C.POINTER(MfccParamsC), C.POINTER(_vp)]),
    "asr_plan_destroy": (None, [_vp]),
    "asr_plan_num_frames": (_i32, [_vp, _i64]),
    "asr_plan_feature_rows": (_i32, [_vp]),
    "asr_plan_uses_fft": (_i32, [_vp]),
    "asr_plan_get_tables": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "asr_plan_set_stage_probe": (C.c_int, [_vp, _vp]),
    "asr_plan_debug_word": (_i32, [_vp, _i32]),
    "asr_fp32_peak_probe": (C.c_int, [_i32, _i32, _vp, _vp]),
    "asr_mlp_forward": (C.c_int, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "asr_cmvn_workspace_bytes": (C.c_size_t, [_i32]),
    "asr_cmvn_partial_sums": (_i32, [_vp, _i32, _i64, _i32, _i64, C.POINTER(NoiseC), _i32, _i32, _i64, _i32, _vp, C.c_size_t, _vp]),
    "asr_cmvn_local_message": (C.c_int, [_vp, C.c_size_t, _i32, _i32, _i64, _i32, _vp, _vp]),
    "asr_cmvn_merge": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "asr_cmvn_exchange_nccl": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "asr_cmvn_p2p_region_bytes": (C.c_size_t, [_i32, _i32]),
    "asr_cmvn_exchange_p2p": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "asr_cmvn_apply2": (C.c_int, [_vp, _i32, _i64, _i32, _i64, C.POINTER(NoiseC), _vp, C.c_size_t, _i32, _i32, _i64, _vp, _vp, _vp,
                                  _vp, _i32, _vp]),
    "asr_tc_selftest": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "asr_mfcc_batch": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _i32, _i32, C.POINTER(NoiseC), _vp, _i32, _i32, _vp,
                                 _vp, C.c_size_t, _vp]),
    "asr_logmel_batch": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _i32, _i32, C.POINTER(NoiseC), _vp, _i32, _vp,
                                   _vp, C.c_size_t, _vp]),
    "asr_mfcc_workspace_bytes": (C.c_size_t, [_vp, _i32, _i32]),
    "asr_plan_launches": (_i32, [_vp, _i32]),
    "asr_plan_set_path": (C.c_int, [_vp, _i32]),
    "asr_plan_path_used": (_i32, [_vp, _i32, _i32]),
    "asr_clip_power": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _vp, _vp]),
    "asr_copy_mapped": (C.c_int, [_vp, _vp, C.c_size_t, _vp]),
    "asr_snr_sigma": (C.c_int, [_vp, _f32, _vp, _i32, _vp]),
    "asr_snr_sigma_host": (C.c_int, [_vp, _vp, _f32, _vp, _i32]),
    "asr_babble_workspace_bytes": (C.c_size_t, [_i32, _i32]),
    "asr_babble_stream": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, C.c_size_t, _vp]),
    "asr_mix_white": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "asr_mix_mixture": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _vp, _vp, _f64, _f64, _f64, _vp, _vp]),
    "asr_mix_rows_white": (C.c_int, [_vp, _i64, _vp, _f64, _vp, _vp]),
    "asr_mix_rows_mixture": (C.c_int, [_vp, _i64, _vp, _vp, _f64, _f64, _f64, _vp, _vp]),
    "asr_randn_f64": (C.c_int, [_u64, _u64, _i64, _vp, _vp]),
    "asr_cmvn_colsum": (C.c_int, [_vp, _i32, _i64, _i32, _i64, _vp, _vp]),
    "asr_cmvn_mean": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "asr_cmvn_colsum_centered": (C.c_int, [_vp, _i32, _i64, _i32, _i64, _vp, _vp, _vp]),
    "asr_cmvn_finalize": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "asr_cmvn_apply": (C.c_int, [_vp, _i32, _i64, _i32, _i64, _vp, _vp, _vp, _i32, _vp]),
    "asr_resample_out_len": (_i64, [_i64, _i32, _i32]),
    "asr_resample_design": (C.c_int, [_i32, _i32, _f64, _vp, _i32, _vp, _vp, _vp, _vp]),
    "asr_resample_batch": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _vp, _vp, _vp]),
    "asr_mfcc_batch_host": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _i32, _i32, _f32, _u64, _vp, _i32, _i32, _vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise AsrError(
            f"{LIB_PATH} is missing - build it with `make -C {_HERE}` (nvcc, sm_100a). "
            "There is no CPU fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library drift apart
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.asr_last_error()
        raise AsrError(f"{what or 'asr_b200'} failed ({rc}): {msg.decode() if msg else ''}")
