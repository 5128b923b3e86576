"""ctypes binding of libmafe.so (include/mafe.h).  No torch, no CUDA call at import time.

The product path has NO CPU fallback: if the library is missing (or there is no B200),
every op raises ``MafeError`` loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmafe.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "mafe.h")

# ---- constants mirrored from include/mafe.h ----
OK = 0
E_UNSUPPORTED = -4
PAD = {"constant": 0, "reflect": 1, "edge": 2, "symmetric": 3}
OUT_COMPLEX, OUT_POWER, OUT_MEL, OUT_LOGMEL, OUT_MFCC = range(5)
LOG_NONE, LOG_LN_EPS_IF_ZERO, LOG_LN_PLUS, LOG_DB = range(4)
WAVE_F32, WAVE_I16 = 0, 1
DBGROUP_NONE, DBGROUP_UTT, DBGROUP_BATCH, DBGROUP_MAP = range(4)
PROF_FBANK_MAIN, PROF_FRAME_MEAN, PROF_CMVN, PROF_OTHER = range(4)


class MafeError(RuntimeError):
    pass


class FrontendDesc(C.Structure):
    _fields_ = [
        ("n_fft", C.c_int32), ("frame_len", C.c_int32), ("hop", C.c_int32), ("center", C.c_int32),
        ("pad_mode", C.c_int32), ("out_kind", C.c_int32), ("window", C.POINTER(C.c_float)),
        ("preemph", C.c_double), ("remove_frame_mean", C.c_int32), ("dither", C.c_float),
        ("dither_seed", C.c_uint64),
        ("power", C.c_float), ("spec_scale", C.c_float),
        ("n_mels", C.c_int32), ("mel_fb", C.POINTER(C.c_float)), ("log_kind", C.c_int32),
        ("log_arg", C.c_float), ("log_mult", C.c_float), ("log_offset", C.c_float), ("top_db", C.c_float),
        ("n_mfcc", C.c_int32), ("dct", C.POINTER(C.c_float)),
        ("utt_cmvn_mean", C.c_int32), ("utt_cmvn_std", C.c_int32),
        ("allow_fast_path", C.c_int32),
        ("utt_scalar_norm", C.c_int32),
    ]


_P = C.c_void_p
_I32, _I64, _F = C.c_int32, C.c_int64, C.c_float
_PROTOS = {
    "mafe_version": (C.c_int, []),
    "mafe_last_error": (C.c_char_p, []),
    "mafe_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "mafe_ctx_destroy": (C.c_int, [_P]),
    "mafe_ctx_set_stream": (C.c_int, [_P, _P]),
    "mafe_ctx_sync": (C.c_int, [_P]),
    "mafe_ctx_sm_count": (C.c_int, [_P]),
    "mafe_ctx_launch_count": (_I64, [_P]),
    "mafe_ctx_profile_enable": (C.c_int, [_P, _I32]),
    "mafe_ctx_profile_read": (C.c_int, [_P, _I32, C.POINTER(C.c_double), C.POINTER(_I64)]),
    "mafe_ctx_profile_reset": (C.c_int, [_P]),
    "mafe_fp32_fma_peak": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "mafe_device_malloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "mafe_device_free": (C.c_int, [_P, _P]),
    "mafe_pinned_malloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "mafe_pinned_free": (C.c_int, [_P, _P]),
    "mafe_memcpy_h2d": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "mafe_memcpy_d2h": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "mafe_memset": (C.c_int, [_P, _P, C.c_int, C.c_size_t]),
    "mafe_memcpy_h2d_gather": (C.c_int, [_P, _P, _P, _P, _I32]),
    "mafe_memcpy_d2h_staged": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "mafe_plan_create": (C.c_int, [_P, C.POINTER(FrontendDesc), C.POINTER(_P)]),
    "mafe_plan_destroy": (C.c_int, [_P]),
    "mafe_plan_num_frames": (_I64, [_P, _I64]),
    "mafe_plan_out_dim": (_I32, [_P]),
    "mafe_plan_is_fast": (_I32, [_P]),
    "mafe_batch_create": (C.c_int, [_P, _P, _P, _I32, _P, C.POINTER(_P)]),
    "mafe_batch_destroy": (C.c_int, [_P]),
    "mafe_batch_refill": (C.c_int, [_P, _P, _P, _P, _I32, _P]),
    "mafe_batch_total_frames": (_I64, [_P]),
    "mafe_batch_total_samples": (_I64, [_P]),
    "mafe_batch_frame_offsets": (C.c_int, [_P, _P]),
    "mafe_batch_frame_offsets_dev": (_P, [_P]),
    "mafe_frontend_run": (C.c_int, [_P, _P, _P, _P, _I32, _F, _P, _I32]),
    "mafe_frontend_run_aux": (C.c_int, [_P, _P, _P, _P, _I32, _F, _P, _I32, _P]),
    "mafe_frontend_run_host": (C.c_int, [_P, _P, _P, _I32, _P, _I32, _F, _P, _P, _I32, _I32]),
    "mafe_magphase": (C.c_int, [_P, _P, _I64, _F, _P, _P]),
    "mafe_amplitude_to_db": (C.c_int, [_P, _P, _P, _I64, _I64, _F, _F, _F, _F]),
    "mafe_db_to_amplitude": (C.c_int, [_P, _P, _P, _I64, _F, _F]),
    "mafe_melscale": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _P, _I32]),
    "mafe_transpose": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I64]),
    "mafe_istft": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P]),
    "mafe_cmvn_utt": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32]),
    "mafe_cmvn_scalar": (C.c_int, [_P, _P, _P, _I32, _I32, _I32]),
    "mafe_cmvn_stats_accumulate": (C.c_int, [_P, _P, _I64, _I32, _P]),
    "mafe_cmvn_apply": (C.c_int, [_P, _P, _I64, _I32, _P, _P]),
    "mafe_compute_deltas": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I64, _I64, _I32, _I32]),
    "mafe_context_window": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32]),
    "mafe_pad_sequence": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, C.c_float, _I32, _P, _P]),
    "mafe_sliding_window_cmn": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32]),
    "mafe_mask_rects": (C.c_int, [_P, _P, _P, _I32, _I32, _P, _I32, C.c_float]),
    "mafe_phase_vocoder": (C.c_int, [_P, _P, _I32, _I32, _I32, C.c_double, _P, _I32, _P]),
    "mafe_median_filter": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32]),
    "mafe_hpss_masks": (C.c_int, [_P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, _I32, _P, _P]),
    "mafe_wav_parse": (C.c_int, [_P, C.c_int64, C.c_double, C.c_double, _I32, _P]),
    "mafe_wav_stage": (C.c_int, [_P, _P, _I32, _I32, _P, _P, _P, C.c_int64, C.POINTER(_I32)]),
    "mafe_wav_files_open": (C.c_int, [_P, _I32, _I32, C.POINTER(_P), _P, _P, C.POINTER(_I32)]),
    "mafe_wav_files_pack": (C.c_int, [_P, _P, C.c_int64]),
    "mafe_wav_files_close": (C.c_int, [_P]),
    "mafe_wav_decode": (C.c_int, [_P, _P, C.c_int64, _I32, _I32, _I32, C.c_double, _P]),
    "mafe_resample_workspace": (C.c_int, [_I32, C.c_int64, C.c_int64, C.POINTER(C.c_size_t)]),
    "mafe_resample_fft": (C.c_int, [_P, _P, _I32, C.c_int64, C.c_int64, _P, _P, C.c_size_t]),
}


class WavInfo(C.Structure):
    """``mafe_wav_info`` (include/mafe.h)."""
    _fields_ = [(n, C.c_int32) for n in ("format_tag", "channels", "sample_rate", "bytes_per_second", "block_align",
                                         "bit_depth", "big_endian", "sample_kind", "bytes_per_sample", "warnings",
                                         "error_kind", "reserved")] + \
               [(n, C.c_int64) for n in ("data_offset", "n_items", "data_chunk_bytes")]


WAV_U8, WAV_I8, WAV_I16, WAV_I24, WAV_I32, WAV_I40, WAV_I48, WAV_I56, WAV_I64, WAV_F32, WAV_F64 = range(1, 12)
WAV_OUT_F32, WAV_OUT_F64, WAV_OUT_I16 = 0, 1, 2
WAV_WARN_UNKNOWN_CHUNK, WAV_WARN_EOF, WAV_WARN_INCOMPLETE_ID = 1, 2, 4
WAV_ERR_VALUE, WAV_ERR_TYPE, WAV_ERR_UNBOUND, WAV_ERR_ZERODIV, WAV_ERR_STRUCT, WAV_ERR_OS = 1, 2, 3, 4, 5, 6

_lib = None


def header_symbols():
    """Every function name declared in include/mafe.h (used by the CPU-side export test)."""
    with open(HEADER_PATH) as fh:
        src = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(mafe_[a-z0-9_]+)\s*\(", src)))


def load():
    """dlopen libmafe.so (no CUDA initialisation happens here) and attach prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise MafeError(
            "libmafe.so not found at %s -- build it with `python -m mindaudio_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback on this path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.mafe_version() != 102:
        raise MafeError("libmafe.so version mismatch: %d" % lib.mafe_version())
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        msg = load().mafe_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        raise MafeError("libmafe error %d: %s" % (rc, msg))
