"""ctypes binding of libsoftspoken_b200.so (the C ABI in include/softspoken_b200.h).

There is no CPU fallback: if the shared library has not been built, importing
this module raises, loudly, with the command that builds it.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SOFTSPOKEN_B200_LIB") or os.path.join(_HERE, "libsoftspoken_b200.so")   # override: A/B tuning builds

SS_OK = 0
SS_E_ARG, SS_E_CUDA, SS_E_BLOB, SS_E_CAPACITY, SS_E_NODEVICE, SS_E_RANGE = -1, -2, -3, -4, -5, -6
MODE_FP32, MODE_BF16, MODE_F16, MODE_F16X3 = 0, 1, 2, 3
MODES = {"fp32": MODE_FP32, "bf16": MODE_BF16, "f16": MODE_F16, "f16x3": MODE_F16X3}
DEFAULT_MODE = "f16x3"   # fp32-grade logits on tensor cores: the mode detections are bit-exact in
ABI_VERSION = 2


class SoftspokenError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"softspoken_b200 error {code}: {message}")
        self.code = code


class Interval(C.Structure):
    _fields_ = [("begin", C.c_int64), ("end", C.c_int64)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: the CUDA library is not built and softspoken_b200 has no CPU fallback. "
        "Build it with `python -c 'import __graft_entry__ as g; g.build()'` (or `make -C softspoken_b200/csrc`).")

lib = C.CDLL(LIB_PATH)

_p = C.c_void_p
_i64 = C.c_int64
_int = C.c_int

_SIGNATURES = {
    "ss_abi_version": (_int, []),
    "ss_last_error": (C.c_char_p, []),
    "ss_get_constant": (_int, [C.c_char_p, C.POINTER(C.c_double)]),
    "ss_device_count": (_int, [C.POINTER(_int)]),
    "ss_launch_count": (_int, [C.POINTER(C.c_uint64)]),
    "ss_ctx_create": (_int, [_int, _p, C.c_size_t, _int, C.POINTER(_p)]),
    "ss_ctx_destroy": (_int, [_p]),
    "ss_ctx_device_bytes": (_int, [_p, C.POINTER(C.c_size_t)]),
    "ss_ctx_reserve": (_int, [_p, _i64, _int]),
    "ss_ctx_set_refine": (_int, [_p, C.c_double, _int]),
    "ss_ctx_refine_stats": (_int, [_p, C.POINTER(C.c_uint64), _int]),
    "ss_plan_windows": (_i64, [_i64]),
    "ss_timeline_bins": (_i64, [_i64]),
    "ss_pad": (_int, [_p, _p, _i64, _p, _p]),
    "ss_features": (_int, [_p, _p, _i64, _p, _int, _p, _p]),
    "ss_resample": (_int, [_p, _p, _i64, _p, _i64, _int, _int, _int, _p, _p]),
    "ss_resample_pcm16": (_int, [_p, _p, _i64, _p, _i64, _int, _int, _int, _p, _p]),
    "ss_spectrogram_frames": (_i64, [_i64]),
    "ss_spectrogram": (_int, [_p, _p, _i64, _p, _p, _p]),
    "ss_spectrogram_pcm16": (_int, [_p, _p, _i64, _p, _p, _p]),
    "ss_spectrogram_db": (_int, [_p, _p, _i64, _p, _p]),
    "ss_classify": (_int, [_p, _p, _int, _p, _p, _int, _p]),
    "ss_average": (_int, [_p, _p, _int, _i64, _p, _p, _p]),
    "ss_regions": (_int, [_p, _p, _p, _i64, C.c_double, _int, _p, _p, _int, _p]),
    "ss_silence": (_int, [_p, _p, _i64, _p, _int, _p]),
    "ss_detect_device": (_int, [_p, _p, _i64, _int, _p, _p, _int, _p, _p]),
    "ss_detect_host": (_int, [_p, _p, _i64, _int, _p, _int, C.POINTER(_int), _p]),
    "ss_detect_host_batch": (_int, [_p, _int, _p, _p, _int, _p, _int, _p]),
    "ss_detect_device_pcm16": (_int, [_p, _p, _i64, _int, _p, _p, _int, _p, _p]),
    "ss_detect_host_pcm16": (_int, [_p, _p, _i64, _int, _p, _int, C.POINTER(_int), _p]),
    "ss_detect_host_batch_pcm16": (_int, [_p, _int, _p, _p, _int, _p, _int, _p]),
    "ss_decode_pcm16": (_int, [_p, _p, _i64, _int, _p, _p]),
    "ss_encode_pcm16": (_int, [_p, _p, _i64, _p, _p]),
    "ss_silence_pcm16": (_int, [_p, _p, _i64, _p, _int, _int, _p]),
    "ss_silence_pcm16_host": (_int, [_p, _p, _i64, _p, _int, _int]),
    "ss_check_health": (_int, [_p, _p]),
    "ss_silence_host": (_int, [_p, _p, _i64, _p, _int]),
    "ss_debug_check_guards": (_int, [_p, C.POINTER(C.c_uint64), C.POINTER(_int)]),
    "ss_debug_tc_profile": (_int, [_p, _int, _p]),
    "ss_debug_activation": (_int, [_p, _int, _int, _p, C.POINTER(_int), C.POINTER(_int), C.POINTER(_int), _p]),
}

EXPORTS = tuple(_SIGNATURES)

for _name, (_res, _args) in _SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = the library does not export the ABI
    _fn.restype = _res
    _fn.argtypes = _args

if lib.ss_abi_version() != ABI_VERSION:
    raise ImportError(f"libsoftspoken_b200.so ABI {lib.ss_abi_version()} != binding ABI {ABI_VERSION}: rebuild")


def check(rc: int) -> None:
    if rc != SS_OK:
        raise SoftspokenError(rc, lib.ss_last_error().decode("utf-8", "replace"))


def get_constant(name: str) -> float:
    v = C.c_double()
    check(lib.ss_get_constant(name.encode(), C.byref(v)))
    return v.value


def launch_count() -> int:
    n = C.c_uint64()
    check(lib.ss_launch_count(C.byref(n)))
    return int(n.value)


def device_count() -> int:
    n = _int()
    check(lib.ss_device_count(C.byref(n)))
    return n.value
