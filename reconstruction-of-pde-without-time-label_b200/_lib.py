"""ctypes binding of libblindno_b200.so -- mirrors include/blindno_b200.h one to one.

There is no CPU fallback: if the library is missing it is built (nvcc), and if that fails the
import of any op raises.  Every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

MAX_LAYERS = 8
PREC_FP32, PREC_TF32, PREC_TF32X3 = 0, 1, 2
ABI_VERSION = 2

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libblindno_b200.so")

_fp = C.c_void_p   # all device pointers travel as integers


class SpectralShape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("ndim", "images", "c_in", "c_out", "hp", "wp", "m1", "m2", "prec")]


class FnoShape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("ndim", "images", "c_in", "width", "c_out", "hidden", "n_layers", "h", "w",
                 "hp", "wp", "out_h", "out_w", "m1", "m2", "prec")]


class FnoParams(C.Structure):
    _fields_ = [("fc0_w", _fp), ("fc0_b", _fp),
                ("conv_w", _fp * MAX_LAYERS), ("conv_b", _fp * MAX_LAYERS),
                ("spec_w1", _fp * MAX_LAYERS), ("spec_w2", _fp * MAX_LAYERS),
                ("fc1_w", _fp), ("fc1_b", _fp), ("fc2_w", _fp), ("fc2_b", _fp)]


FnoGrads = FnoParams   # same layout, writable pointers


class LiftInput(C.Structure):
    _fields_ = [("x_cl", _fp), ("bags", _fp), ("idx", _fp), ("grid", _fp),
                ("n_bags", C.c_int32), ("bag_len", C.c_int32), ("n_keep", C.c_int32), ("grid_dim", C.c_int32)]


EXPORTS = {
    # name: (restype, argtypes)
    "bdn_abi_version": (C.c_int, []),
    "bdn_last_error": (C.c_char_p, []),
    "bdn_pad_amount": (C.c_int, [C.c_int]),
    "bdn_kernel_launches": (C.c_int64, []),
    "bdn_device_sm_count": (C.c_int, []),
    "bdn_prepare_plan": (C.c_int, [C.c_int32] * 5),
    "bdn_profile_begin": (C.c_int, []),
    "bdn_profile_end": (C.c_long, [C.c_char_p, C.c_size_t]),
    "bdn_spectral_workspace_bytes": (C.c_size_t, [C.POINTER(SpectralShape)]),
    "bdn_spectral_forward": (C.c_int, [C.POINTER(SpectralShape), _fp, _fp, _fp, _fp, _fp, _fp, C.c_size_t, _fp]),
    "bdn_spectral_backward": (C.c_int, [C.POINTER(SpectralShape), _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                                        C.c_size_t, _fp]),
    "bdn_stage_wfwd": (C.c_int, [C.c_int32] * 5 + [_fp, _fp, C.c_int32, C.c_int32, _fp]),
    "bdn_fno_act_floats": (C.c_size_t, [C.POINTER(FnoShape)]),
    "bdn_fno_spec_floats": (C.c_size_t, [C.POINTER(FnoShape)]),
    "bdn_fno_workspace_bytes": (C.c_size_t, [C.POINTER(FnoShape)]),
    "bdn_fno_forward": (C.c_int, [C.POINTER(FnoShape), C.POINTER(FnoParams), C.POINTER(LiftInput), _fp, _fp, _fp,
                                  _fp, C.c_size_t, _fp]),
    "bdn_fno_backward": (C.c_int, [C.POINTER(FnoShape), C.POINTER(FnoParams), C.POINTER(LiftInput), _fp, C.c_int32,
                                   C.c_int32, _fp, _fp, C.POINTER(FnoGrads), _fp, _fp, C.c_size_t, _fp]),
    "bdn_stage_lift_forward": (C.c_int, [C.POINTER(FnoShape), _fp, _fp, C.POINTER(LiftInput), _fp, _fp]),
    "bdn_stage_lift_backward": (C.c_int, [C.POINTER(FnoShape), _fp, _fp, C.POINTER(LiftInput), _fp, _fp, _fp, _fp, _fp]),
    "bdn_stage_layer_workspace_bytes": (C.c_size_t, [C.POINTER(FnoShape)]),
    "bdn_stage_layer_forward": (C.c_int, [C.POINTER(FnoShape), _fp, C.c_int32, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                                          C.c_size_t, _fp]),
    "bdn_stage_layer_backward": (C.c_int, [C.POINTER(FnoShape), _fp, _fp, C.c_int32, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                                           _fp, _fp, _fp, C.c_size_t, _fp]),
    "bdn_fno_layer_path": (C.c_int, [C.POINTER(FnoShape)]),
    "bdn_stage_project_forward": (C.c_int, [C.POINTER(FnoShape)] + [_fp] * 7),
    "bdn_stage_project_backward": (C.c_int, [C.POINTER(FnoShape)] + [_fp] * 6 + [C.c_int32, C.c_int32] + [_fp] * 6),
    "bdn_bag_pool_lift_forward": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_int32, _fp]),
    "bdn_bag_pool_lift_backward": (C.c_int, [_fp, _fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _fp]),
    "bdn_nio_tail_forward": (C.c_int, [_fp] * 8 + [C.c_int32] * 6 + [_fp]),
    "bdn_nio_tail_backward": (C.c_int, [_fp] * 8 + [C.c_int32] * 6 + [_fp]),
    "bdn_bag_attention_saved_floats": (C.c_size_t, [C.c_int32, C.c_int32]),
    "bdn_bag_attention_workspace_floats": (C.c_size_t, [C.c_int32, C.c_int32]),
    "bdn_bag_attention_mean_forward": (C.c_int, [_fp] * 5 + [C.c_int32] * 3 + [C.c_float, _fp]),
    "bdn_bag_attention_mean_backward": (C.c_int, [_fp] * 7 + [C.c_int32] * 3 + [_fp]),
    "bdn_mse_heads_forward": (C.c_int, [C.POINTER(_fp), C.c_int32, C.c_int32, C.c_int64, _fp, _fp, C.POINTER(_fp), _fp, _fp]),
    "bdn_mse_heads_backward": (C.c_int, [C.POINTER(_fp), C.c_int32, C.c_int32, C.c_int64, _fp, _fp, C.POINTER(_fp), _fp]),
    "bdn_adam_step": (C.c_int, [_fp, _fp, _fp, _fp, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_int32, C.c_float, _fp]),
}

_lock = threading.Lock()
_lib = None


class BlindnoError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (building first if needed) the shared library.  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        _build.ensure_current()       # rebuilds a missing or stale library; raises if it cannot
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)     # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        if handle.bdn_abi_version() != ABI_VERSION:
            raise BlindnoError("libblindno_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise BlindnoError(f"{what} failed (status {rc}): {lib().bdn_last_error().decode()}")


def pad_amount(n: int) -> int:
    """int(round(n / 4)) with Python's round-half-to-even (FNOModules.py:105, :222-223)."""
    return int(round(n * 0.25))
