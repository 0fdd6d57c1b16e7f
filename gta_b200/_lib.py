"""ctypes binding of libgta_b200.so (C ABI declared in include/gta_b200.h).

The library is built in-tree by `python -m gta_b200.build` (nvcc, sm_100a).  Loading fails loudly when the
shared object is missing: there is no CPU / PyTorch fallback on the product path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# GTA_B200_LIB selects a tuning build of the same library (python -m gta_b200.build -D... --out ...)
LIB_PATH = os.environ.get("GTA_B200_LIB") or os.path.join(_HERE, "libgta_b200.so")

GTA_DTYPE_BF16, GTA_DTYPE_F32 = 0, 1
GTA_FLAG_SKIP_STAGE = 2
GTA_FLAG_STAGE_ONLY = 4
GTA_FLAG_FAST_FP32 = 128
GTA_FLAG_V1_PIPELINE = 16
GTA_FLAG_SINGLE_LAUNCH = 32
GTA_FLAG_TWO_LAUNCH = 1024
GTA_FLAG_RUNTIME_LAYOUT = 2048
GTA_FLAG_BWD_SPLIT = 4096
GTA_PIPELINE_NAMES = {0: "two launches", 1: "single launch", 2: "split precision", 3: "generic"}
GTA_FLAG_V3_PRESTAGED = 64
GTA_FLAG_V4_PIPELINE = 256
GTA_FLAG_V5_PIPELINE = 512


class GtaReps(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("se3_q", "se3_k", "so3_q", "so3_k", "so2_q", "so2_k", "se3_qi", "t2_q", "t2_k")]


class GtaAttnParams(ctypes.Structure):
    _fields_ = (
        [("q", c_void_p), ("k", c_void_p), ("v", c_void_p)]
        + [(n, c_int64) for n in ("q_stride_b", "q_stride_h", "q_stride_t", "k_stride_b", "k_stride_h",
                                  "k_stride_t", "v_stride_b", "v_stride_h", "v_stride_t")]
        + [("out", c_void_p), ("lse", c_void_p)]
        + [(n, c_int) for n in ("B", "H", "Tq", "Tk", "D", "Nq", "Nk", "triv", "se3", "so3", "so2")]
        + [("reps", GtaReps), ("trans_coeff", c_void_p), ("scale", c_float)]
        + [(n, c_int) for n in ("in_dtype", "out_dtype", "v_transform")]
        + [("workspace", c_void_p), ("workspace_bytes", c_size_t), ("flags", c_int), ("debug_clocks", c_void_p)]
        + [("t2", c_int), ("euclid", c_int)]
    )


class GtaAttnBwdParams(ctypes.Structure):
    _fields_ = [("fwd", GtaAttnParams), ("dout", c_void_p), ("dq", c_void_p), ("dk", c_void_p), ("dv", c_void_p),
                ("dtrans_coeff", c_void_p), ("workspace", c_void_p), ("workspace_bytes", c_size_t)]


# every symbol include/gta_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "gta_attn_fwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "gta_attn_fwd_workspace_bytes_ex": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "gta_attn_fwd_workspace_bytes_p": (c_size_t, [POINTER(GtaAttnParams)]),
    "gta_attn_fwd": (c_int, [POINTER(GtaAttnParams), c_void_p]),
    "gta_attn_fwd_pipeline": (c_int, [POINTER(GtaAttnParams)]),
    "gta_attn_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "gta_attn_bwd_workspace_bytes_p": (c_size_t, [POINTER(GtaAttnParams)]),
    "gta_attn_bwd": (c_int, [POINTER(GtaAttnBwdParams), c_void_p]),
    "gta_attn_probs_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "gta_attn_probs": (c_int, [POINTER(GtaAttnParams), c_void_p, c_void_p]),
    "gta_rotate_debug": (c_int, [POINTER(GtaAttnParams), c_void_p, c_void_p, c_void_p, c_void_p]),
    "gta_build_reps": (c_int, [c_void_p] * 4 + [c_int] * 6 + [c_float, c_float, c_int, c_int] + [c_void_p] * 7),
    "gta_so2_mats": (c_int, [c_void_p, c_int64, c_int, c_float, c_float, c_int, c_void_p, c_void_p]),
    "gta_wigner_d": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "gta_se3_inverse": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "gta_t2_mats": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "gta_last_error": (c_char_p, []),
    "gta_abi_version": (c_int, []),
}

# development hooks (include/gta_b200_dev.h -> libgta_b200_dev.so): probes, micro-benchmarks, the first-generation kernel
DEV_LIB_PATH = os.path.join(_HERE, "libgta_b200_dev.so")
DEV_SYMBOLS = {
    "gta_dev_umma_probe": (c_int, [c_void_p] * 4 + [c_int, c_int] + [c_void_p] * 3),
    "gta_dev_umma_bench": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gta_dev_softmax_bench": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gta_dev_attn_fwd_v0": (c_int, [POINTER(GtaAttnParams), c_void_p]),
    "gta_dev_last_error": (c_char_p, []),
}

_lib = None
_dev = None


class GtaError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GtaError(
                "gta_b200: %s is missing — build it with `python -m gta_b200.build` (needs nvcc). "
                "There is no CPU fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)            # AttributeError if the .so does not export it
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def dev_lib() -> ctypes.CDLL:
    """The development library (tests / tools only)."""
    global _dev
    if _dev is None:
        if not os.path.exists(DEV_LIB_PATH):
            raise GtaError("gta_b200: %s is missing — build it with `python -m gta_b200.build`" % DEV_LIB_PATH)
        l = ctypes.CDLL(DEV_LIB_PATH)
        for name, (res, args) in DEV_SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _dev = l
    return _dev


def check_dev(rc: int, what: str = "") -> None:
    if rc != 0:
        raise GtaError("%s failed (rc=%d): %s" % (what or "gta_b200 dev call", rc, dev_lib().gta_dev_last_error().decode()))


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise GtaError("%s failed (rc=%d): %s" % (what or "gta_b200 call", rc, lib().gta_last_error().decode()))
