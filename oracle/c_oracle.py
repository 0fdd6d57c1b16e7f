"""TEST INFRASTRUCTURE — ctypes wrapper over oracle/libgta_oracle.so (the C restatement in
oracle/gta_oracle.c).  Imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgta_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gta_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libgta_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _f(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def gta_attention(cfg, q, k, v, extr_q, extr_k, coord_q, coord_k, trans_coeff=0.01, tau=1.0,
                  return_rotated=False):
    """q,k,v: array-likes [B,H,T,D] (any strides; copied to contiguous fp32).  Returns out [B,H,Tq,D]."""
    q, pq = _f(q); k, pk = _f(k); v, pv = _f(v)
    eq, peq = _f(extr_q); ek, pek = _f(extr_k); cq, pcq = _f(coord_q); ck, pck = _f(coord_k)
    B, H, Tq, D = q.shape
    Tk = k.shape[2]
    triv, se3, so3, so2 = cfg.dims()
    out = np.empty_like(q)
    rot = [np.empty_like(q), np.empty_like(k), np.empty_like(v)] if return_rotated else [None] * 3
    pf = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if a is not None else None
    rc = lib().oracle_gta_attention_ex(
        pq, pk, pv, peq, pek, pcq, pck,
        B, H, Tq, Tk, D, eq.shape[1], ek.shape[1], triv, se3, so3, so2, int(cfg.t2_dim()), int(bool(cfg.euclid)),
        int(cfg.so2), ctypes.c_float(cfg.max_freq_h), ctypes.c_float(cfg.max_freq_w),
        int(cfg.shared_freqs), ctypes.c_float(trans_coeff), ctypes.c_float(cfg.head_dim ** -0.5 / tau),
        int(cfg.v_transform), pf(out), pf(rot[0]), pf(rot[1]), pf(rot[2]))
    if rc:
        raise RuntimeError("oracle_gta_attention failed rc=%d" % rc)
    return (out, *rot) if return_rotated else out


def build_reps(cfg, extr_q, extr_k, coord_q, coord_k):
    """Packed fp32 rep tables (layout of include/gta_b200.h)."""
    eq, peq = _f(extr_q); ek, pek = _f(extr_k); cq, pcq = _f(coord_q); ck, pck = _f(coord_k)
    B, Nq, Nk = eq.shape[0], eq.shape[1], ek.shape[1]
    Tq, Tk = cq.shape[1], ck.shape[1]
    C = 2 * int(cfg.so2)
    o = dict(se3_q=np.zeros((B, Nq, 16), np.float32), se3_k=np.zeros((B, Nk, 16), np.float32),
             so3_q=np.zeros((B, Nq, 34), np.float32), so3_k=np.zeros((B, Nk, 34), np.float32),
             so2_q=np.zeros((B, Tq, max(C, 1), 2), np.float32), so2_k=np.zeros((B, Tk, max(C, 1), 2), np.float32))
    pf = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    rc = lib().oracle_build_reps(peq, pek, pcq, pck, B, Nq, Nk, Tq, Tk, int(cfg.so2),
                                 ctypes.c_float(cfg.max_freq_h), ctypes.c_float(cfg.max_freq_w),
                                 int(cfg.shared_freqs), pf(o["se3_q"]), pf(o["se3_k"]), pf(o["so3_q"]),
                                 pf(o["so3_k"]), pf(o["so2_q"]), pf(o["so2_k"]))
    if rc:
        raise RuntimeError("oracle_build_reps failed rc=%d" % rc)
    return o


def num_threads() -> int:
    return int(lib().oracle_num_threads())
