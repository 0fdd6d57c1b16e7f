"""torch-facing wrappers over the C ABI: device memory, streams and strides come from torch, everything
else is the CUDA library.  All tensors must live on a CUDA device."""
from __future__ import annotations

import ctypes
import dataclasses
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import GtaAttnBwdParams, GtaAttnParams, GtaReps, check, lib


FLAG_FAST_FP32 = _lib.GTA_FLAG_FAST_FP32


def _stream(dev=None) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _launch(dev, name: str, *args) -> None:
    """Every library call runs with the tensors' device current and on that device's current stream (the C ABI
    launches on the calling thread's current device)."""
    with torch.cuda.device(dev):
        check(getattr(lib(), name)(*args, _stream(dev)), name)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


@dataclasses.dataclass
class PackedReps:
    """Packed fp32 rep tables on the device (layout documented at GtaReps in include/gta_b200.h)."""
    se3_q: Optional[torch.Tensor] = None   # [B,Nq,16]  E_q (unscaled)
    se3_k: Optional[torch.Tensor] = None   # [B,Nk,16]  inv(E_k)
    so3_q: Optional[torch.Tensor] = None   # [B,Nq,34]
    so3_k: Optional[torch.Tensor] = None
    so2_q: Optional[torch.Tensor] = None   # [B,Tq,C,2]
    so2_k: Optional[torch.Tensor] = None
    se3_qi: Optional[torch.Tensor] = None  # [B,Nq,16]  inv(E_q) (euclid_sim only)
    t2_q: Optional[torch.Tensor] = None    # [B,Tq,2]   patch coordinates of the t2 block
    t2_k: Optional[torch.Tensor] = None
    n_q_views: int = 1
    n_k_views: int = 1

    _TABLES = ("se3_q", "se3_k", "so3_q", "so3_k", "so2_q", "so2_k", "se3_qi", "t2_q", "t2_k")

    def batch_slice(self, lo: int, hi: int) -> "PackedReps":
        """The tables of batch elements [lo, hi) (views: every table has the batch as its leading dim)."""
        kw = {n: (None if getattr(self, n) is None else getattr(self, n)[lo:hi]) for n in self._TABLES}
        return PackedReps(n_q_views=self.n_q_views, n_k_views=self.n_k_views, **kw)

    def c_struct(self) -> GtaReps:
        return GtaReps(*[_ptr(getattr(self, n)) for n in ("se3_q", "se3_k", "so3_q", "so3_k", "so2_q", "so2_k",
                                                           "se3_qi", "t2_q", "t2_k")])


def build_reps(extr_q: torch.Tensor, extr_k: torch.Tensor, coord_q: torch.Tensor, coord_k: torch.Tensor, *,
               so2_nfreqs: int, so3_maxdeg: int, max_freq_h: float = 1.0, max_freq_w: float = 1.0,
               shared_freqs: bool = False, se3: bool = True, t2: bool = False, euclid: bool = False) -> PackedReps:
    """Device-side pre_compute_reps (reference: source/encoder.py:183-265, source/decoder.py:247-353).
    extr_* [B,N,4,4], coord_* [B,T,2] (or [B,N,t,2])."""
    dev = extr_k.device
    assert dev.type == "cuda", "gta_b200 runs on CUDA devices only"
    eq, ek = _f32c(extr_q), _f32c(extr_k)
    B, Nq, Nk = eq.shape[0], eq.shape[1], ek.shape[1]
    cq = _f32c(coord_q).reshape(B, -1, 2)
    ck = _f32c(coord_k).reshape(B, -1, 2)
    Tq, Tk = cq.shape[1], ck.shape[1]
    same = (extr_q is extr_k) and (coord_q is coord_k)
    f = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
    r = PackedReps(n_q_views=Nq, n_k_views=Nk)
    if se3 or so3_maxdeg:
        r.se3_q, r.se3_k = f(B, Nq, 16), f(B, Nk, 16)
    if so3_maxdeg:
        r.so3_q, r.so3_k = f(B, Nq, 34), f(B, Nk, 34)
    if so2_nfreqs:
        r.so2_k = f(B, Tk, 2 * so2_nfreqs, 2)
        r.so2_q = r.so2_k if same else f(B, Tq, 2 * so2_nfreqs, 2)
    _launch(dev, "gta_build_reps", _ptr(eq), _ptr(ek), _ptr(cq), _ptr(ck), B, Nq, Nk, Tq, Tk, int(so2_nfreqs),
                               float(max_freq_h), float(max_freq_w), int(shared_freqs), int(so3_maxdeg),
                               _ptr(r.se3_q), _ptr(r.se3_k), _ptr(r.so3_q), _ptr(r.so3_k), _ptr(r.so2_q),
                               _ptr(r.so2_k))
    if euclid and se3:          # the euclid branch multiplies the query points by inv(E_q) itself (gta.py:140,153)
        r.se3_qi = r.se3_k if same else f(B, Nq, 16)
        if not same:
            _launch(dev, "gta_se3_inverse", _ptr(eq), B * Nq, _ptr(r.se3_qi))
    if t2:                      # make_T2mats needs nothing but the coordinates (gta.py:72-89)
        r.t2_q, r.t2_k = cq, ck
    return r


_DT = {torch.bfloat16: _lib.GTA_DTYPE_BF16, torch.float32: _lib.GTA_DTYPE_F32}
_ws_cache: Dict[tuple, torch.Tensor] = {}


def _workspace(dev, nbytes: int) -> torch.Tensor:
    """Scratch for the staged K'/V' tile images, one buffer per (device, stream): two streams running the op
    concurrently never share staged tiles, and work queued on one stream reuses its buffer in stream order."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), _stream(dev))
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes + 1024, device=dev, dtype=torch.uint8)
        _ws_cache[key] = ws
    return ws


def _params(q, k, v, out, reps: PackedReps, f_dims, trans_coeff, scale, v_transform, flags, lse=None, euclid=False):
    assert q.is_cuda and k.is_cuda and v.is_cuda, "gta_b200 runs on CUDA devices only"
    assert q.dtype == k.dtype == v.dtype and q.dtype in _DT, "q/k/v must all be bf16 or all fp32"
    B, H, Tq, D = q.shape
    Tk = k.shape[2]
    for t in (q, k, v):
        assert t.stride(3) == 1, "head dim must be contiguous"
    g = lambda n: int(f_dims.get(n, 0) or 0)
    p = GtaAttnParams()
    p.q, p.k, p.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    p.q_stride_b, p.q_stride_h, p.q_stride_t = q.stride(0), q.stride(1), q.stride(2)
    p.k_stride_b, p.k_stride_h, p.k_stride_t = k.stride(0), k.stride(1), k.stride(2)
    p.v_stride_b, p.v_stride_h, p.v_stride_t = v.stride(0), v.stride(1), v.stride(2)
    p.out = _ptr(out)
    p.lse = _ptr(lse)
    p.B, p.H, p.Tq, p.Tk, p.D = B, H, Tq, Tk, D
    p.Nq, p.Nk = reps.n_q_views, reps.n_k_views
    p.triv, p.se3, p.so3, p.so2, p.t2 = g("triv"), g("se3"), g("so3"), g("so2"), g("t2")
    p.euclid = int(bool(euclid))
    p.reps = reps.c_struct()
    p.trans_coeff = _ptr(trans_coeff)
    p.scale = float(scale)
    p.in_dtype = _DT[q.dtype]
    p.out_dtype = _DT[out.dtype] if out is not None else _DT[q.dtype]
    p.v_transform = int(bool(v_transform))
    p.flags = int(flags)
    return p


def gta_attention_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, reps: PackedReps, f_dims: dict, *,
                      trans_coeff: Optional[torch.Tensor] = None, scale: Optional[float] = None,
                      v_transform: bool = True, out_dtype: Optional[torch.dtype] = None,
                      return_lse: bool = False, flags: int = 0, debug_clocks: Optional[torch.Tensor] = None,
                      euclid: bool = False, out: Optional[torch.Tensor] = None):
    """q [B,H,Tq,D], k,v [B,H,Tk,D] (strided views allowed) -> out [B,H,Tq,D] as a permuted view of a
    contiguous [B,Tq,H,D] buffer (so the reference's 'b h n d -> b n (h d)' is free).  `out`: optional preallocated
    contiguous [B,Tq,H,D] destination."""
    B, H, Tq, D = q.shape
    dev = q.device
    if scale is None:
        scale = D ** -0.5
    if trans_coeff is not None:
        trans_coeff = trans_coeff.detach().to(device=dev, dtype=torch.float32).reshape(-1)[:1].contiguous()
    if out is None:
        out = torch.empty(B, Tq, H, D, device=dev, dtype=out_dtype or q.dtype)
    else:
        assert out.shape == (B, Tq, H, D) and out.is_contiguous() and out.device == dev, "out must be contiguous [B,Tq,H,D]"
    lse = torch.empty(B, H, Tq, device=dev, dtype=torch.float32) if return_lse else None
    p = _params(q, k, v, out, reps, f_dims, trans_coeff, scale, v_transform, flags, lse, euclid)
    nbytes = lib().gta_attn_fwd_workspace_bytes_p(p)
    ws = _workspace(dev, nbytes)
    base = ws.data_ptr()
    p.workspace = (base + 1023) // 1024 * 1024
    p.workspace_bytes = nbytes
    p.debug_clocks = _ptr(debug_clocks)
    _launch(dev, "gta_attn_fwd", p)
    res = out.permute(0, 2, 1, 3)
    return (res, lse) if return_lse else res


def gta_attention_bwd(dout: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor,
                      lse: torch.Tensor, reps: PackedReps, f_dims: dict, *, trans_coeff: Optional[torch.Tensor] = None,
                      scale: Optional[float] = None, v_transform: bool = True,
                      debug_clocks: Optional[torch.Tensor] = None, flags: int = 0, euclid: bool = False):
    """Backward of gta_attention_fwd.  `out` is the forward result ([B,H,Tq,D] view of a [B,Tq,H,D] buffer), `lse` its
    log-sum-exp, `dout` the gradient w.r.t. `out` (any layout).  Returns (dq, dk, dv, dtrans_coeff) with dq/dk/dv shaped
    like q/k/v (views of contiguous [B,T,H,D] buffers) and dtrans_coeff a [1] fp32 tensor (None without an se3 block).
    Head dims <= 96 run ONE fused kernel (dK, dV, bulk-reduced dQ partial sums) + a dq finishing kernel; `flags`:
    GTA_FLAG_BWD_SPLIT keeps the dK/dV + dQ kernel pair, GTA_FLAG_SINGLE_LAUNCH forces the fused kernel for calls of less than
    one wave of key tiles, GTA_FLAG_RUNTIME_LAYOUT its run-time-layout epilogue.  `euclid` / a t2 block / unaligned blocks: the
    element-wise rep passes around the same kernels (generic path)."""
    B, H, Tq, D = q.shape
    Tk = k.shape[2]
    dev = q.device
    if scale is None:
        scale = D ** -0.5
    if trans_coeff is not None:
        trans_coeff = trans_coeff.detach().to(device=dev, dtype=torch.float32).reshape(-1)[:1].contiguous()
    o_c = out.permute(0, 2, 1, 3).contiguous()                    # [B,Tq,H,D] (no copy for the forward's own buffer)
    do_c = dout.to(o_c.dtype).permute(0, 2, 1, 3).contiguous()
    dq = torch.empty(B, Tq, H, D, device=dev, dtype=q.dtype)
    dk = torch.empty(B, Tk, H, D, device=dev, dtype=q.dtype)
    dv = torch.empty(B, Tk, H, D, device=dev, dtype=q.dtype)
    has_se3 = bool(int(f_dims.get("se3", 0) or 0))
    dtc = torch.zeros(1, device=dev, dtype=torch.float32) if has_se3 else None
    bp = GtaAttnBwdParams()
    bp.fwd = _params(q, k, v, o_c, reps, f_dims, trans_coeff, scale, v_transform, flags, lse.contiguous(), euclid)
    bp.fwd.debug_clocks = _ptr(debug_clocks)
    bp.dout, bp.dq, bp.dk, bp.dv, bp.dtrans_coeff = _ptr(do_c), _ptr(dq), _ptr(dk), _ptr(dv), _ptr(dtc)
    nbytes = lib().gta_attn_bwd_workspace_bytes_p(ctypes.byref(bp.fwd))     # (covers the generic-path layouts: t2, unaligned blocks)
    ws = _workspace(dev, nbytes)
    bp.workspace = (ws.data_ptr() + 1023) // 1024 * 1024
    bp.workspace_bytes = nbytes
    _launch(dev, "gta_attn_bwd", bp)
    return dq.permute(0, 2, 1, 3), dk.permute(0, 2, 1, 3), dv.permute(0, 2, 1, 3), dtc


def gta_attention_probs(q: torch.Tensor, k: torch.Tensor, lse: torch.Tensor, reps: PackedReps, f_dims: dict, *,
                        trans_coeff: Optional[torch.Tensor] = None, scale: Optional[float] = None) -> torch.Tensor:
    """The attention map the reference returns next to `out` (source/layers.py:207-211): fp32 [B,H,Tq,Tk] from q, k and
    the log-sum-exp of a forward call (`gta_attention_fwd(..., return_lse=True)`).  Visualisation path."""
    B, H, Tq, D = q.shape
    Tk = k.shape[2]
    dev = q.device
    if scale is None:
        scale = D ** -0.5
    if trans_coeff is not None:
        trans_coeff = trans_coeff.detach().to(device=dev, dtype=torch.float32).reshape(-1)[:1].contiguous()
    attn = torch.empty(B, H, Tq, Tk, device=dev, dtype=torch.float32)
    dummy = torch.empty(16, device=dev, dtype=q.dtype)
    p = _params(q, k, k, dummy, reps, f_dims, trans_coeff, scale, True, 0, lse.contiguous())
    nbytes = lib().gta_attn_probs_workspace_bytes(B, H, Tq, Tk, D)
    ws = _workspace(dev, nbytes)
    p.workspace = (ws.data_ptr() + 1023) // 1024 * 1024
    p.workspace_bytes = nbytes
    _launch(dev, "gta_attn_probs", p, _ptr(attn))
    return attn


def rotate_debug(q, k, v, reps: PackedReps, f_dims: dict, *, trans_coeff=None, v_transform=True, euclid=False):
    """fp32 rotated operands (q', k', v') as contiguous [B,H,T,D] tensors — for tests."""
    dev = q.device
    if trans_coeff is not None:
        trans_coeff = trans_coeff.detach().to(device=dev, dtype=torch.float32).reshape(-1)[:1].contiguous()
    B, H, Tq, D = q.shape
    Tk = k.shape[2]
    qt = torch.empty(B, H, Tq, D, device=dev, dtype=torch.float32)
    kt = torch.empty(B, H, Tk, D, device=dev, dtype=torch.float32)
    vt = torch.empty(B, H, Tk, D, device=dev, dtype=torch.float32)
    dummy = torch.empty(16, device=dev, dtype=q.dtype)
    p = _params(q, k, v, dummy, reps, f_dims, trans_coeff, 1.0, v_transform, 0, euclid=euclid)
    _launch(dev, "gta_rotate_debug", p, _ptr(qt), _ptr(kt), _ptr(vt))
    return qt, kt, vt


def so2_mats(coord: torch.Tensor, nfreqs: int, max_freqs=(1, 1), shared_freqs: bool = False) -> torch.Tensor:
    """[..., 2] -> [..., 2*nfreqs, 2, 2] (pair index = freq*2 + axis)."""
    assert coord.is_cuda and coord.shape[-1] == 2
    c = _f32c(coord).reshape(-1, 2)
    out = torch.empty(c.shape[0], 2 * nfreqs, 2, 2, device=c.device, dtype=torch.float32)
    _launch(c.device, "gta_so2_mats", _ptr(c), c.shape[0], int(nfreqs), float(max_freqs[0]), float(max_freqs[1]),
                             int(shared_freqs), _ptr(out))
    return out.reshape(*coord.shape[:-1], 2 * nfreqs, 2, 2)


def t2_mats(coord: torch.Tensor, with_inverse: bool = False):
    """[..., 2] -> [..., 3, 3] = [[1,0,0],[0,1,0],[x,y,1]] (make_T2mats, source/utils/gta.py:72-89), optionally with
    the inverses the callers obtain from torch.linalg.inv (source/encoder.py:212)."""
    assert coord.is_cuda and coord.shape[-1] == 2
    c = _f32c(coord).reshape(-1, 2)
    m = torch.empty(c.shape[0], 3, 3, device=c.device, dtype=torch.float32)
    mi = torch.empty_like(m) if with_inverse else None
    _launch(c.device, "gta_t2_mats", _ptr(c), c.shape[0], _ptr(m), _ptr(mi))
    m = m.reshape(*coord.shape[:-1], 3, 3)
    return (m, mi.reshape(*coord.shape[:-1], 3, 3)) if with_inverse else m


def se3_inverse(extr: torch.Tensor) -> torch.Tensor:
    """[..., 4, 4] -> inverse (torch.linalg.inv at source/encoder.py:219), fp64 Gauss-Jordan on the device."""
    assert extr.is_cuda and extr.shape[-2:] == (4, 4)
    e = _f32c(extr).reshape(-1, 16)
    out = torch.empty_like(e)
    _launch(e.device, "gta_se3_inverse", _ptr(e), e.shape[0], _ptr(out))
    return out.reshape(extr.shape)


def wigner_d(R: torch.Tensor):
    """R [n,3,3] -> (D1 [n,3,3], D2 [n,5,5])."""
    assert R.is_cuda and R.shape[-2:] == (3, 3)
    r = _f32c(R).reshape(-1, 3, 3)
    d1 = torch.empty(r.shape[0], 3, 3, device=r.device, dtype=torch.float32)
    d2 = torch.empty(r.shape[0], 5, 5, device=r.device, dtype=torch.float32)
    _launch(r.device, "gta_wigner_d", _ptr(r), r.shape[0], _ptr(d1), _ptr(d2))
    return d1, d2


def umma_probe(A, Bm, P, V, p_in_tmem: bool):
    """tcgen05 descriptor / tile-image self-test (development library)."""
    D = A.shape[1]
    outS = torch.empty(128, 128, device=A.device, dtype=torch.float32)
    outO = torch.empty(128, D, device=A.device, dtype=torch.float32)
    with torch.cuda.device(A.device):
        _lib.check_dev(_lib.dev_lib().gta_dev_umma_probe(_ptr(A), _ptr(Bm), _ptr(P), _ptr(V), D, int(p_in_tmem), _ptr(outS),
                                                         _ptr(outO), _stream(A.device)), "gta_dev_umma_probe")
    return outS, outO


def pipeline_of(q, k, v, reps: PackedReps, f_dims: dict, *, flags: int = 0, euclid: bool = False) -> str:
    """Name of the pipeline gta_attn_fwd selects for this call (gta_attn_fwd_pipeline)."""
    dummy = torch.empty(16, device=q.device, dtype=q.dtype)
    p = _params(q, k, v, dummy, reps, f_dims, None, 1.0, True, flags, euclid=euclid)
    return _lib.GTA_PIPELINE_NAMES.get(lib().gta_attn_fwd_pipeline(p), "?")
