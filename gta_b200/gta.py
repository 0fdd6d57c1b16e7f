"""Host-side mirror of the reference's GTA operator interface (source/utils/gta.py), backed by the CUDA library.

Same names, argument meaning and return convention as the reference so that it drops in behind
`source.layers.Attention.forward` (source/layers.py:419-428) without touching encoder/decoder code:

    import gta_b200.gta as fast
    fast.install()        # rebinds source.layers.multihead_geometric_transform_attention (+ source.utils.gta)

Forward and backward: under autograd the op is a torch.autograd.Function whose backward is the library's fused
gta_attn_bwd (gradients for q, k, v and trans_coeff).  Every f_dims layout and flag of the reference function is served by the CUDA
library (the `t2` block, `euclid_sim` and layouts with blocks that are not multiples of 8 through its generic path).
Calls that need autograd through the generic path or a CPU tensor are delegated to the original reference function when it was captured by
install(), and raise NotImplementedError otherwise — there is no silent CPU or PyTorch fallback inside this package.
"""
from __future__ import annotations

import warnings
from typing import Optional

import numpy as np
import torch

from . import ops
from .ops import PackedReps
from .synth import make_2dcoord  # noqa: F401  (same values as source/utils/gta.py:9-16)

_original = None          # the reference implementation captured by install()
# The reference returns the full [B,H,Tq,Tk] attention map as its second output; it is only read when a Transformer is
# called with return_last_attmap=True (source/layers.py:478-480, heads == 1, off in every shipped config — SURVEY T7).
# Set this to True to have the drop-in materialise it (gta_attn_probs); the default returns None in its place.
RETURN_ATTENTION_MAP = False
_PACK_KEY = "_gta_b200_packed"


def scale_mask(trans_coeff, device):
    """[[1,1,1,tc]x3,[0,0,0,1]] (source/utils/gta.py:40-44); the kernels apply it on load."""
    msk = torch.ones(4, 4, device=device)
    msk[:3, 3] = msk[:3, 3] * torch.as_tensor(trans_coeff, device=device, dtype=torch.float32).reshape(())
    msk[3, :3] = 0
    return msk


def make_SO2mats(coord, nfreqs, max_freqs=(1, 1), shared_freqs=False):
    """coord [..., 2] -> [..., nfreqs, 2, 2, 2] exactly as source/utils/gta.py:47-69 returns it
    (callers flatten dims -4,-3 into the pair index freq*2+axis, encoder.py:195)."""
    if coord.shape[-1] != 2:
        raise NotImplementedError("gta_b200.make_SO2mats: only 2-d coordinates are implemented")
    m = ops.so2_mats(coord, nfreqs, max_freqs, shared_freqs)          # [..., 2*nfreqs, 2, 2]
    return m.reshape(*coord.shape[:-1], nfreqs, 2, 2, 2)


def make_T2mats(coord):
    """coord [..., 2] -> [..., 3, 3] exactly as source/utils/gta.py:72-89 returns it."""
    return ops.t2_mats(coord)


def _pack_reps(reps: dict, f_dims: dict, B: int, euclid: bool = False) -> PackedReps:
    """Reference-format rep tensors (the `extras` dict) -> packed fp32 tables.  Cached in the dict, keyed on the
    identity of the source tensors (the decoder overwrites the *_q entries, decoder.py:259-346)."""
    g = lambda n: int(f_dims.get(n, 0) or 0)
    src = tuple(id(reps.get(n)) for n in ("inv_se3rep_q", "se3rep_k", "so3rep_q", "so3rep_k", "so2rep_q", "so2rep_k",
                                          "se3rep_q", "t2rep_q", "t2rep_k")) + (bool(euclid),)
    hit = reps.get(_PACK_KEY)
    if hit is not None and hit[0] == src:
        return hit[1]
    f = lambda t: t.detach().to(torch.float32)
    p = PackedReps()
    if g("se3"):
        p.se3_q = f(reps["inv_se3rep_q"]).reshape(B, -1, 16).contiguous()
        p.se3_k = f(reps["se3rep_k"]).reshape(B, -1, 16).contiguous()
        p.n_q_views, p.n_k_views = p.se3_q.shape[1], p.se3_k.shape[1]
        if euclid:
            p.se3_qi = f(reps["se3rep_q"]).reshape(B, -1, 16).contiguous()
    if g("t2"):
        # make_T2mats puts the token's coordinates in the last row of an otherwise constant matrix (gta.py:85-89)
        xy = lambda m: f(m)[..., 2, :2].reshape(B, -1, 2).contiguous()
        p.t2_q = xy(reps["t2rep_q"])
        p.t2_k = p.t2_q if reps["t2rep_k"] is reps["t2rep_q"] else xy(reps["t2rep_k"])
    if g("so3"):
        dq, dk = reps["so3rep_q"], reps["so3rep_k"]
        if len(dq) != 2 or dq[0].shape[-1] != 3 or dq[1].shape[-1] != 5:
            raise NotImplementedError("gta_b200: so3 reps must be [D_1, D_2]")
        p.so3_q = torch.cat([f(dq[0]).reshape(B, -1, 9), f(dq[1]).reshape(B, -1, 25)], -1).contiguous()
        p.so3_k = torch.cat([f(dk[0]).reshape(B, -1, 9), f(dk[1]).reshape(B, -1, 25)], -1).contiguous()
        p.n_q_views, p.n_k_views = p.so3_q.shape[1], p.so3_k.shape[1]
    if g("so2"):
        def cs(m):  # [B,T,C,2,2] = [[c,-s],[s,c]] -> [B,T,C,2]
            m = f(m)
            return torch.stack([m[..., 0, 0], m[..., 1, 0]], -1).contiguous()
        p.so2_q = cs(reps["so2rep_q"])
        p.so2_k = p.so2_q if reps["so2rep_k"] is reps["so2rep_q"] else cs(reps["so2rep_k"])
    reps[_PACK_KEY] = (src, p)
    return p


def _delegate(reason, *args, **kwargs):
    if _original is None:
        raise NotImplementedError("gta_b200: %s is not implemented by the fused path" % reason)
    warnings.warn("gta_b200: %s -> delegating to the reference implementation" % reason, stacklevel=3)
    return _original(*args, **kwargs)


class _FusedGtaAttention(torch.autograd.Function):
    """Differentiable wrapper: forward = gta_attn_fwd (+ log-sum-exp), backward = gta_attn_bwd.  Gradients flow to q, k, v
    and to the layer's trans_coeff parameter (source/layers.py:188-191); the rep tensors are constants, as in the
    reference (SO(3) reps are detached at gta.py:194-197, the SE(3) / SO(2) / coordinates come from the batch)."""

    @staticmethod
    def forward(ctx, q, k, v, trans_coeff, packed, f_dims, scale, v_transform):
        tc = None if trans_coeff is None else trans_coeff.detach()
        out, lse = ops.gta_attention_fwd(q, k, v, packed, f_dims, trans_coeff=tc, scale=scale, v_transform=v_transform,
                                         return_lse=True)
        ctx.save_for_backward(q, k, v, out, lse, tc if tc is not None else q.new_empty(0))
        ctx.packed, ctx.f_dims, ctx.scale, ctx.v_transform = packed, f_dims, scale, v_transform
        ctx.has_tc = tc is not None
        ctx.tc_shape = None if trans_coeff is None else trans_coeff.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse, tc = ctx.saved_tensors
        dq, dk, dv, dtc = ops.gta_attention_bwd(dout, q, k, v, out, lse, ctx.packed, ctx.f_dims,
                                                trans_coeff=tc if ctx.has_tc else None, scale=ctx.scale,
                                                v_transform=ctx.v_transform)
        gtc = None
        if ctx.has_tc and ctx.needs_input_grad[3] and dtc is not None:
            gtc = dtc.reshape(ctx.tc_shape).to(tc.dtype)
        return dq, dk, dv, gtc, None, None, None, None


def multihead_geometric_transform_attention(q, k, v, attn_fn, f_dims, reps, trans_coeff=1.0, v_transform=True,
                                            euclid=False, **kwargs):
    """Drop-in for source/utils/gta.py:92-279.  q [B,H,Tq,C], k,v [B,H,Tk,C] (strided views are consumed
    as they are); returns (out [B,H,Tq,C], None) — the attention map is never materialised."""
    args = (q, k, v, attn_fn, f_dims, reps)
    kw = dict(trans_coeff=trans_coeff, v_transform=v_transform, euclid=euclid, **kwargs)
    g = lambda n: int(f_dims.get(n, 0) or 0)
    if not q.is_cuda:
        return _delegate("a non-CUDA tensor", *args, **kw)
    if q.dtype not in (torch.bfloat16, torch.float32) or k.dtype != q.dtype or v.dtype != q.dtype:
        return _delegate("dtype %s" % q.dtype, *args, **kw)
    needs_grad = torch.is_grad_enabled() and (any(t.requires_grad for t in (q, k, v)) or
                                              (torch.is_tensor(trans_coeff) and trans_coeff.requires_grad))
    if needs_grad and (euclid or g("t2") or any(g(n) % 8 for n in ("triv", "se3", "so3", "so2"))):
        return _delegate("autograd through the t2 / euclid_sim / unaligned-block path", *args, **kw)
    B, H, Tq, D = q.shape
    packed = _pack_reps(reps, f_dims, B, euclid)
    if g("so3") and not g("se3"):
        return _delegate("so3 without se3 (undefined in the reference as well, SURVEY T5)", *args, **kw)
    if not g("se3") and not g("so3"):
        packed.n_q_views = packed.n_k_views = 1
    if Tq % packed.n_q_views or k.shape[2] % packed.n_k_views:
        raise ValueError("token count must be divisible by the number of views")
    tau = kwargs.get("tau", 1.0)
    scale = float(getattr(attn_fn, "scale", D ** -0.5)) / float(tau)
    tc = None
    if g("se3"):
        tc = trans_coeff if torch.is_tensor(trans_coeff) else torch.tensor([float(trans_coeff)], device=q.device)
    if needs_grad:
        tc_param = tc if (tc is not None and torch.is_tensor(trans_coeff)) else tc
        return _FusedGtaAttention.apply(q, k, v, tc_param, packed, dict(f_dims), scale, bool(v_transform)), None
    if RETURN_ATTENTION_MAP and not euclid:
        out, lse = ops.gta_attention_fwd(q, k, v, packed, f_dims, trans_coeff=tc, scale=scale, v_transform=v_transform,
                                         return_lse=True)
        return out, ops.gta_attention_probs(q, k, lse, packed, f_dims, trans_coeff=tc, scale=scale)
    out = ops.gta_attention_fwd(q, k, v, packed, f_dims, trans_coeff=tc, scale=scale, v_transform=v_transform,
                                euclid=euclid)
    return out, None


def install(layers_module=None, gta_module=None):
    """Rebind the reference's module globals to the fused op (source/layers.py:6,419 resolves the name at call
    time).  Returns the original function."""
    global _original
    import importlib
    layers_module = layers_module or importlib.import_module("source.layers")
    gta_module = gta_module or importlib.import_module("source.utils.gta")
    if _original is None:
        _original = gta_module.multihead_geometric_transform_attention
    layers_module.multihead_geometric_transform_attention = multihead_geometric_transform_attention
    gta_module.multihead_geometric_transform_attention = multihead_geometric_transform_attention
    return _original


def uninstall(layers_module=None, gta_module=None):
    global _original
    import importlib
    if _original is None:
        return
    layers_module = layers_module or importlib.import_module("source.layers")
    gta_module = gta_module or importlib.import_module("source.utils.gta")
    layers_module.multihead_geometric_transform_attention = _original
    gta_module.multihead_geometric_transform_attention = _original
    _original = None
