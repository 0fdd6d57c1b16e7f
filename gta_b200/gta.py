"""Host-side mirror of the reference's GTA operator interface (source/utils/gta.py), backed by the CUDA library.

Same names, argument meaning and return convention as the reference so that it drops in behind
`source.layers.Attention.forward` (source/layers.py:419-428) without touching encoder/decoder code:

    import gta_b200.gta as fast
    fast.install()        # rebinds source.layers.multihead_geometric_transform_attention (+ source.utils.gta) and the
                          # two pre_compute_reps methods (source/encoder.py:183, source/decoder.py:247)

Forward and backward: under autograd the op is a torch.autograd.Function whose backward is the library's fused
gta_attn_bwd (gradients for q, k, v and trans_coeff).  Every f_dims layout and flag of the reference function is served
by the CUDA library (the `t2` block, `euclid_sim` and layouts with blocks that are not multiples of 8 through its generic
path).  Calls the library does not implement (autograd through the generic path, a learnable softmax temperature that
requires a gradient, CPU tensors, fp16, per-token SE(3) reps) are delegated to the original reference function when it
was captured by install(), and raise NotImplementedError otherwise — there is no silent CPU or PyTorch fallback inside
this package.
"""
from __future__ import annotations

import os
import sys
import warnings
from typing import Optional

import torch

from . import ops
from .ops import PackedReps
from .synth import make_2dcoord  # noqa: F401  (same values as source/utils/gta.py:9-16)

_original = None          # the reference implementation captured by install()
_original_reps = {}       # class -> the reference's pre_compute_reps captured by install()
# The reference returns the full [B,H,Tq,Tk] attention map as its second output; the caller reads it only when
# Attention.forward was called with return_attmap=True (source/layers.py:441-444; Transformer does that for its last
# layer under return_last_attmap, layers.py:478-480 — heads == 1, off in every shipped config).  The drop-in
# looks at its caller's `return_attmap` and materialises the map (gta_attn_probs) exactly then; this global forces it
# on for direct callers of the function.
RETURN_ATTENTION_MAP = False
# fp32 q/k/v under autograd (runs/clevrtr/* train with mixed_prec: False): "bf16" = tensor-core math in bf16 with fp32
# accumulation in BOTH directions (the library's backward has no split-precision variant; forward and backward then see
# the same rounded operands, exactly like the reference under its own bf16 autocast); "reference" = delegate such calls
# to the reference function.  Forward-only fp32 calls always run the split-precision kernel (1e-3 budget).
FP32_TRAINING = os.environ.get("GTA_B200_FP32_TRAIN", "bf16")
_PACK_KEY = "_gta_b200_packed"      # extras[...] = (cache key, PackedReps)
_LAZY_KEY = "_gta_b200_lazy_reps"   # extras[...] = [(original pre_compute_reps, module, attn_kwargs), ...] not yet run
_warned = set()


def _warn_once(msg):
    if msg not in _warned:
        _warned.add(msg)
        warnings.warn("gta_b200: " + msg, stacklevel=3)


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


# ----------------------------------------------------------------------------------------------- rep tables
class _Unsupported(Exception):
    """Rep tensors in a format the packed tables cannot express -> the call is delegated to the reference."""


_REP_KEYS = ("inv_se3rep_q", "se3rep_k", "so3rep_q", "so3rep_k", "so2rep_q", "so2rep_k", "se3rep_q", "t2rep_q", "t2rep_k")


def _tensor_key(t):
    if t is None:
        return None
    if isinstance(t, (list, tuple)):
        return tuple(_tensor_key(x) for x in t)
    return (t.data_ptr(), t._version, tuple(t.shape), t.dtype, t.device.index)


def _pack_reps(reps: dict, f_dims: dict, B: int, euclid: bool = False) -> PackedReps:
    """Reference-format rep tensors (the `extras` dict as the reference's pre_compute_reps leaves it) -> packed fp32
    tables.  Cached in the dict; the key covers storage, version counter and shape of every source tensor, the active
    blocks and the batch (the decoder overwrites the *_q entries, decoder.py:259-346)."""
    g = lambda n: int(f_dims.get(n, 0) or 0)
    blocks = tuple(bool(g(n)) for n in ("se3", "so3", "so2", "t2"))
    key = tuple(_tensor_key(reps.get(n)) for n in _REP_KEYS) + (blocks, bool(euclid), int(B))
    hit = reps.get(_PACK_KEY)
    if hit is not None and hit[0] == key:
        return hit[1]
    f = lambda t: t.detach().to(torch.float32)
    p = PackedReps()
    views = []
    if g("se3"):
        eq, ik = reps["inv_se3rep_q"], reps["se3rep_k"]
        if eq.dim() != 4 or ik.dim() != 4:          # ray_to_se3: per-token [B,N,T,4,4] matrices (encoder.py:220-231)
            raise _Unsupported("per-token SE(3) reps (ray_to_se3)")
        p.se3_q = f(eq).reshape(B, -1, 16).contiguous()
        p.se3_k = f(ik).reshape(B, -1, 16).contiguous()
        views.append((p.se3_q.shape[1], p.se3_k.shape[1]))
        if euclid:
            p.se3_qi = f(reps["se3rep_q"]).reshape(B, -1, 16).contiguous()
    if g("t2"):
        # make_T2mats puts the token's coordinates in the last row of an otherwise constant matrix (gta.py:85-89)
        xy = lambda m: f(m)[..., 2, :2].reshape(B, -1, 2).contiguous()
        p.t2_q = xy(reps["t2rep_q"])
        p.t2_k = p.t2_q if reps["t2rep_k"] is reps["t2rep_q"] else xy(reps["t2rep_k"])
    if g("so3"):
        dq, dk = reps["so3rep_q"], reps["so3rep_k"]
        if len(dq) != 2 or len(dk) != 2 or dq[0].shape[-1] != 3 or dq[1].shape[-1] != 5:
            raise _Unsupported("so3 reps other than [D_1, D_2]")
        p.so3_q = torch.cat([f(dq[0]).reshape(B, -1, 9), f(dq[1]).reshape(B, -1, 25)], -1).contiguous()
        p.so3_k = torch.cat([f(dk[0]).reshape(B, -1, 9), f(dk[1]).reshape(B, -1, 25)], -1).contiguous()
        views.append((p.so3_q.shape[1], p.so3_k.shape[1]))
    if any(v != views[0] for v in views):
        raise _Unsupported("se3 and so3 reps with different view counts")
    if views:
        p.n_q_views, p.n_k_views = views[0]
    if g("so2"):
        def cs(m):  # [B,T,C,2,2] = [[c,-s],[s,c]] -> [B,T,C,2]
            m = f(m)
            return torch.stack([m[..., 0, 0], m[..., 1, 0]], -1).contiguous()
        p.so2_q = cs(reps["so2rep_q"])
        p.so2_k = p.so2_q if reps["so2rep_k"] is reps["so2rep_q"] else cs(reps["so2rep_k"])
    reps[_PACK_KEY] = (key, p)
    return p


def _packed_from_extras(reps: dict, f_dims: dict, B: int, euclid: bool) -> PackedReps:
    """Tables built on the device by the installed pre_compute_reps (one gta_build_reps launch), else packed from the
    reference-format tensors."""
    hit = reps.get(_PACK_KEY)
    if hit is not None and hit[0] == "built":
        return hit[1]
    return _pack_reps(reps, f_dims, B, euclid)


def _fast_pre_compute_reps(side):
    """Replacement for ImprovedSRTEncoder.pre_compute_reps (source/encoder.py:183-265, side='enc') and
    ImprovedSRTDecoder.pre_compute_reps (source/decoder.py:247-353, side='dec'): the packed fp32 tables come from ONE
    gta_build_reps launch on the raw poses / coordinates in `extras` — instead of ~60 ATen launches incl. a batched LU —
    and stay fp32 under the reference's bf16 autocast.  The reference-format tensors are produced lazily, by
    the original method, only if a later call has to be delegated to the reference function."""
    def pre_compute_reps(self, attn_kwargs, extras):
        orig = _original_reps.get(side)
        f_dims = attn_kwargs["f_dims"]
        g = lambda n: int(f_dims.get(n, 0) or 0)
        unsupported = [k for k in ("ray_to_se3", "zeroout_so3", "id_so3", "elementwise_mul") if attn_kwargs.get(k)]
        ek = extras.get("input_transforms")
        if unsupported or ek is None or not ek.is_cuda or (g("so3") and int(attn_kwargs.get("so3", 0)) != 2):
            return orig(self, attn_kwargs, extras)
        eq = ek if side == "enc" else extras["target_transforms"]
        ck = extras["input_coord"] if (g("so2") or g("t2")) else ek.new_zeros(ek.shape[0], ek.shape[1], 2)
        cq = ck if side == "enc" else (extras["target_coord"] if (g("so2") or g("t2")) else
                                       eq.new_zeros(eq.shape[0], eq.shape[1], 2))
        with torch.autocast("cuda", enabled=False):
            packed = ops.build_reps(eq, ek, cq, ck, so2_nfreqs=int(attn_kwargs.get("so2", 0) or 0) if g("so2") else 0,
                                    so3_maxdeg=2 if g("so3") else 0,
                                    max_freq_h=float(attn_kwargs.get("max_freq_h", 1)),
                                    max_freq_w=float(attn_kwargs.get("max_freq_w", 1)),
                                    shared_freqs=bool(attn_kwargs.get("shared_freqs", False)),
                                    se3=bool(g("se3")), t2=bool(g("t2")),
                                    euclid=bool(attn_kwargs.get("euclid_sim", False)))
        extras[_PACK_KEY] = ("built", packed)
        extras.setdefault(_LAZY_KEY, []).append((orig, self, attn_kwargs))
    pre_compute_reps.__name__ = "pre_compute_reps"
    return pre_compute_reps


def _materialise_reference_reps(reps: dict):
    """Run the reference's own pre_compute_reps calls that the installed fast versions skipped (delegation needs the
    reference-format tensors and the einsum lambdas)."""
    for orig, mod, kw in reps.pop(_LAZY_KEY, []):
        orig(mod, kw, reps)


def _delegate(reason, q, k, v, attn_fn, f_dims, reps, **kw):
    if _original is None:
        raise NotImplementedError("gta_b200: %s is not implemented by the fused path" % reason)
    _warn_once("%s -> delegating to the reference implementation" % reason)
    if isinstance(reps, dict):
        _materialise_reference_reps(reps)
    return _original(q, k, v, attn_fn, f_dims, reps, **kw)


# ----------------------------------------------------------------------------------------------- softmax temperature
def _closure_tau(attn_fn):
    """`tau` is not an argument of the reference function: AttnFn.forward closes over it (source/layers.py:195-211; a
    float 1.0, or the nn.Parameter of TemperatureAdjsutableSoftmax under `softmax: adjustable`).  Returns
    (tau, found)."""
    fwd = getattr(type(attn_fn), "forward", None) or getattr(attn_fn, "forward", None)
    fwd = getattr(fwd, "__func__", fwd)
    code, cells = getattr(fwd, "__code__", None), getattr(fwd, "__closure__", None)
    if code is None or not cells or "tau" not in code.co_freevars:
        return 1.0, False
    try:
        return cells[code.co_freevars.index("tau")].cell_contents, True
    except ValueError:       # empty cell
        return 1.0, False


def _caller_wants_attmap() -> bool:
    """Attention.forward(x, z, return_attmap, extras) (source/layers.py:292,441-444) is the function's only caller; its
    `return_attmap` argument decides whether the second return value is read."""
    try:
        f = sys._getframe(2)
    except ValueError:
        return False
    return f.f_code.co_name == "forward" and bool(f.f_locals.get("return_attmap", False))


class _FusedGtaAttention(torch.autograd.Function):
    """Differentiable wrapper: forward = gta_attn_fwd (+ log-sum-exp), backward = gta_attn_bwd.  Gradients flow to q, k, v
    and to the layer's trans_coeff parameter (source/layers.py:188-191); the rep tensors are constants, as in the
    reference (SO(3) reps are detached at gta.py:194-197, the SE(3) / SO(2) / coordinates come from the batch)."""

    @staticmethod
    def forward(ctx, q, k, v, trans_coeff, packed, f_dims, scale, v_transform, flags, euclid=False, tau=None):
        # tau: the learnable softmax temperature (layers.py:195-200) when it needs a gradient; `scale` already holds 1 / tau
        ctx.tau = tau
        tc = None if trans_coeff is None else trans_coeff.detach()
        out, lse = ops.gta_attention_fwd(q, k, v, packed, f_dims, trans_coeff=tc, scale=scale, v_transform=v_transform,
                                         return_lse=True, flags=flags, euclid=euclid)
        ctx.save_for_backward(q, k, v, out, lse, tc if tc is not None else q.new_empty(0))
        ctx.packed, ctx.f_dims, ctx.scale, ctx.v_transform, ctx.euclid = packed, f_dims, scale, v_transform, euclid
        ctx.has_tc = tc is not None
        ctx.tc_shape = None if trans_coeff is None else trans_coeff.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse, tc = ctx.saved_tensors
        dq, dk, dv, dtc = ops.gta_attention_bwd(dout, q, k, v, out, lse, ctx.packed, ctx.f_dims,
                                                trans_coeff=tc if ctx.has_tc else None, scale=ctx.scale,
                                                v_transform=ctx.v_transform, euclid=ctx.euclid)
        gtc = None
        if ctx.has_tc and ctx.needs_input_grad[3] and dtc is not None:
            gtc = dtc.reshape(ctx.tc_shape).to(tc.dtype)
        gtau = None
        if ctx.tau is not None and ctx.needs_input_grad[10]:
            # logits L = scale0 (q'.k') / tau  =>  dL/dtau = -L / tau, and sum_j G_ij L_ij = q'_i . dq'_i = q_i . dq_i (the reps
            # are linear maps: q . J^T g = (J q) . g), so d(loss)/d(tau) = -(1 / tau) sum(q * dq): one reduction, no extra kernel
            gtau = (-(q.float() * dq.float()).sum() / ctx.tau.detach().float().reshape(-1)[0]).reshape(ctx.tau.shape).to(ctx.tau.dtype)
        return dq, dk, dv, gtc, None, None, None, None, None, None, gtau


def multihead_geometric_transform_attention(q, k, v, attn_fn, f_dims, reps, trans_coeff=1.0, v_transform=True,
                                            euclid=False, **kwargs):
    """Drop-in for source/utils/gta.py:92-279.  q [B,H,Tq,C], k,v [B,H,Tk,C] (strided views are consumed
    as they are); returns (out [B,H,Tq,C], attn) where attn is the [B,H,Tq,Tk] map only when the caller asked for it
    (return_attmap) and None otherwise — the fused kernel never materialises it."""
    kw = dict(trans_coeff=trans_coeff, v_transform=v_transform, euclid=euclid, **kwargs)
    deleg = lambda why: _delegate(why, q, k, v, attn_fn, f_dims, reps, **kw)
    g = lambda n: int(f_dims.get(n, 0) or 0)
    if not q.is_cuda:
        return deleg("a non-CUDA tensor")
    if q.dtype not in (torch.bfloat16, torch.float32) or k.dtype != q.dtype or v.dtype != q.dtype:
        return deleg("dtype %s" % q.dtype)
    grad_on = torch.is_grad_enabled()
    needs_grad = grad_on and (any(t.requires_grad for t in (q, k, v)) or
                              (torch.is_tensor(trans_coeff) and trans_coeff.requires_grad))
    if g("so3") and not g("se3"):
        return deleg("so3 without se3 (undefined in the reference as well, SURVEY T5)")
    # softmax temperature: a closure variable of attn_fn.forward, never a keyword of this function
    tau, _ = _closure_tau(attn_fn)
    if "tau" in kwargs:
        tau = kwargs["tau"]
    tau_param = None
    if torch.is_tensor(tau):
        if grad_on and tau.requires_grad:
            if euclid:      # the |q'|^2, |k'|^2 terms of EuclidAttnFn are divided by tau as well: not covered by the q . dq identity
                return deleg("a learnable softmax temperature that needs a gradient, with euclid_sim")
            tau_param = tau
        tau = float(tau.detach().reshape(-1)[0])           # host read of a scalar parameter
    needs_grad = needs_grad or tau_param is not None
    if needs_grad and q.dtype == torch.float32 and FP32_TRAINING == "reference":
        return deleg("fp32 training (GTA_B200_FP32_TRAIN=reference)")
    B, H, Tq, D = q.shape
    try:
        packed = _packed_from_extras(reps, f_dims, B, euclid)
    except _Unsupported as e:
        return deleg(str(e))
    if not g("se3") and not g("so3"):
        packed.n_q_views = packed.n_k_views = 1
    if Tq % packed.n_q_views or k.shape[2] % packed.n_k_views:
        raise ValueError("token count must be divisible by the number of views")
    scale = float(getattr(attn_fn, "scale", D ** -0.5)) / float(tau)
    tc = None
    if g("se3"):
        tc = trans_coeff if torch.is_tensor(trans_coeff) else torch.tensor([float(trans_coeff)], device=q.device)
    want_map = RETURN_ATTENTION_MAP or _caller_wants_attmap()
    if want_map and euclid:
        return deleg("the attention map of euclid_sim")
    if needs_grad:
        flags = 0
        if q.dtype == torch.float32:
            _warn_once("fp32 q/k/v under autograd: forward and backward multiply in bf16 with fp32 accumulation "
                       "(1e-2 budget, as under bf16 autocast); set GTA_B200_FP32_TRAIN=reference to train through the "
                       "reference function instead")
            flags = ops.FLAG_FAST_FP32
        out = _FusedGtaAttention.apply(q, k, v, tc, packed, dict(f_dims), scale, bool(v_transform), flags, bool(euclid), tau_param)
        attn = None
        if want_map:
            with torch.no_grad():
                _, lse = ops.gta_attention_fwd(q, k, v, packed, f_dims, trans_coeff=tc, scale=scale,
                                               v_transform=v_transform, return_lse=True)
                attn = ops.gta_attention_probs(q, k, lse, packed, f_dims, trans_coeff=tc, scale=scale)
        return out, attn
    if want_map:
        out, lse = ops.gta_attention_fwd(q, k, v, packed, f_dims, trans_coeff=tc, scale=scale, v_transform=v_transform,
                                         return_lse=True)
        return out, ops.gta_attention_probs(q, k, lse, packed, f_dims, trans_coeff=tc, scale=scale)
    out = ops.gta_attention_fwd(q, k, v, packed, f_dims, trans_coeff=tc, scale=scale, v_transform=v_transform,
                                euclid=euclid)
    return out, None


def install(layers_module=None, gta_module=None, reps: bool = True):
    """Rebind the reference's module globals to the fused op (source/layers.py:6,419 resolves the name at call time)
    and, with reps=True, the two pre_compute_reps methods to the device rep builder.  Returns the original function."""
    global _original
    import importlib
    layers_module = layers_module or importlib.import_module("source.layers")
    gta_module = gta_module or importlib.import_module("source.utils.gta")
    if _original is None:
        _original = gta_module.multihead_geometric_transform_attention
    layers_module.multihead_geometric_transform_attention = multihead_geometric_transform_attention
    gta_module.multihead_geometric_transform_attention = multihead_geometric_transform_attention
    if reps:
        for side, mod, cls in (("enc", "source.encoder", "ImprovedSRTEncoder"), ("dec", "source.decoder", "ImprovedSRTDecoder")):
            m = sys.modules.get(mod)
            if m is None:
                try:
                    m = importlib.import_module(mod)
                except Exception:      # source.encoder does not import as shipped; the caller may have stubbed it
                    continue
            c = getattr(m, cls, None)
            if c is not None and side not in _original_reps:
                _original_reps[side] = c.pre_compute_reps
                _original_reps[side + "_cls"] = c
                c.pre_compute_reps = _fast_pre_compute_reps(side)
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
    for side in ("enc", "dec"):
        if side in _original_reps:
            _original_reps.pop(side + "_cls").pre_compute_reps = _original_reps.pop(side)
    _original = None
