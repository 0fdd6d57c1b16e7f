"""TEST INFRASTRUCTURE — not product code.  Drives the *unmodified* reference (autonomousvision/gta,
mounted read-only at /root/reference in the build container; its model files installed unmodified into
baseline/_ref for the GPU box, see baseline/install_ref.py) to validate the oracle restatements, to generate
the golden vectors under tests/golden/, and to time the reference's own CPU path (bench.py --impl reference
and the cpu_baseline leg).  Never imported by the product package.

What it calls in the reference (no reference source is copied here):
  * ImprovedSRTEncoder.pre_compute_reps   source/encoder.py:183-265   (self-attention reps)
  * ImprovedSRTDecoder.pre_compute_reps   source/decoder.py:247-353   (cross-attention query reps)
  * multihead_geometric_transform_attention  source/utils/gta.py:92-279
  * AttnFn == softmax(q k^T * scale / tau) v   source/layers.py:202-211 (re-stated inline below
    because the class is local to Attention.__init__)

Import quirks handled: J_dense.pt is loaded from a CWD-relative path at import
time, and source.encoder imports a symbol `ray2rotation` that does not exist at this commit.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))
from baseline import ref_loader  # noqa: E402  ($GTA_REF, /root/reference or baseline/_ref; handles the import quirks)


def available() -> bool:
    return ref_loader.available()


def load():
    """The reference modules: namespace(gta, wigner_d, layers, encoder, decoder, models_nvs, root)."""
    return ref_loader.load()


class _AttnFn:
    """softmax(q k^T * scale / tau) v — what AttnFn.forward computes (source/layers.py:207-211)."""

    def __init__(self, scale, tau=1.0, euclid=False):
        self.scale, self.tau, self.euclid = scale, tau, euclid

    def __call__(self, q, k, v):
        sim = q @ k.transpose(-1, -2)
        if self.euclid:   # EuclidAttnFn.forward (source/layers.py:219-223), also local to Attention.__init__
            sim = sim - 0.5 * q.pow(2).sum(-1)[..., None] - 0.5 * k.pow(2).sum(-1)[..., None, :]
        attn = torch.softmax(sim * self.scale / self.tau, dim=-1)
        return attn @ v, attn


def attn_args(cfg):
    a = dict(f_dims=dict(cfg.f_dims), so2=cfg.so2, so3=cfg.so3, max_freq_h=cfg.max_freq_h,
             max_freq_w=cfg.max_freq_w, shared_freqs=cfg.shared_freqs)
    if not a["f_dims"].get("so2"):
        # The reference's pre_compute_reps only defines `NqTq` inside its so2 branch and then uses it in the se3
        # branch (encoder.py:196,238-242): without an so2 block (runs/*/GTA/gta_t2) it raises UnboundLocalError.
        # For REP CONSTRUCTION ONLY, ask for a one-frequency so2 table as well; the attention call still gets the
        # real f_dims, so the extra table is never read.
        a["f_dims"]["so2"] = 4
        a["so2"] = 1
    return a


def ref_reps(cfg, extr_q, extr_k, coord_q, coord_k, cross: bool):
    """Run the reference's own pre_compute_reps and return its `extras` dict."""
    m = load()
    extras = {}
    B = extr_k.shape[0]
    Nk, Nq = extr_k.shape[1], extr_q.shape[1]
    extras["input_transforms"] = extr_k
    extras["input_coord"] = coord_k.reshape(B, Nk, -1, 2)
    enc = m.encoder.ImprovedSRTEncoder.__new__(m.encoder.ImprovedSRTEncoder)
    m.encoder.ImprovedSRTEncoder.pre_compute_reps(enc, attn_args(cfg), extras)
    if cross:
        extras["target_transforms"] = extr_q
        extras["target_coord"] = coord_q.reshape(B, Nq, -1, 2)
        dec = m.decoder.ImprovedSRTDecoder.__new__(m.decoder.ImprovedSRTDecoder)
        m.decoder.ImprovedSRTDecoder.pre_compute_reps(dec, attn_args(cfg), extras)
    return extras


def ref_gta_attention(cfg, inp, trans_coeff=0.01, tau=1.0, dtype=torch.float32):
    """Reference forward on the given inputs (dict from gta_b200.synth.make_inputs)."""
    m = load()
    cross = inp["extr_q"] is not inp["extr_k"]
    c = lambda t: t.to(dtype)
    extras = ref_reps(cfg, c(inp["extr_q"]), c(inp["extr_k"]), c(inp["coord_q"]), c(inp["coord_k"]), cross)
    fn = _AttnFn(cfg.head_dim ** -0.5, tau, euclid=getattr(cfg, "euclid", False))
    tc = torch.tensor([trans_coeff], dtype=dtype)
    out, attn = m.gta.multihead_geometric_transform_attention(
        c(inp["q"]), c(inp["k"]), c(inp["v"]), attn_fn=fn, f_dims=dict(cfg.f_dims), reps=extras,
        trans_coeff=tc, v_transform=cfg.v_transform, euclid=getattr(cfg, "euclid", False))
    return out, extras
