"""TEST INFRASTRUCTURE — a from-scratch torch restatement of the reference's GTA-attention path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker / CPU baseline.  The product path (gta_b200/) never does.

Parity status: PINNED.  tests/test_oracle.py checks this port against tests/golden/*.npz, which
were produced by running the unmodified reference in the build container
(tests/golden/gen_golden.py, via oracle/ref_harness.py), and — when /root/reference is present —
against the live reference.

Each function cites the reference lines it restates (paths relative to the reference root):
  rep construction   source/encoder.py:183-265, source/decoder.py:247-353
  SO(2) tables       source/utils/gta.py:47-69
  scale mask         source/utils/gta.py:40-44
  Wigner-D           source/utils/wigner_d.py:16-58 (+ J_dense.pt[1], [2])
  rep application    source/utils/gta.py:92-279
  softmax attention  source/layers.py:202-211
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

# J_dense.pt[1] and [2] (exact values; symmetric involutions) — wigner_d.py:30.
_S3 = math.sqrt(3.0) / 2.0
J1 = [[0.0, -1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 0.0, 1.0]]
J2 = [[0.0, 0.0, 0.0, -1.0, 0.0],
      [0.0, 1.0, 0.0, 0.0, 0.0],
      [0.0, 0.0, -0.5, 0.0, -_S3],
      [-1.0, 0.0, 0.0, 0.0, 0.0],
      [0.0, 0.0, -_S3, 0.0, 0.5]]
EULER_EPS = 1e-5  # wigner_d.py:37


def zyz_euler(R: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """ZYZ Euler angles with the two gimbal cases (wigner_d.py:39-49)."""
    g1 = torch.atan2(R[..., 2, 1], -R[..., 2, 0])
    g2 = torch.atan2(torch.sqrt(R[..., 0, 2] ** 2 + R[..., 1, 2] ** 2), R[..., 2, 2])
    g3 = torch.atan2(R[..., 1, 2], R[..., 0, 2])
    up = (R[..., 2, 2] - 1).abs() < EULER_EPS
    dn = (R[..., 2, 2] + 1).abs() < EULER_EPS
    g1 = torch.where(up, torch.atan2(R[..., 1, 0], R[..., 0, 0]), g1)
    g1 = torch.where(dn, torch.atan2(-R[..., 1, 0], -R[..., 0, 0]), g1)
    g3 = torch.where(up | dn, torch.zeros_like(g3), g3)
    return g1, g2, g3


def z_rot(angle: torch.Tensor, l: int) -> torch.Tensor:
    """(2l+1)x(2l+1): cos(f_i a) on the diagonal, sin(f_i a) on the anti-diagonal, f = l..-l
    (wigner_d.py:16-25)."""
    n = 2 * l + 1
    f = torch.arange(l, -l - 1, -1, dtype=angle.dtype, device=angle.device)
    m = angle.new_zeros(angle.shape + (n, n))
    idx = torch.arange(n, device=angle.device)
    th = angle[..., None] * f
    m[..., idx, n - 1 - idx] = torch.sin(th)
    m[..., idx, idx] = torch.cos(th)
    return m


def wigner_d(l: int, g1, g2, g3) -> torch.Tensor:
    """D_l = Z(g3) J Z(g2) J Z(g1)  (wigner_d.py:28-35)."""
    J = torch.tensor(J1 if l == 1 else J2, dtype=g1.dtype, device=g1.device)
    return z_rot(g3, l) @ J @ z_rot(g2, l) @ J @ z_rot(g1, l)


def so2_angles(coord: torch.Tensor, nfreqs: int, max_freqs, shared: bool) -> torch.Tensor:
    """theta[..., j*naxes + axis] = max_freq_axis * 2*pi * coord_axis * freq_j,
    freq_j = 2^(j+1) / 2^nfreqs for j = 0..nfreqs-1 (gta.py:57-63).  NB the pair index is
    frequency-major / axis-minor: make_SO2mats stacks the per-axis tables at dim -3, i.e. *after*
    the frequency dim (gta.py:68), and the callers flatten (freq, axis) (encoder.py:195)."""
    if shared:
        freqs = torch.ones(nfreqs, dtype=coord.dtype, device=coord.device)
    else:
        freqs = 2.0 ** torch.arange(1, nfreqs + 1, dtype=coord.dtype, device=coord.device) / 2.0 ** nfreqs
    th = [max_freqs[d] * 2 * math.pi * coord[..., d:d + 1] * freqs for d in range(coord.shape[-1])]
    return torch.stack(th, -1).flatten(-2, -1)


def build_reps(cfg, extr_q, extr_k, coord_q, coord_k) -> Dict[str, torch.Tensor]:
    """Everything pre_compute_reps produces for this path, in a flat dict."""
    triv, se3, so3, so2 = cfg.dims()
    r: Dict[str, torch.Tensor] = {}
    if se3:
        r["se3_k"] = torch.linalg.inv(extr_k)          # encoder.py:219
        r["se3_qinv"] = extr_q                          # inv_se3rep_q, encoder.py:236
    if so3:
        for side, E in (("q", extr_q), ("k", extr_k)):
            R = torch.linalg.inv(E)[..., :3, :3]        # encoder.py:247
            g = zyz_euler(R)
            r["so3_d1_" + side] = wigner_d(1, *g)
            r["so3_d2_" + side] = wigner_d(2, *g)
    if so2:
        mf = [cfg.max_freq_h, cfg.max_freq_w]
        r["so2_th_q"] = so2_angles(coord_q, cfg.so2, mf, cfg.shared_freqs)   # [B,T,2n]
        r["so2_th_k"] = so2_angles(coord_k, cfg.so2, mf, cfg.shared_freqs)
    if se3 and getattr(cfg, "euclid", False):
        r["se3_q"] = torch.linalg.inv(extr_q)          # se3rep_q, encoder.py:235 (used un-transposed by the euclid branch)
    if cfg.t2_dim():
        r["t2_q"], r["t2_k"] = t2_mats(coord_q), t2_mats(coord_k)             # encoder.py:208-213
        r["t2_qinv"] = torch.linalg.inv(r["t2_q"])
    return r


def _scale_translation(M: torch.Tensor, tc) -> torch.Tensor:
    """M * scale_mask(tc): the translation column (rows 0..2 of column 3) is multiplied by tc
    (gta.py:40-44,140-141)."""
    one = torch.tensor([[1.0, 1.0, 1.0, 0.0]] * 3 + [[0.0, 0.0, 0.0, 1.0]], dtype=M.dtype, device=M.device)
    col = torch.tensor([[0.0, 0.0, 0.0, 1.0]] * 3 + [[0.0, 0.0, 0.0, 0.0]], dtype=M.dtype, device=M.device)
    return M * (one + col * tc)        # out of place: differentiable in tc (a learnable parameter, layers.py:188-191)


def _per_view(x: torch.Tensor, N: int) -> torch.Tensor:
    B, H, T, C = x.shape
    return x.reshape(B, H, N, T // N, C)


def t2_mats(coord: torch.Tensor) -> torch.Tensor:
    """[[1,0,0],[0,1,0],[x,y,1]] per token (make_T2mats, gta.py:72-89)."""
    m = torch.eye(3, dtype=coord.dtype, device=coord.device).repeat(*coord.shape[:-1], 1, 1)
    m[..., 2, 0] = coord[..., 0]
    m[..., 2, 1] = coord[..., 1]
    return m


def _apply_blocks(x, cfg, se3_mat, d1, d2, th, inverse_so2: bool, N: int, t2_mat=None):
    """Block-diagonal rep applied to x [B,H,T,D]; se3_mat [B,N,4,4], d1 [B,N,3,3], d2 [B,N,5,5],
    th [B,T,C], t2_mat [B,T,3,3] (gta.py:127-242 for q/k/v, :246-276 for the output)."""
    triv, se3, so3, so2 = cfg.dims()
    t2 = cfg.t2_dim()
    B, H, T, D = x.shape
    parts: List[torch.Tensor] = []
    o = 0
    if triv:
        parts.append(x[..., :triv]); o += triv
    if se3 and getattr(cfg, "euclid", False):
        # homogenised 3-vectors; the appended coordinate is dropped again (gta.py:146-156, 251-253)
        xs = _per_view(x[..., o:o + se3], N).reshape(B, H, N, T // N, se3 // 3, 3)
        xs = torch.cat([xs, torch.ones_like(xs[..., :1])], -1)
        ys = torch.einsum("bnij,bhntcj->bhntci", se3_mat, xs)[..., :-1]
        parts.append(ys.reshape(B, H, T, se3)); o += se3
    elif se3:
        xs = _per_view(x[..., o:o + se3], N).reshape(B, H, N, T // N, se3 // 4, 4)
        ys = torch.einsum("bnij,bhntcj->bhntci", se3_mat, xs)
        parts.append(ys.reshape(B, H, T, se3)); o += se3
    if so3:
        xs = _per_view(x[..., o:o + so3], N).reshape(B, H, N, T // N, so3 // 8, 8)
        y1 = torch.einsum("bnij,bhntcj->bhntci", d1, xs[..., 0:3])
        y2 = torch.einsum("bnij,bhntcj->bhntci", d2, xs[..., 3:8])
        parts.append(torch.cat([y1, y2], -1).reshape(B, H, T, so3)); o += so3
    if so2:
        xs = x[..., o:o + so2].reshape(B, H, T, so2 // 2, 2)
        c, s = torch.cos(th)[:, None], torch.sin(th)[:, None]
        if inverse_so2:
            s = -s
        y0 = c * xs[..., 0] - s * xs[..., 1]
        y1 = s * xs[..., 0] + c * xs[..., 1]
        parts.append(torch.stack([y0, y1], -1).reshape(B, H, T, so2)); o += so2
    if t2:
        xs = x[..., o:o + t2].reshape(B, H, T, t2 // 3, 3)
        ys = torch.einsum("btij,bhtcj->bhtci", t2_mat, xs)                    # t2fn, encoder.py:214
        parts.append(ys.reshape(B, H, T, t2)); o += t2
    return torch.cat(parts, -1)


def transform_qkv(cfg, q, k, v, reps, trans_coeff):
    """Returns (q', k', v') = (rho_q^{-T} q, rho_k k, rho_k v)  (gta.py:134-242)."""
    triv, se3, so3, so2 = cfg.dims()
    Nq, Nk = cfg.n_q_views, cfg.n_k_views
    Aq = Ak = None
    if se3 and getattr(cfg, "euclid", False):
        Aq = _scale_translation(reps["se3_q"], trans_coeff)                        # c_q, gta.py:140,153
        Ak = _scale_translation(reps["se3_k"], trans_coeff)
    elif se3:
        Aq = _scale_translation(reps["se3_qinv"], trans_coeff).transpose(-1, -2)   # gta.py:165
        Ak = _scale_translation(reps["se3_k"], trans_coeff)                        # gta.py:166
    d = lambda n: reps.get(n)
    t2q = reps["t2_qinv"].transpose(-1, -2) if cfg.t2_dim() else None              # gta.py:233
    qt = _apply_blocks(q, cfg, Aq, d("so3_d1_q"), d("so3_d2_q"), d("so2_th_q"), False, Nq, t2q)
    kt = _apply_blocks(k, cfg, Ak, d("so3_d1_k"), d("so3_d2_k"), d("so2_th_k"), False, Nk, d("t2_k"))
    vt = _apply_blocks(v, cfg, Ak, d("so3_d1_k"), d("so3_d2_k"), d("so2_th_k"), False, Nk, d("t2_k")) \
        if cfg.v_transform else v
    return qt, kt, vt


def gta_attention(cfg, q, k, v, extr_q, extr_k, coord_q, coord_k, trans_coeff=0.01, tau=1.0,
                  reps=None):
    """Full forward: O = rho_q^{-1} softmax(Q'K'^T * d^-1/2 / tau) V'   (gta.py:92-279 + layers.py:207-211)."""
    if reps is None:
        reps = build_reps(cfg, extr_q, extr_k, coord_q, coord_k)
    qt, kt, vt = transform_qkv(cfg, q, k, v, reps, trans_coeff)
    scale = cfg.head_dim ** -0.5
    sim = qt @ kt.transpose(-1, -2)
    if getattr(cfg, "euclid", False):          # EuclidAttnFn, layers.py:219-223
        sim = sim - 0.5 * qt.pow(2).sum(-1)[..., None] - 0.5 * kt.pow(2).sum(-1)[..., None, :]
    attn = torch.softmax(sim * (scale / tau), dim=-1)
    out = attn @ vt
    if not cfg.v_transform:
        return out
    triv, se3, so3, so2 = cfg.dims()
    Ao = _scale_translation(reps["se3_qinv"], trans_coeff) if se3 else None         # gta.py:255-257
    d1t = reps["so3_d1_q"].transpose(-1, -2) if so3 else None                       # gta.py:188
    d2t = reps["so3_d2_q"].transpose(-1, -2) if so3 else None
    return _apply_blocks(out, cfg, Ao, d1t, d2t, reps.get("so2_th_q"), True, cfg.n_q_views, reps.get("t2_qinv"))


def so2_mats(th: torch.Tensor) -> torch.Tensor:
    """[[cos,-sin],[sin,cos]] per angle — the layout make_SO2mats returns (gta.py:64-68)."""
    c, s = torch.cos(th), torch.sin(th)
    return torch.stack([torch.stack([c, -s], -1), torch.stack([s, c], -1)], -2)
