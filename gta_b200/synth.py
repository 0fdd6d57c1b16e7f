"""Seeded synthetic inputs for the GTA-attention hot path.

Everything here is plain numpy/torch on the host; it is shared by the tests, the
golden-vector generator and bench.py so that every consumer sees the same tensors.

Shapes follow the reference:
  q        [B, H, Tq, D]   (source/layers.py:394-395; strided views of the projection output)
  k, v     [B, H, Tk, D]
  extr_q   [B, Nq, 4, 4]   camera extrinsics of the query views  (extras['input_transforms'] /
  extr_k   [B, Nk, 4, 4]    extras['target_transforms'], source/encoder.py:218, decoder.py:293)
  coord_q  [B, Tq, 2]      patch coordinates in [0,1)             (extras['input_coord'/'target_coord'])
  coord_k  [B, Tk, 2]
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional

import numpy as np
import torch


@dataclasses.dataclass(frozen=True)
class GtaConfig:
    """Static description of one GTA attention call (mirrors attn_args.method.args of the YAMLs)."""
    heads: int
    head_dim: int
    f_dims: Dict[str, int]          # order inside a head is fixed: triv | se3 | so3 | so2 | t2 (gta.py:115)
    so2: int = 0                    # number of SO(2) frequencies per coordinate axis
    so3: int = 0                    # max Wigner-D degree (only 2 is used by the shipped configs)
    max_freq_h: float = 1.0
    max_freq_w: float = 1.0
    shared_freqs: bool = False
    v_transform: bool = True
    euclid: bool = False            # euclid_sim: se3 block = homogenised 3-vectors, similarity -|q'-k'|^2/2 (layers.py:213-231)
    n_q_views: int = 1
    n_k_views: int = 1
    name: str = ""

    def dims(self):
        g = lambda k: int(self.f_dims.get(k, 0) or 0)
        return g("triv"), g("se3"), g("so3"), g("so2")

    def t2_dim(self):
        return int(self.f_dims.get("t2", 0) or 0)

    def validate(self):
        triv, se3, so3, so2 = self.dims()
        t2 = self.t2_dim()
        assert triv + se3 + so3 + so2 + t2 == self.head_dim, "f_dims must sum to head_dim"
        assert se3 % (3 if self.euclid else 4) == 0 and so3 % 8 == 0 and so2 % 2 == 0 and t2 % 3 == 0
        if so2:
            assert so2 == 4 * self.so2, "so2 dims must equal 2 axes * nfreqs * 2"
        if so3:
            assert self.so3 == 2 and se3 > 0, "so3 needs max degree 2 and an se3 block (SURVEY T5)"


# The shipped configs the benchmark is quoted on (runs/*/GTA/*/config.yaml).
MSN_SO3 = dict(heads=8, head_dim=96, f_dims={"triv": 0, "se3": 48, "so3": 24, "so2": 24}, so2=6, so3=2)
CLEVR = dict(heads=6, head_dim=64, f_dims={"se3": 32, "so2": 32}, so2=8, so3=0)
CFG1_A = dict(heads=4, head_dim=32, f_dims={"se3": 16, "so2": 16}, so2=4, so3=0)
CFG1_B = dict(heads=4, head_dim=32, f_dims={"se3": 16, "so3": 8, "so2": 8}, so2=2, so3=2)
# ablation configs served by the generic path (runs/clevrtr/GTA/gta_t2, gta_euclid; runs/msn/GTA/gta_t2, gta_so3_euclid)
CLEVR_T2 = dict(heads=6, head_dim=64, f_dims={"triv": 2, "se3": 32, "t2": 30}, so2=0, so3=0)
CLEVR_EUCLID = dict(heads=6, head_dim=64, f_dims={"triv": 2, "se3": 30, "so2": 32}, so2=8, so3=0, euclid=True)
MSN_T2 = dict(heads=8, head_dim=96, f_dims={"triv": 0, "se3": 48, "t2": 48}, so2=0, so3=0)
MSN_SO3_EUCLID = dict(heads=8, head_dim=96, f_dims={"triv": 0, "se3": 48, "so3": 24, "so2": 24}, so2=6, so3=2, euclid=True)


def make_2dcoord(H: int, W: int) -> np.ndarray:
    """coord[i, j] = (i/H, j/W); same values as the reference helper (source/utils/gta.py:9-16)."""
    ii = (np.arange(H, dtype=np.float32) / np.float32(H))[:, None]
    jj = (np.arange(W, dtype=np.float32) / np.float32(W))[None, :]
    out = np.empty((H, W, 2), dtype=np.float32)
    out[..., 0] = ii
    out[..., 1] = jj
    return out


def patch_coords(h: int, w: int, down: int = 8) -> np.ndarray:
    """Key-side coordinates: full-resolution grid sampled at stride//2::stride
    (source/utils/common.py:105-110 applied by multishapenet.py:169-174)."""
    full = make_2dcoord(down * h, down * w)
    return full[down // 2::down, down // 2::down].reshape(h * w, 2)


def random_extrinsics(gen: torch.Generator, B: int, N: int, first_identity: bool = True) -> torch.Tensor:
    """Rigid transforms: rotation = QR of a Gaussian matrix with det fixed to +1, translation ~ N(0,1)."""
    A = torch.randn(B, N, 3, 3, generator=gen, dtype=torch.float64)
    Q, R = torch.linalg.qr(A)
    Q = Q * torch.sign(torch.diagonal(R, dim1=-2, dim2=-1)).unsqueeze(-2)
    det = torch.linalg.det(Q)
    Q[..., :, 0] = Q[..., :, 0] * det[..., None]
    E = torch.zeros(B, N, 4, 4, dtype=torch.float64)
    E[..., :3, :3] = Q
    E[..., :3, 3] = torch.randn(B, N, 3, generator=gen, dtype=torch.float64)
    E[..., 3, 3] = 1.0
    if first_identity:
        E[:, 0] = torch.eye(4, dtype=torch.float64)
    return E.to(torch.float32)


def make_inputs(cfg: GtaConfig, B: int, tq_per_view: int, tk_per_view: int, *, cross: bool,
                seed: int = 0, grid_hw: Optional[tuple] = None, packed_layout: bool = True,
                dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    """Host tensors for one call.  Self-attention (cross=False): q/k/v are strided views of one
    [B, T, 3*H*D] buffer exactly as `to_qkv(x).chunk(3)` + rearrange delivers them
    (source/layers.py:389-395).  Cross-attention: q from [B,Tq,H*D], k/v from one [B,Tk,2*H*D]."""
    cfg.validate()
    gen = torch.Generator().manual_seed(seed)
    H, D = cfg.heads, cfg.head_dim
    Nq, Nk = cfg.n_q_views, cfg.n_k_views
    Tq, Tk = Nq * tq_per_view, Nk * tk_per_view

    def heads_view(x):  # 'b n (h d) -> b h n d' as a view
        return x.view(x.shape[0], x.shape[1], H, D).permute(0, 2, 1, 3)

    if not cross:
        assert Nq == Nk and tq_per_view == tk_per_view
        buf = torch.randn(B, Tq, 3 * H * D, generator=gen, dtype=torch.float32).to(dtype)
        if packed_layout:
            q, k, v = (heads_view(t) for t in buf.chunk(3, dim=-1))
        else:
            q, k, v = (heads_view(t).contiguous() for t in buf.chunk(3, dim=-1))
    else:
        bq = torch.randn(B, Tq, H * D, generator=gen, dtype=torch.float32).to(dtype)
        bkv = torch.randn(B, Tk, 2 * H * D, generator=gen, dtype=torch.float32).to(dtype)
        q = heads_view(bq)
        k, v = (heads_view(t) for t in bkv.chunk(2, dim=-1))
        if not packed_layout:
            q, k, v = q.contiguous(), k.contiguous(), v.contiguous()

    extr_k = random_extrinsics(gen, B, Nk)
    if cross:
        extr_q = random_extrinsics(gen, B, Nq, first_identity=False)
    else:
        extr_q = extr_k

    if grid_hw is None:
        s = int(round(math.sqrt(tk_per_view)))
        grid_hw = (s, s) if s * s == tk_per_view else None
    if grid_hw is not None:
        ck = torch.from_numpy(patch_coords(*grid_hw))            # [tk_per_view, 2]
    else:  # non-square view (e.g. CLEVR 15x20)
        ck = torch.rand(tk_per_view, 2, generator=gen)
    coord_k = ck[None, None].expand(B, Nk, tk_per_view, 2).reshape(B, Tk, 2).contiguous()
    if cross:
        coord_q = torch.rand(B, Tq, 2, generator=gen)
    else:
        coord_q = coord_k
    return dict(q=q, k=k, v=v, extr_q=extr_q, extr_k=extr_k, coord_q=coord_q, coord_k=coord_k)
