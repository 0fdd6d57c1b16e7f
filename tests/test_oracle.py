"""CPU tests: the oracle restatements (oracle/torch_port.py, oracle/gta_oracle.c) against the golden
vectors produced by the unmodified reference, plus the invariants of SURVEY.md §4."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from gta_b200.synth import (CFG1_A, CFG1_B, CLEVR, CLEVR_EUCLID, CLEVR_T2, MSN_SO3, MSN_SO3_EUCLID, MSN_T2, GtaConfig,
                            make_inputs)
from oracle import c_oracle, ref_harness, torch_port as tp
from tests.golden.gen_golden import ABLATION_CASES, CASES

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
T = torch.from_numpy


def _case(name):
    for c in CASES + ABLATION_CASES:
        if c[0] == name:
            _, base, nq, nk, tq, tk, cross, B, tc, seed, vt = c
            return GtaConfig(**base, n_q_views=nq, n_k_views=nk, v_transform=vt), cross
    raise KeyError(name)


@pytest.mark.parametrize("name", [c[0] for c in CASES + ABLATION_CASES])
def test_torch_port_matches_golden(name):
    cfg, cross = _case(name)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = tp.gta_attention(cfg, T(g["q"]), T(g["k"]), T(g["v"]), T(g["extr_q"]), T(g["extr_k"]),
                           T(g["coord_q"]), T(g["coord_k"]), trans_coeff=float(g["trans_coeff"]))
    assert np.abs(out.numpy() - g["out"]).max() < 2e-6


@pytest.mark.parametrize("name", [c[0] for c in CASES + ABLATION_CASES])
def test_c_oracle_matches_golden(name):
    cfg, cross = _case(name)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = c_oracle.gta_attention(cfg, g["q"], g["k"], g["v"], g["extr_q"], g["extr_k"], g["coord_q"],
                                 g["coord_k"], trans_coeff=float(g["trans_coeff"]))
    assert np.abs(out - g["out"]).max() < 5e-6 * max(1.0, float(np.abs(g["out"]).max()))
    if cfg.t2_dim() or cfg.euclid:
        return          # the packed rep tables below belong to the fused-path layouts
    r = c_oracle.build_reps(cfg, g["extr_q"], g["extr_k"], g["coord_q"], g["coord_k"])
    if "ref_se3rep_k" in g:
        assert np.abs(r["se3_k"].reshape(g["ref_se3rep_k"].shape) - g["ref_se3rep_k"]).max() < 1e-6
        assert np.abs(r["se3_q"].reshape(g["ref_inv_se3rep_q"].shape) - g["ref_inv_se3rep_q"]).max() < 1e-6
    if "ref_so3rep_k_d1" in g:
        B, N = g["extr_k"].shape[:2]
        assert np.abs(r["so3_k"][..., :9].reshape(B, N, 3, 3) - g["ref_so3rep_k_d1"]).max() < 2e-6
        assert np.abs(r["so3_k"][..., 9:].reshape(B, N, 5, 5) - g["ref_so3rep_k_d2"]).max() < 2e-6
        Bq, Nq = g["extr_q"].shape[:2]
        assert np.abs(r["so3_q"][..., 9:].reshape(Bq, Nq, 5, 5) - g["ref_so3rep_q_d2"]).max() < 2e-6
    if "ref_so2rep_q" in g:
        m = g["ref_so2rep_q"]                       # [B,T,C,2,2] = [[c,-s],[s,c]]
        assert np.abs(r["so2_q"][..., 0] - m[..., 0, 0]).max() < 2e-6
        assert np.abs(r["so2_q"][..., 1] - m[..., 1, 0]).max() < 2e-6
        m = g["ref_so2rep_k"]
        assert np.abs(r["so2_k"][..., 1] + m[..., 0, 1]).max() < 2e-6


def test_reps_gimbal_golden():
    g = np.load(os.path.join(GOLDEN, "reps_gimbal.npz"))
    E = T(g["extr"])
    R = torch.linalg.inv(E)[..., :3, :3].flatten(0, 1)
    ang = tp.zyz_euler(R)
    assert (tp.wigner_d(1, *ang) - T(g["d1"])).abs().max() < 1e-6
    assert (tp.wigner_d(2, *ang) - T(g["d2"])).abs().max() < 1e-6
    cfg = GtaConfig(**MSN_SO3, n_q_views=5, n_k_views=5)
    r = c_oracle.build_reps(cfg, g["extr"], g["extr"], g["coord"], g["coord"])
    assert np.abs(r["so3_k"][0, :, :9].reshape(5, 3, 3) - g["d1"]).max() < 2e-6
    assert np.abs(r["so3_k"][0, :, 9:].reshape(5, 5, 5) - g["d2"]).max() < 2e-6
    assert np.abs(r["se3_k"].reshape(1, 5, 4, 4) - g["inv"]).max() < 1e-6
    m = g["so2_n6"]
    assert np.abs(r["so2_k"][..., 0] - m[..., 0, 0]).max() < 2e-6
    assert np.abs(r["so2_k"][..., 1] - m[..., 1, 0]).max() < 2e-6
    th = tp.so2_angles(T(g["coord"]), 3, [2, 0.5], True)
    assert (tp.so2_mats(th) - T(g["so2_n3_shared_f2_05"])).abs().max() < 1e-6
    from gta_b200.synth import make_2dcoord
    assert np.array_equal(make_2dcoord(5, 7), g["coord2d_5x7"])


def _run64(cfg, inp, tc=0.01):
    d = lambda t: t.double()
    return tp.gta_attention(cfg, d(inp["q"]), d(inp["k"]), d(inp["v"]), d(inp["extr_q"]), d(inp["extr_k"]),
                            d(inp["coord_q"]), d(inp["coord_k"]), trans_coeff=tc)


def test_identity_pose_zero_coord_is_plain_attention():
    cfg = GtaConfig(**MSN_SO3, n_q_views=2, n_k_views=2)
    inp = make_inputs(cfg, 1, 9, 9, cross=False, seed=1)
    inp["extr_q"] = inp["extr_k"] = torch.eye(4).expand(1, 2, 4, 4).contiguous()
    inp["coord_q"] = inp["coord_k"] = torch.zeros(1, 18, 2)
    out = _run64(cfg, inp)
    q, k, v = (inp[n].double() for n in "qkv")
    ref = torch.softmax(q @ k.transpose(-1, -2) * cfg.head_dim ** -0.5, -1) @ v
    assert (out - ref).abs().max() < 1e-12
    o2 = c_oracle.gta_attention(cfg, inp["q"], inp["k"], inp["v"], inp["extr_q"], inp["extr_k"],
                                inp["coord_q"], inp["coord_k"])
    assert np.abs(o2 - ref.float().numpy()).max() < 1e-6


def test_frame_and_shift_invariance():
    cfg = GtaConfig(**MSN_SO3, n_q_views=3, n_k_views=2)
    inp = make_inputs(cfg, 2, 6, 10, cross=True, seed=2)
    base = _run64(cfg, inp, tc=1.0)
    from gta_b200.synth import random_extrinsics
    G = random_extrinsics(torch.Generator().manual_seed(9), 1, 1, first_identity=False)[0, 0].double()
    inp2 = dict(inp)
    inp2["extr_q"] = (inp["extr_q"].double() @ G).float()
    inp2["extr_k"] = (inp["extr_k"].double() @ G).float()
    assert (_run64(cfg, inp2, tc=1.0) - base).abs().max() < 2e-5      # fp32-rounded extrinsics
    inp3 = dict(inp)
    shift = torch.tensor([0.125, 0.5])   # exactly representable: so2 periods are 2^n / 2^j
    inp3["coord_q"] = inp["coord_q"] + shift
    inp3["coord_k"] = inp["coord_k"] + shift
    assert (_run64(cfg, inp3, tc=1.0) - base).abs().max() < 1e-5


def test_wigner_properties():
    gen = torch.Generator().manual_seed(3)
    from gta_b200.synth import random_extrinsics
    R1 = random_extrinsics(gen, 1, 6, False)[0, :, :3, :3].double()
    R2 = random_extrinsics(gen, 1, 6, False)[0, :, :3, :3].double()
    for l in (1, 2):
        D1 = tp.wigner_d(l, *tp.zyz_euler(R1)); D2 = tp.wigner_d(l, *tp.zyz_euler(R2))
        D12 = tp.wigner_d(l, *tp.zyz_euler(R1 @ R2))
        I = torch.eye(2 * l + 1, dtype=torch.float64)
        assert (D1 @ D1.transpose(-1, -2) - I).abs().max() < 1e-6   # R comes from fp32 extrinsics
        assert (D1 @ D2 - D12).abs().max() < 1e-6
    P = [1, 2, 0]   # D_1(R) = P R P^T with basis order (y, z, x)
    D = tp.wigner_d(1, *tp.zyz_euler(R1))
    assert (D - R1[:, P][:, :, P]).abs().max() < 1e-6


def test_c_oracle_rotated_tensors_match_port():
    cfg = GtaConfig(**CFG1_B, n_q_views=2, n_k_views=2)
    inp = make_inputs(cfg, 1, 16, 16, cross=False, seed=4)
    out, qt, kt, vt = c_oracle.gta_attention(cfg, inp["q"], inp["k"], inp["v"], inp["extr_q"], inp["extr_k"],
                                             inp["coord_q"], inp["coord_k"], return_rotated=True)
    reps = tp.build_reps(cfg, inp["extr_q"], inp["extr_k"], inp["coord_q"], inp["coord_k"])
    q2, k2, v2 = tp.transform_qkv(cfg, inp["q"], inp["k"], inp["v"], reps, 0.01)
    for a, b in ((qt, q2), (kt, k2), (vt, v2)):
        assert np.abs(a - b.numpy()).max() < 2e-6


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("base,nq,nk,tq,tk,cross", [
    (CFG1_A, 2, 2, 64, 64, False), (MSN_SO3, 5, 5, 16, 16, False), (CLEVR, 3, 2, 11, 30, True)])
def test_against_live_reference(base, nq, nk, tq, tk, cross):
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk)
    inp = make_inputs(cfg, 2, tq, tk, cross=cross, seed=21)
    ref, _ = ref_harness.ref_gta_attention(cfg, inp, trans_coeff=0.3)
    out = c_oracle.gta_attention(cfg, inp["q"], inp["k"], inp["v"], inp["extr_q"], inp["extr_k"],
                                 inp["coord_q"], inp["coord_k"], trans_coeff=0.3)
    assert np.abs(out - ref.numpy()).max() < 5e-6
    o2 = tp.gta_attention(cfg, inp["q"], inp["k"], inp["v"], inp["extr_q"], inp["extr_k"],
                          inp["coord_q"], inp["coord_k"], trans_coeff=0.3)
    assert (o2 - ref).abs().max() < 2e-6


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("base,nq,nk,tq,tk,cross,vt", [
    (CLEVR_T2, 2, 2, 30, 30, False, True), (MSN_T2, 3, 2, 8, 16, True, False), (CLEVR_EUCLID, 3, 2, 11, 30, True, True),
    (MSN_SO3_EUCLID, 5, 5, 16, 16, False, True), (MSN_SO3_EUCLID, 3, 2, 8, 16, True, False)])
def test_ablation_blocks_against_live_reference(base, nq, nk, tq, tk, cross, vt):
    """t2 block and euclid_sim: the torch restatement against the unmodified reference function."""
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk, v_transform=vt)
    inp = make_inputs(cfg, 2, tq, tk, cross=cross, seed=22)
    ref, ex = ref_harness.ref_gta_attention(cfg, inp, trans_coeff=0.3)
    o2 = tp.gta_attention(cfg, inp["q"], inp["k"], inp["v"], inp["extr_q"], inp["extr_k"],
                          inp["coord_q"], inp["coord_k"], trans_coeff=0.3)
    assert (o2 - ref).abs().max() < 2e-6
    o3 = c_oracle.gta_attention(cfg, inp["q"], inp["k"], inp["v"], inp["extr_q"], inp["extr_k"],
                                inp["coord_q"], inp["coord_k"], trans_coeff=0.3)
    assert np.abs(o3 - ref.numpy()).max() < 5e-6
    if cfg.t2_dim():
        assert (tp.t2_mats(inp["coord_k"]) - ex["t2rep_k"]).abs().max() == 0


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("base,nq,nk,tq,tk,cross", [(MSN_SO3, 3, 2, 8, 16, True), (CLEVR, 2, 2, 21, 21, False),
                                                    (CLEVR_T2, 2, 2, 21, 21, False), (CLEVR_EUCLID, 3, 2, 8, 16, True),
                                                    (MSN_SO3_EUCLID, 2, 2, 9, 9, False)])
def test_grad_oracle_matches_reference_autograd(base, nq, nk, tq, tk, cross):
    """The gradient checker of the GPU tests (autograd through oracle/torch_port.py) against autograd through the
    UNMODIFIED reference function: dq, dk, dv and d(trans_coeff)."""
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk)
    inp = make_inputs(cfg, 2, tq, tk, cross=cross, seed=23)
    dout = torch.randn(inp["q"].shape, generator=torch.Generator().manual_seed(1))
    grads = []
    for impl in ("ref", "port"):
        q, k, v = (inp[n].clone().requires_grad_(True) for n in "qkv")
        tc = torch.tensor([0.3], requires_grad=True)
        if impl == "ref":
            m = ref_harness.load()
            extras = ref_harness.ref_reps(cfg, inp["extr_q"], inp["extr_k"], inp["coord_q"], inp["coord_k"], cross)
            out, _ = m.gta.multihead_geometric_transform_attention(
                q, k, v, attn_fn=ref_harness._AttnFn(cfg.head_dim ** -0.5, euclid=cfg.euclid), f_dims=dict(cfg.f_dims), reps=extras,
                trans_coeff=tc, v_transform=True, euclid=cfg.euclid)
        else:
            out = tp.gta_attention(cfg, q, k, v, inp["extr_q"], inp["extr_k"], inp["coord_q"], inp["coord_k"], trans_coeff=tc)
        out.backward(dout)
        grads.append((q.grad, k.grad, v.grad, tc.grad))
    for a, b in zip(*grads):
        assert (a - b).abs().max() < 5e-6 * max(1.0, float(b.abs().max()))


@pytest.mark.parametrize("name", [c[0] for c in CASES + ABLATION_CASES])
def test_dropin_packs_reference_format_reps(name):
    """Host logic of the drop-in (gta_b200.gta._pack_reps): the rep tensors exactly as the reference's pre_compute_reps
    left them in `extras` (stored in the golden files) are packed into the tables the CUDA library reads — checked
    against the C oracle's tables built from the raw poses / coordinates."""
    from gta_b200 import gta as fast
    cfg, cross = _case(name)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    extras = {}
    for key in ("se3rep_q", "se3rep_k", "inv_se3rep_q", "so2rep_q", "so2rep_k", "t2rep_q", "t2rep_k", "inv_t2rep_q"):
        if "ref_" + key in g:
            extras[key] = T(g["ref_" + key])
    for side in ("q", "k"):
        if f"ref_so3rep_{side}_d1" in g:
            extras[f"so3rep_{side}"] = [T(g[f"ref_so3rep_{side}_d1"]), T(g[f"ref_so3rep_{side}_d2"])]
    if not cross:       # self-attention: the reference stores the same objects under the _q and _k keys (encoder.py:196,213,235)
        for key in ("so2rep", "t2rep"):
            if key + "_q" in extras:
                extras[key + "_k"] = extras[key + "_q"]
    B = g["q"].shape[0]
    p = fast._pack_reps(extras, cfg.f_dims, B, cfg.euclid)
    r = c_oracle.build_reps(cfg, g["extr_q"], g["extr_k"], g["coord_q"], g["coord_k"])
    triv, se3, so3, so2 = cfg.dims()
    if se3:
        assert np.abs(p.se3_q.numpy() - r["se3_q"]).max() < 1e-6 and np.abs(p.se3_k.numpy() - r["se3_k"]).max() < 1e-6
        assert (p.n_q_views, p.n_k_views) == (cfg.n_q_views, cfg.n_k_views)
        if cfg.euclid:
            assert np.abs(p.se3_qi.numpy() - np.linalg.inv(g["extr_q"].astype(np.float64)).reshape(B, -1, 16)).max() < 1e-5
    if so3:
        assert np.abs(p.so3_q.numpy() - r["so3_q"]).max() < 2e-6 and np.abs(p.so3_k.numpy() - r["so3_k"]).max() < 2e-6
    if so2:
        assert np.abs(p.so2_q.numpy() - r["so2_q"]).max() < 2e-6 and np.abs(p.so2_k.numpy() - r["so2_k"]).max() < 2e-6
        assert (p.so2_q is p.so2_k) == (not cross)
    if cfg.t2_dim():
        assert np.array_equal(p.t2_q.numpy(), g["coord_q"].reshape(B, -1, 2))
        assert np.array_equal(p.t2_k.numpy(), g["coord_k"].reshape(B, -1, 2))
    assert extras[fast._PACK_KEY][1] is p and fast._pack_reps(extras, cfg.f_dims, B, cfg.euclid) is p    # cached


@pytest.mark.skipif(not ref_harness.available(), reason="reference tree not mounted (GPU box)")
def test_install_rebinds_reference_globals_and_delegates_cpu_calls():
    """gta_b200.gta.install(): source.layers resolves the operator by name at call time (layers.py:6,419), so rebinding
    the module globals swaps the implementation; a CPU call is delegated to the captured reference function (same
    result), and uninstall() restores the originals."""
    import warnings
    from gta_b200 import gta as fast
    m = ref_harness.load()
    import source.layers as rlayers
    orig = m.gta.multihead_geometric_transform_attention
    cfg = GtaConfig(**MSN_SO3, n_q_views=2, n_k_views=2)
    inp = make_inputs(cfg, 1, 9, 9, cross=False, seed=3)
    extras = ref_harness.ref_reps(cfg, inp["extr_q"], inp["extr_k"], inp["coord_q"], inp["coord_k"], False)
    fn = ref_harness._AttnFn(cfg.head_dim ** -0.5)
    tc = torch.tensor([0.05])
    ref, _ = orig(inp["q"], inp["k"], inp["v"], attn_fn=fn, f_dims=dict(cfg.f_dims), reps=extras, trans_coeff=tc)
    try:
        assert fast.install() is orig
        assert rlayers.multihead_geometric_transform_attention is fast.multihead_geometric_transform_attention
        assert m.gta.multihead_geometric_transform_attention is fast.multihead_geometric_transform_attention
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            out, attn = rlayers.multihead_geometric_transform_attention(
                inp["q"], inp["k"], inp["v"], attn_fn=fn, f_dims=dict(cfg.f_dims), reps=extras, trans_coeff=tc)
        assert any("delegating" in str(x.message) for x in w)
        assert torch.equal(out, ref) and attn is not None
    finally:
        fast.uninstall()
    assert rlayers.multihead_geometric_transform_attention is orig and m.gta.multihead_geometric_transform_attention is orig
    with pytest.raises(NotImplementedError):      # without install() there is nothing to delegate to — and no CPU fallback
        fast.multihead_geometric_transform_attention(inp["q"], inp["k"], inp["v"], fn, dict(cfg.f_dims), extras, trans_coeff=tc)
