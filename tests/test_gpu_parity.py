"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs, against the committed golden vectors of the unmodified reference, and — at full benchmark
sizes — through size-independent properties (frame invariance, linearity in V, identity-pose == plain softmax).

Tolerances (BASELINE.json north_star), enforced as ABSOLUTE max-abs bounds: 1e-2 for the bf16 tensor-core path and 1e-3
for fp32 inputs, on N(0,1) inputs at the configured trans_coeff (0.01, and every case below 0.5).  Only for
trans_coeff >= 0.5 — where the O(1) camera translations enter the SE(3) features and the output grows with them — is
the bound relative to |ref|_max (and doubled for bf16); those cases print their measured error.  Every comparison
prints its measured max-abs error (pytest -s / -rP shows them)."""
import glob
import os

import numpy as np
import pytest
import torch

from gta_b200.synth import (CFG1_A, CFG1_B, CLEVR, CLEVR_EUCLID, CLEVR_T2, MSN_SO3, MSN_SO3_EUCLID, MSN_T2, GtaConfig,
                            make_inputs)

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
BF16_TOL = 1e-2
FP32_TOL = 1e-3
# GTA_FLAG_* pipeline selectors of the product library: the default two-launch pipeline (staging kernel + persistent
# attention kernel), the single-launch pipeline (K/V rotation by staging warps of the attention kernel), the
# streaming-softmax / epilogue-warpgroup kernel, the spare-P-buffer kernel, V1 (non-persistent two-tile)
PIPELINES, PIPELINE_IDS = [1024, 32, 256, 512, 16, 0], ["v2_two_launch", "v3_single_launch", "v4_streaming", "v5_spare_p", "v1", "auto"]


def _ops():
    from gta_b200 import ops
    return ops


def _dev_reps(cfg, inp):
    ops = _ops()
    ek, ck = inp["extr_k"].cuda(), inp["coord_k"].cuda()
    eq = ek if inp["extr_q"] is inp["extr_k"] else inp["extr_q"].cuda()
    cq = ck if inp["coord_q"] is inp["coord_k"] else inp["coord_q"].cuda()
    return ops.build_reps(eq, ek, cq, ck, so2_nfreqs=cfg.so2, so3_maxdeg=cfg.so3, max_freq_h=cfg.max_freq_h,
                          max_freq_w=cfg.max_freq_w, shared_freqs=cfg.shared_freqs, t2=bool(cfg.t2_dim()),
                          euclid=cfg.euclid)


def _run(cfg, inp, tc=0.01, flags=0, out_dtype=None):
    ops = _ops()
    reps = _dev_reps(cfg, inp)
    out = ops.gta_attention_fwd(inp["q"].cuda(), inp["k"].cuda(), inp["v"].cuda(), reps, cfg.f_dims,
                                trans_coeff=torch.tensor([tc], device="cuda"), v_transform=cfg.v_transform,
                                flags=flags, out_dtype=out_dtype, euclid=cfg.euclid)
    torch.cuda.synchronize()
    return out.float().cpu().numpy()


def _oracle(cfg, inp, tc=0.01):
    from oracle import c_oracle
    return c_oracle.gta_attention(cfg, inp["q"].float(), inp["k"].float(), inp["v"].float(), inp["extr_q"],
                                  inp["extr_k"], inp["coord_q"], inp["coord_k"], trans_coeff=tc)


def _tol(ref, tc=0.01, base=BF16_TOL):
    """Absolute north-star bound below trans_coeff 0.5; relative to |ref|_max (doubled for bf16) from there on, where the
    O(1) camera translations feed the se3 features and the output grows with them."""
    if tc < 0.5:
        return base
    return base * max(1.0, float(np.abs(ref).max())) * (2.0 if base == BF16_TOL else 1.0)


def _check(out, ref, tc=0.01, base=BF16_TOL, what="", rel=False):
    """rel=True: bound relative to max(1, |ref|_max) without the trans_coeff doubling (outputs well above 1, where the
    bf16 rounding of the OUTPUT alone is |out| * 2^-9)."""
    err = float(np.abs(out - ref).max())
    tol = base * max(1.0, float(np.abs(ref).max())) if rel else _tol(ref, tc, base)
    print("%s max-abs err %.3e (bound %.1e%s, |ref|max %.2f)" % (what, err, tol, "" if tc < 0.5 else " relative, tc=%g" % tc,
                                                              float(np.abs(ref).max())))
    assert np.isfinite(out).all()
    assert err < tol, (what, err, tol)
    return err


def test_umma_probe_exact():
    """tcgen05 descriptors / tile images: S = A B^T and O = P V are exact in fp32 accumulation."""
    ops = _ops()
    torch.manual_seed(0)
    for D in (32, 64, 96, 128):
        A = torch.randn(128, D, device="cuda").bfloat16(); Bm = torch.randn(128, D, device="cuda").bfloat16()
        P = torch.rand(128, 128, device="cuda").bfloat16(); V = torch.randn(128, D, device="cuda").bfloat16()
        for tm in (False, True):
            S, O = ops.umma_probe(A, Bm, P, V, tm)
            assert (S - A.float() @ Bm.float().T).abs().max() < 1e-4
            assert (O - P.float() @ V.float()).abs().max() < 1e-4


def test_build_reps_matches_oracle():
    from oracle import c_oracle
    cfg = GtaConfig(**MSN_SO3, n_q_views=3, n_k_views=5)
    inp = make_inputs(cfg, 2, 16, 64, cross=True, seed=1)
    r = _dev_reps(cfg, inp)
    o = c_oracle.build_reps(cfg, inp["extr_q"], inp["extr_k"], inp["coord_q"], inp["coord_k"])
    for k in ("se3_q", "se3_k", "so3_q", "so3_k", "so2_q", "so2_k"):
        assert np.abs(getattr(r, k).cpu().numpy() - o[k]).max() < 2e-6, k


def test_reps_golden_gimbal():
    ops = _ops()
    g = np.load(os.path.join(GOLDEN, "reps_gimbal.npz"))
    E = torch.from_numpy(g["extr"]).cuda()
    c = torch.from_numpy(g["coord"]).cuda()
    r = ops.build_reps(E, E, c, c, so2_nfreqs=6, so3_maxdeg=2)
    assert np.abs(r.so3_k[0, :, :9].reshape(5, 3, 3).cpu().numpy() - g["d1"]).max() < 2e-6
    assert np.abs(r.so3_k[0, :, 9:].reshape(5, 5, 5).cpu().numpy() - g["d2"]).max() < 2e-6
    assert np.abs(r.se3_k.reshape(1, 5, 4, 4).cpu().numpy() - g["inv"]).max() < 1e-6
    from gta_b200 import gta as fast, wigner_d as fw
    m = fast.make_SO2mats(c, 6).flatten(-4, -3)
    assert np.abs(m.cpu().numpy() - g["so2_n6"]).max() < 2e-6
    m = fast.make_SO2mats(c, 3, [2, 0.5], shared_freqs=True).flatten(-4, -3)
    assert np.abs(m.cpu().numpy() - g["so2_n3_shared_f2_05"]).max() < 2e-6
    R = torch.linalg.inv(torch.from_numpy(g["extr"]))[..., :3, :3].flatten(0, 1).cuda()
    D = fw.rotmat_to_wigner_d_matrices(2, R)
    assert np.abs(D[0].cpu().numpy() - g["d0"]).max() == 0
    assert np.abs(D[1].cpu().numpy() - g["d1"]).max() < 2e-6
    assert np.abs(D[2].cpu().numpy() - g["d2"]).max() < 2e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_rotated_operands_match_oracle(dtype):
    from oracle import c_oracle
    ops = _ops()
    cfg = GtaConfig(**MSN_SO3, n_q_views=3, n_k_views=2)
    inp = make_inputs(cfg, 2, 8, 16, cross=True, seed=2, dtype=dtype)
    reps = _dev_reps(cfg, inp)
    qt, kt, vt = ops.rotate_debug(inp["q"].cuda(), inp["k"].cuda(), inp["v"].cuda(), reps, cfg.f_dims,
                                  trans_coeff=torch.tensor([0.3], device="cuda"))
    _, q2, k2, v2 = c_oracle.gta_attention(cfg, inp["q"].float(), inp["k"].float(), inp["v"].float(), inp["extr_q"],
                                           inp["extr_k"], inp["coord_q"], inp["coord_k"], trans_coeff=0.3,
                                           return_rotated=True)
    for a, b in ((qt, q2), (kt, k2), (vt, v2)):
        assert np.abs(a.cpu().numpy() - b).max() < 5e-6


def _golden_names(ablation):
    from tests.golden.gen_golden import ABLATION_CASES, CASES
    return [c[0] for c in (ABLATION_CASES if ablation else CASES)]


@pytest.mark.parametrize("name", _golden_names(True))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_golden_vectors_ablation_blocks(name, dtype):
    """t2 block / euclid_sim / head layouts with blocks that are not multiples of 8 (generic path) against the committed
    outputs of the unmodified reference."""
    from tests.golden.gen_golden import ABLATION_CASES
    _, base, nq, nk, tq, tk, cross, B, tc, seed, vt = [c for c in ABLATION_CASES if c[0] == name][0]
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk, v_transform=vt)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    inp = {k: torch.from_numpy(g[k]) for k in ("q", "k", "v", "extr_q", "extr_k", "coord_q", "coord_k")}
    if not cross:
        inp["extr_q"], inp["coord_q"] = inp["extr_k"], inp["coord_k"]
    if dtype == torch.bfloat16:
        inp = dict(inp, **{n: inp[n].to(dtype) for n in "qkv"})
        ref = _oracle(cfg, inp, float(g["trans_coeff"]))          # same (rounded) inputs
    else:
        ref = g["out"]
    out = _run(cfg, inp, tc=float(g["trans_coeff"]))
    # fp32 inputs run the split-precision path (fp32 budget) unless the padded head dim of euclid_sim exceeds 96
    hp = dtype == torch.float32 and cfg.head_dim + (32 if cfg.euclid else 0) <= 96
    # euclid_sim logits carry the -|k'|^2/2 terms (tens of units, not O(1)): their bf16 products set the error, so the
    # bf16-math bound of that ablation is relative to |ref|_max like the trans_coeff >= 0.5 cases
    tc_eff = 1.0 if (cfg.euclid and not hp) else float(g["trans_coeff"])
    _check(out, ref, tc_eff, FP32_TOL if hp else BF16_TOL, name)


def _golden_case(name):
    from tests.golden.gen_golden import CASES
    _, base, nq, nk, tq, tk, cross, B, tc, seed, vt = [c for c in CASES if c[0] == name][0]
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk, v_transform=vt)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    inp = {k: torch.from_numpy(g[k]) for k in ("q", "k", "v", "extr_q", "extr_k", "coord_q", "coord_k")}
    if not cross:
        inp["extr_q"], inp["coord_q"] = inp["extr_k"], inp["coord_k"]
    return cfg, g, inp


@pytest.mark.parametrize("name", _golden_names(False))
def test_golden_vectors_fp32(name):
    """Committed outputs of the unmodified reference (fp32, CPU) vs the library fed the same fp32 inputs: the
    split-precision kernel (attn_fwd_hp_kernel) at the fp32 budget of 1e-3."""
    cfg, g, inp = _golden_case(name)
    out = _run(cfg, inp, tc=float(g["trans_coeff"]))
    _check(out, g["out"], float(g["trans_coeff"]), FP32_TOL, name + " [fp32, hp kernel]")


@pytest.mark.parametrize("name", _golden_names(False))
@pytest.mark.parametrize("flags", PIPELINES, ids=PIPELINE_IDS)
def test_golden_vectors_bf16_pipelines(name, flags):
    """The same golden cases with q/k/v rounded to bf16, so that the bf16 tensor-core kernels themselves (the default
    persistent pipeline and the older generations) run on the reference's vectors.  Two checks: against the oracle
    evaluated on the rounded inputs (kernel error alone), and against the reference's raw fp32 golden output (kernel
    error + input rounding) at the same bound."""
    cfg, g, inp = _golden_case(name)
    tc = float(g["trans_coeff"])
    inp = dict(inp, **{n: inp[n].to(torch.bfloat16) for n in "qkv"})
    out = _run(cfg, inp, tc=tc, flags=flags)
    _check(out, _oracle(cfg, inp, tc), tc, BF16_TOL, name + " [bf16 vs oracle on rounded inputs]")
    _check(out, g["out"], tc, BF16_TOL, name + " [bf16 vs raw fp32 golden]")


CASES_GPU = [
    # base, Nq, Nk, tq, tk, cross, B, dtype, tc
    (CFG1_A, 2, 2, 64, 64, False, 1, torch.bfloat16, 0.01),       # exactly one key tile
    (CFG1_B, 2, 2, 16, 16, False, 2, torch.float32, 0.01),        # a single partial tile
    (CFG1_B, 1, 1, 1, 1, False, 1, torch.bfloat16, 0.01),         # one token
    (CFG1_A, 2, 2, 1024, 1024, False, 2, torch.bfloat16, 0.01),   # BASELINE config 1 shape (T=2048)
    (MSN_SO3, 5, 5, 64, 64, False, 1, torch.bfloat16, 0.01),      # 3 tiles, tail 64
    (MSN_SO3, 5, 5, 256, 256, False, 2, torch.bfloat16, 0.01),    # MSN encoder shape
    (MSN_SO3, 5, 5, 512, 256, True, 1, torch.bfloat16, 0.01),     # MSN decoder shape
    (CLEVR, 2, 2, 300, 300, False, 2, torch.bfloat16, 0.01),      # CLEVR encoder shape (views straddle tiles)
    (CLEVR, 3, 2, 853, 300, True, 1, torch.bfloat16, 0.01),       # CLEVR decoder shape (Tq=2559)
    (CLEVR, 3, 2, 853, 300, True, 1, torch.float32, 1.0),         # fp32 I/O, trans_coeff 1
    (MSN_SO3, 1, 5, 1, 256, True, 2, torch.bfloat16, 0.01),       # render path: one query token
]


@pytest.mark.parametrize("case", CASES_GPU, ids=lambda c: f"D{c[0]['head_dim']}_{c[1]}x{c[3]}_{c[2]}x{c[4]}_{'x' if c[5] else 's'}_{str(c[7])[6:]}")
@pytest.mark.parametrize("flags", PIPELINES, ids=PIPELINE_IDS)
def test_fused_attention_matches_oracle(case, flags):
    base, nq, nk, tq, tk, cross, B, dtype, tc = case
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk)
    inp = make_inputs(cfg, B, tq, tk, cross=cross, seed=7, dtype=dtype)
    ref = _oracle(cfg, inp, tc)
    out = _run(cfg, inp, tc, flags)
    # fp32 inputs take the split-precision kernel whatever the pipeline flag says: fp32 budget
    _check(out, ref, tc, FP32_TOL if dtype == torch.float32 else BF16_TOL, "fused vs oracle")


@pytest.mark.parametrize("case", [
    (CFG1_B, 2, 2, 16, 16, False, 2, 0.01), (CFG1_A, 2, 2, 1024, 1024, False, 1, 0.01),
    (MSN_SO3, 5, 5, 256, 256, False, 2, 0.01), (MSN_SO3, 5, 5, 512, 256, True, 1, 1.0),
    (CLEVR, 2, 2, 300, 300, False, 2, 0.01), (CLEVR, 3, 2, 853, 300, True, 1, 1.0), (MSN_SO3, 1, 5, 1, 256, True, 2, 0.01),
], ids=["cfg1b", "cfg1a_T2048", "msn_enc", "msn_dec_tc1", "clevr_enc", "clevr_dec_tc1", "render_1q"])
def test_fp32_inputs_are_fp32_accurate(case):
    """fp32 q/k/v run the split-precision pipeline (bf16 hi + residual operands, fp32 accumulation): the north-star
    budget for fp32 is 1e-3 max-abs (relative to max(1, |ref|_max) when trans_coeff = 1 scales the output up)."""
    base, nq, nk, tq, tk, cross, B, tc = case
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk)
    inp = make_inputs(cfg, B, tq, tk, cross=cross, seed=17, dtype=torch.float32)
    ref = _oracle(cfg, inp, tc)
    out = _run(cfg, inp, tc)
    e_hp = _check(out, ref, tc, FP32_TOL, "fp32 split-precision")
    # peaked attention (large logits): one key dominates each row, so P/V rounding cannot average out
    inp2 = dict(inp)
    inp2["q"] = inp["q"] * 6.0
    ref2 = _oracle(cfg, inp2, tc)
    out2 = _run(cfg, inp2, tc)
    _check(out2, ref2, tc, FP32_TOL, "fp32 split-precision, peaked softmax")
    # the opt-out flag multiplies in plain bf16 (bf16 budget)
    from gta_b200 import _lib
    out3 = _run(cfg, inp, tc, flags=_lib.GTA_FLAG_FAST_FP32)
    e_fast = _check(out3, ref, tc, BF16_TOL, "fp32 inputs, GTA_FLAG_FAST_FP32")
    assert e_fast > e_hp


def test_contiguous_and_strided_inputs_agree():
    cfg = GtaConfig(**MSN_SO3, n_q_views=2, n_k_views=2)
    a = make_inputs(cfg, 2, 100, 100, cross=False, seed=3, dtype=torch.bfloat16, packed_layout=True)
    b = make_inputs(cfg, 2, 100, 100, cross=False, seed=3, dtype=torch.bfloat16, packed_layout=False)
    assert not a["q"].is_contiguous() and b["q"].is_contiguous()
    assert np.array_equal(_run(cfg, a), _run(cfg, b))


def test_v_transform_false_and_lse():
    ops = _ops()
    cfg = GtaConfig(**CLEVR, n_q_views=3, n_k_views=2, v_transform=False)
    inp = make_inputs(cfg, 1, 50, 70, cross=True, seed=4, dtype=torch.bfloat16)
    ref = _oracle(cfg, inp)
    out = _run(cfg, inp)
    _check(out, ref, what="v_transform=False")
    reps = _dev_reps(cfg, inp)
    o2, lse = ops.gta_attention_fwd(inp["q"].cuda(), inp["k"].cuda(), inp["v"].cuda(), reps, cfg.f_dims,
                                    trans_coeff=torch.tensor([0.01], device="cuda"), v_transform=False, return_lse=True)
    from oracle import torch_port as tp
    r = tp.build_reps(cfg, inp["extr_q"], inp["extr_k"], inp["coord_q"], inp["coord_k"])
    qt, kt, _ = tp.transform_qkv(cfg, inp["q"].float(), inp["k"].float(), inp["v"].float(), r, 0.01)
    lse_ref = torch.logsumexp(qt @ kt.transpose(-1, -2) * cfg.head_dim ** -0.5, -1)
    assert (lse.cpu() - lse_ref).abs().max() < 5e-2


def test_identity_pose_is_plain_attention_full_size():
    """Size-independent property at the MSN benchmark shape: identity extrinsics + zero coords => softmax(QK^T)V."""
    cfg = GtaConfig(**MSN_SO3, n_q_views=5, n_k_views=5)
    inp = make_inputs(cfg, 4, 256, 256, cross=False, seed=5, dtype=torch.bfloat16)
    inp["extr_q"] = inp["extr_k"] = torch.eye(4).expand(4, 5, 4, 4).contiguous()
    inp["coord_q"] = inp["coord_k"] = torch.zeros(4, 1280, 2)
    out = _run(cfg, inp)
    q, k, v = (inp[n].cuda().float() for n in "qkv")
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).cpu().numpy()
    assert np.abs(out - ref).max() < BF16_TOL


def test_frame_invariance_and_linearity_full_size():
    """E_i -> E_i G for one global SE(3) G leaves the output unchanged; the op is linear in V."""
    from gta_b200.synth import random_extrinsics
    cfg = GtaConfig(**MSN_SO3, n_q_views=5, n_k_views=5)
    inp = make_inputs(cfg, 2, 256, 256, cross=False, seed=6, dtype=torch.bfloat16)
    base = _run(cfg, inp, out_dtype=torch.float32)
    G = random_extrinsics(torch.Generator().manual_seed(1), 1, 1, False)[0, 0].double()
    inp2 = dict(inp)
    inp2["extr_q"] = inp2["extr_k"] = (inp["extr_k"].double() @ G).float()
    assert np.abs(_run(cfg, inp2, out_dtype=torch.float32) - base).max() < BF16_TOL
    inp3 = dict(inp)
    inp3["v"] = (inp["v"].float() * 2).to(torch.bfloat16)      # exact in bf16
    assert np.abs(_run(cfg, inp3, out_dtype=torch.float32) - 2 * base).max() < 1e-5


@pytest.mark.parametrize("flags", PIPELINES, ids=PIPELINE_IDS)
@pytest.mark.parametrize("case", [(CLEVR, 2, 2, 300, 300, False, 32), (MSN_SO3, 5, 5, 256, 256, False, 24),
                                  (CLEVR, 3, 2, 853, 300, True, 8)],
                         ids=["clevr_enc_B32", "msn_enc_B24", "clevr_dec_B8"])
def test_many_items_per_cta(case, flags):
    """Persistent kernels: several work items per CTA (ring / barrier phases across item boundaries, items without
    a second query tile) at the benchmark batch sizes."""
    base, nq, nk, tq, tk, cross, B = case
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk)
    inp = make_inputs(cfg, B, tq, tk, cross=cross, seed=12, dtype=torch.bfloat16)
    ref = _oracle(cfg, inp)
    out = _run(cfg, inp, flags=flags)
    _check(out, ref, what="many items per CTA")


def test_long_sequence_row_subset():
    """Sweep-style length (BASELINE config 4 family) where the reference cannot materialise [H,L,L]: the fused result
    on ALL rows is checked against the oracle evaluated on a subset of query rows with all keys."""
    from oracle import c_oracle
    cfg = GtaConfig(**MSN_SO3, n_q_views=2, n_k_views=2)
    tpv = 64 * 64
    inp = make_inputs(cfg, 1, tpv, tpv, cross=False, seed=9, dtype=torch.bfloat16)       # L = 8192, 64 key tiles
    out = _run(cfg, inp)
    rows = torch.cat([torch.arange(v * tpv + 1000, v * tpv + 1064) for v in range(2)])
    ref = c_oracle.gta_attention(cfg, inp["q"][:, :, rows].float(), inp["k"].float(), inp["v"].float(), inp["extr_k"],
                                 inp["extr_k"], inp["coord_k"][:, rows], inp["coord_k"], trans_coeff=0.01)
    _check(out[:, :, rows.numpy()], ref, what="L=8192 row subset")


def test_sweep_length_row_subset():
    """BASELINE config 4, first point of the sweep (B = 1, 2 views x 128x128 tokens, L = 32 768, d = 768): 256 key
    tiles per work item — K/V ring phases, the lazy-rescale threshold and the running sums over a long key axis — checked
    on query rows drawn from every part of the sequence against the oracle with ALL keys (the reference itself cannot
    materialise [H, L, L]).  Inputs are generated on the device like bench.py does."""
    from oracle import c_oracle
    ops = _ops()
    cfg = GtaConfig(**MSN_SO3, n_q_views=2, n_k_views=2)
    tpv = 128 * 128
    inp = make_inputs(cfg, 1, tpv, tpv, cross=False, seed=10, dtype=torch.bfloat16)
    out = _run(cfg, inp)
    assert np.isfinite(out).all()
    rows = torch.cat([torch.arange(s, s + 16) for s in (0, 5000, 16368, 16384, 24000, 32752)])   # 48 rows per view (the oracle maps row -> view by position), tile / view boundaries included
    ref = c_oracle.gta_attention(cfg, inp["q"][:, :, rows].float(), inp["k"].float(), inp["v"].float(), inp["extr_k"],
                                 inp["extr_k"], inp["coord_k"][:, rows], inp["coord_k"], trans_coeff=0.01)
    _check(out[:, :, rows.numpy()], ref, what="L=32768 row subset")
    # peaked rows: scaled queries make the running maximum jump by more than the lazy-rescale threshold (2^8) between key
    # tiles.  With logits of ~+-30 the bf16 rounding of the rotated OPERANDS alone moves a peaked softmax by several
    # 1e-2 (inherent to bf16 inputs of a tensor-core product, kernel-independent), so the kernel is checked against the
    # oracle evaluated on operands rounded exactly as the kernel rounds them (q', k', v' -> bf16), at the bf16 bound;
    # the distance to the unrounded oracle is printed for reference.
    from oracle import torch_port as tp
    inp2 = dict(inp)
    inp2["q"] = (inp["q"].float() * 8.0).to(torch.bfloat16)
    out2 = _run(cfg, inp2)[:, :, rows.numpy()]
    r = tp.build_reps(cfg, inp["extr_k"], inp["extr_k"], inp["coord_k"][:, rows], inp["coord_k"])
    qt, kt, vt = tp.transform_qkv(cfg, inp2["q"][:, :, rows].float(), inp["k"].float(), inp["v"].float(), r, 0.01)
    rb = lambda t: t.to(torch.bfloat16).double()
    o = torch.softmax(rb(qt) @ rb(kt).transpose(-1, -2) * cfg.head_dim ** -0.5, -1) @ rb(vt)
    T = lambda n: r[n].double().transpose(-1, -2)
    ref2 = tp._apply_blocks(o, cfg, tp._scale_translation(r["se3_qinv"].double(), 0.01), T("so3_d1_q"), T("so3_d2_q"),
                            r["so2_th_q"].double(), True, cfg.n_q_views).float().numpy()
    _check(out2, ref2, what="L=32768 row subset, peaked (|out| up to ~4), vs oracle on bf16-rounded operands", rel=True)
    exact = c_oracle.gta_attention(cfg, inp2["q"][:, :, rows].float(), inp["k"].float(), inp["v"].float(), inp["extr_k"],
                                   inp["extr_k"], inp["coord_k"][:, rows], inp["coord_k"], trans_coeff=0.01)
    print("  (vs the unrounded oracle: %.3e, |ref|max %.2f)" % (float(np.abs(out2 - exact).max()), float(np.abs(exact).max())))


def test_msn_headline_batch_subset():
    """The headline benchmark configuration itself (MSN gta_so3 encoder, B = 64: 2560 work items, 17-18 per CTA): the
    oracle is evaluated on a subset of the batch elements — first, last and the ones whose items straddle CTA rounds —
    and compared with the corresponding slices of the full-batch result."""
    cfg = GtaConfig(**MSN_SO3, n_q_views=5, n_k_views=5)
    inp = make_inputs(cfg, 64, 256, 256, cross=False, seed=13, dtype=torch.bfloat16)
    out = _run(cfg, inp)
    assert np.isfinite(out).all()
    for b in (0, 3, 18, 37, 63):
        sub = {k: (v[b:b + 1] if torch.is_tensor(v) else v) for k, v in inp.items()}
        sub["extr_q"], sub["coord_q"] = sub["extr_k"], sub["coord_k"]
        _check(out[b:b + 1], _oracle(cfg, sub), what="MSN B=64, batch element %d" % b)


def test_dropin_signature_with_reference_format_reps():
    """multihead_geometric_transform_attention(q,k,v,attn_fn,f_dims,reps,...) fed rep tensors laid out as the
    reference's pre_compute_reps leaves them in `extras`."""
    from gta_b200 import gta as fast
    from oracle import torch_port as tp
    cfg = GtaConfig(**MSN_SO3, n_q_views=3, n_k_views=2)
    inp = make_inputs(cfg, 2, 40, 64, cross=True, seed=8, dtype=torch.bfloat16)
    r = tp.build_reps(cfg, inp["extr_q"], inp["extr_k"], inp["coord_q"], inp["coord_k"])
    extras = {
        "se3rep_q": torch.linalg.inv(inp["extr_q"]).cuda(), "se3rep_k": r["se3_k"].cuda(),
        "inv_se3rep_q": r["se3_qinv"].cuda(),
        "so3rep_q": [r["so3_d1_q"].cuda(), r["so3_d2_q"].cuda()], "so3rep_k": [r["so3_d1_k"].cuda(), r["so3_d2_k"].cuda()],
        "so2rep_q": tp.so2_mats(r["so2_th_q"]).cuda(), "so2rep_k": tp.so2_mats(r["so2_th_k"]).cuda(),
    }

    class AttnFn:
        scale = cfg.head_dim ** -0.5
    tc = torch.nn.Parameter(torch.tensor([0.01], device="cuda"))
    with torch.no_grad():
        out, attn = fast.multihead_geometric_transform_attention(
            inp["q"].cuda(), inp["k"].cuda(), inp["v"].cuda(), attn_fn=AttnFn(), f_dims=cfg.f_dims, reps=extras,
            trans_coeff=tc, v_transform=True, euclid=False)
    assert attn is None and out.shape == inp["q"].shape
    ref = _oracle(cfg, inp)
    _check(out.float().cpu().numpy(), ref, what="drop-in signature")
    assert "_gta_b200_packed" in extras      # packed once, reused by the next layer
    # euclid_sim through the same signature (se3 = 48 = 16 homogenised 3-vectors; needs extras['se3rep_q'])
    cfg_e = GtaConfig(**MSN_SO3_EUCLID, n_q_views=3, n_k_views=2)
    with torch.no_grad():
        out_e, _ = fast.multihead_geometric_transform_attention(
            inp["q"].cuda(), inp["k"].cuda(), inp["v"].cuda(), AttnFn(), cfg_e.f_dims, extras, trans_coeff=tc, euclid=True)
    ref_e = _oracle(cfg_e, inp)
    _check(out_e.float().cpu().numpy(), ref_e, what="drop-in signature, euclid_sim")


def test_dropin_t2_reference_format_reps():
    """t2 block fed the reference-format [B,T,3,3] matrices (make_T2mats / torch.linalg.inv, encoder.py:208-215)."""
    from gta_b200 import gta as fast
    cfg = GtaConfig(**CLEVR_T2, n_q_views=3, n_k_views=2)
    inp = make_inputs(cfg, 2, 40, 64, cross=True, seed=18, dtype=torch.bfloat16)
    t2q, t2k = fast.make_T2mats(inp["coord_q"].cuda()), fast.make_T2mats(inp["coord_k"].cuda())
    from oracle import torch_port as tp
    assert (t2q.cpu() - tp.t2_mats(inp["coord_q"])).abs().max() == 0
    extras = {"se3rep_q": torch.linalg.inv(inp["extr_q"]).cuda(), "se3rep_k": torch.linalg.inv(inp["extr_k"]).cuda(),
              "inv_se3rep_q": inp["extr_q"].cuda(), "t2rep_q": t2q, "t2rep_k": t2k, "inv_t2rep_q": torch.linalg.inv(t2q)}

    class AttnFn:
        scale = cfg.head_dim ** -0.5
    with torch.no_grad():
        out, _ = fast.multihead_geometric_transform_attention(
            inp["q"].cuda(), inp["k"].cuda(), inp["v"].cuda(), AttnFn(), cfg.f_dims, extras,
            trans_coeff=torch.tensor([0.01], device="cuda"))
    ref = _oracle(cfg, inp)
    _check(out.float().cpu().numpy(), ref, what="drop-in t2")


@pytest.mark.parametrize("case", [
    (CLEVR_T2, 2, 2, 300, 300, False, 2, torch.bfloat16, 0.01), (CLEVR_EUCLID, 3, 2, 853, 300, True, 1, torch.bfloat16, 0.01),
    (MSN_T2, 5, 5, 256, 256, False, 1, torch.bfloat16, 0.01), (MSN_SO3_EUCLID, 5, 5, 256, 256, False, 1, torch.bfloat16, 0.01),
    (MSN_SO3_EUCLID, 5, 5, 128, 256, True, 1, torch.float32, 0.3), (CLEVR_EUCLID, 2, 2, 300, 300, False, 1, torch.float32, 1.0),
], ids=["clevr_t2_enc", "clevr_euclid_dec", "msn_t2_enc", "msn_so3_euclid_enc", "msn_so3_euclid_f32", "clevr_euclid_f32_tc1"])
def test_ablation_blocks_at_model_shapes(case):
    base, nq, nk, tq, tk, cross, B, dtype, tc = case
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk)
    inp = make_inputs(cfg, B, tq, tk, cross=cross, seed=19, dtype=dtype)
    ref = _oracle(cfg, inp, tc)
    out = _run(cfg, inp, tc)
    _check(out, ref, tc, BF16_TOL, "ablation blocks")


def test_generic_rotated_operands_match_oracle():
    from oracle import torch_port as tp
    ops = _ops()
    for base in (CLEVR_T2, MSN_SO3_EUCLID, CLEVR_EUCLID):
        cfg = GtaConfig(**base, n_q_views=3, n_k_views=2)
        inp = make_inputs(cfg, 2, 8, 16, cross=True, seed=2, dtype=torch.float32)
        reps = _dev_reps(cfg, inp)
        qt, kt, vt = ops.rotate_debug(inp["q"].cuda(), inp["k"].cuda(), inp["v"].cuda(), reps, cfg.f_dims,
                                      trans_coeff=torch.tensor([0.3], device="cuda"), euclid=cfg.euclid)
        r = tp.build_reps(cfg, inp["extr_q"], inp["extr_k"], inp["coord_q"], inp["coord_k"])
        q2, k2, v2 = tp.transform_qkv(cfg, inp["q"], inp["k"], inp["v"], r, 0.3)
        for a, b in ((qt, q2), (kt, k2), (vt, v2)):
            assert (a.cpu() - b).abs().max() < 5e-6


# ------------------------------------------------------------------------------------------------ backward (f1)
def _grad_oracle(cfg, inp, tc, dout):
    """Gradients of the torch restatement (fp64 autograd) — pinned to the reference's autograd by
    tests/test_oracle.py::test_grad_oracle_matches_reference_autograd."""
    from oracle import torch_port as tp
    q, k, v = (inp[n].double().clone().requires_grad_(True) for n in "qkv")
    tcv = torch.tensor([tc], dtype=torch.float64, requires_grad=True)
    d = lambda t: t.double()
    reps = tp.build_reps(cfg, d(inp["extr_q"]), d(inp["extr_k"]), d(inp["coord_q"]), d(inp["coord_k"]))
    qt, kt, vt = tp.transform_qkv(cfg, q, k, v, reps, tcv)
    for t in (qt, kt, vt):
        if t.requires_grad and not t.is_leaf:
            t.retain_grad()
    # the tail of tp.gta_attention on these very tensors (so that qt/kt/vt.grad are populated)
    sim = qt @ kt.transpose(-1, -2)
    if cfg.euclid:                          # EuclidAttnFn, layers.py:219-223
        sim = sim - 0.5 * qt.pow(2).sum(-1)[..., None] - 0.5 * kt.pow(2).sum(-1)[..., None, :]
    out = torch.softmax(sim * cfg.head_dim ** -0.5, -1) @ vt
    if cfg.v_transform:
        has = lambda n: reps[n].transpose(-1, -2) if n in reps else None
        Ao = tp._scale_translation(reps["se3_qinv"], tcv) if cfg.dims()[1] else None
        out = tp._apply_blocks(out, cfg, Ao, has("so3_d1_q"), has("so3_d2_q"), reps.get("so2_th_q"), True, cfg.n_q_views,
                               reps.get("t2_qinv"))
    out.backward(dout.double())
    # d(trans_coeff) is a sum of four strongly cancelling parts (query, key, value and output side, e.g. 5.8 + 36.0 - 16.6
    # - 22.5 = 2.7): its error budget is relative to the sum of their magnitudes
    triv, se3, _, _ = cfg.dims()
    tc_scale = 0.0
    if se3:
        B, H = q.shape[:2]
        n = 3 if cfg.euclid else 4
        v4 = lambda x, N: x.detach()[..., triv:triv + se3].reshape(B, H, N, -1, se3 // n, n)
        col = lambda g, M: g[..., 0] * M[..., 0, 3] + g[..., 1] * M[..., 1, 3] + g[..., 2] * M[..., 2, 3]
        Eq, Ek = reps["se3_qinv"][:, None, :, None, None], reps["se3_k"][:, None, :, None, None]
        Nq, Nk = cfg.n_q_views, cfg.n_k_views
        if cfg.euclid:      # y = A x + t tc for every side (homogenised 3-vectors): d/dtc = gradient . translation column
            Cq = reps["se3_q"][:, None, :, None, None]
            parts = [col(v4(qt.grad, Nq), Cq).sum(), col(v4(kt.grad, Nk), Ek).sum()]
            if cfg.v_transform:
                parts += [col(v4(vt.grad, Nk), Ek).sum(), col(v4(dout.double(), Nq), Eq).sum()]
        else:
            Q4 = v4(q, Nq)
            parts = [(v4(qt.grad, Nq)[..., 3] * (Eq[..., 0, 3] * Q4[..., 0] + Eq[..., 1, 3] * Q4[..., 1] + Eq[..., 2, 3] * Q4[..., 2])).sum(),
                     (col(v4(kt.grad, Nk), Ek) * v4(k, Nk)[..., 3]).sum()]
            if cfg.v_transform:
                parts += [(col(v4(vt.grad, Nk), Ek) * v4(v, Nk)[..., 3]).sum(),
                          (col(v4(dout.double(), Nq), Eq) * v4(out, Nq)[..., 3] / Eq[..., 3, 3]).sum()]
        assert abs(float(sum(parts)) - float(tcv.grad)) < 1e-6 * max(1.0, float(sum(p.abs() for p in parts)))
        tc_scale = float(sum(p.abs() for p in parts))
    g = lambda t: None if t is None else t.float().numpy()
    return out.detach().float().numpy(), g(q.grad), g(k.grad), g(v.grad), (g(tcv.grad), tc_scale)


@pytest.mark.parametrize("case", [
    (CFG1_B, 2, 2, 16, 16, False, 2, torch.bfloat16, 0.3, True),        # one partial tile, D = 32
    (MSN_SO3, 3, 2, 40, 64, True, 2, torch.bfloat16, 0.3, True),        # cross, ragged tiles, D = 96
    (MSN_SO3, 5, 5, 256, 256, False, 1, torch.bfloat16, 0.01, True),    # MSN encoder shape
    (CLEVR, 3, 2, 171, 300, True, 1, torch.bfloat16, 0.5, True),        # D = 64, views straddle tiles, Tq = 513
    (CLEVR, 2, 2, 150, 150, False, 1, torch.bfloat16, 0.5, False),      # v_transform = False
    (MSN_SO3, 3, 2, 40, 64, True, 1, torch.float32, 0.3, True),         # fp32 I/O (bf16 tensor-core math in the backward)
    (CLEVR, 3, 2, 853, 300, True, 1, torch.bfloat16, 0.01, True),       # CLEVR decoder shape (Tq = 2559, ragged everywhere)
    (MSN_SO3, 5, 5, 512, 256, True, 1, torch.bfloat16, 0.01, True),     # MSN decoder shape (Tq = 2560, Tk = 1280)
    (CLEVR_T2, 3, 2, 171, 300, True, 1, torch.bfloat16, 0.3, True),     # generic path: triv 2 | se3 32 | t2 30 (runs/clevrtr/GTA/gta_t2)
    (MSN_T2, 2, 2, 100, 100, False, 2, torch.bfloat16, 0.3, True),      # generic path: se3 48 | t2 48 (runs/msn/GTA/gta_t2)
    (CLEVR_T2, 2, 2, 150, 150, False, 1, torch.float32, 0.5, False),    # generic path, fp32 I/O, v_transform = False
    (CLEVR_EUCLID, 3, 2, 171, 300, True, 1, torch.bfloat16, 0.3, True), # euclid_sim: triv 2 | se3 30 | so2 32 (runs/clevrtr/GTA/gta_euclid), padded head dim 96
    (MSN_SO3_EUCLID, 2, 2, 100, 100, False, 2, torch.bfloat16, 0.3, True),  # euclid_sim + so3, padded head dim 128 (kernel pair)
    (CLEVR_EUCLID, 2, 2, 150, 150, False, 1, torch.bfloat16, 0.5, False),   # euclid_sim, v_transform = False
], ids=["cfg1b", "msn_cross", "msn_enc", "clevr_dec", "clevr_novt", "msn_cross_f32", "clevr_dec_full", "msn_dec_full",
        "clevr_t2_cross", "msn_t2", "clevr_t2_f32_novt", "clevr_euclid_cross", "msn_so3_euclid", "clevr_euclid_novt"])
def test_fused_backward_matches_autograd_oracle(case):
    """dq, dk, dv and d(trans_coeff) of the fused backward, through the public drop-in under autograd, against fp64
    autograd of the oracle on the same (rounded) inputs.  bf16 tensor-core math: errors relative to the gradient scale."""
    from gta_b200 import gta as fast
    from oracle import torch_port as tp
    base, nq, nk, tq, tk, cross, B, dtype, tc, vt = case
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk, v_transform=vt)
    inp = make_inputs(cfg, B, tq, tk, cross=cross, seed=41, dtype=dtype)
    gen = torch.Generator().manual_seed(7)
    dout = torch.randn(inp["q"].shape, generator=gen).to(dtype)
    ref_out, rq, rk, rv, rtc = _grad_oracle(cfg, inp, tc, dout.float())

    r = tp.build_reps(cfg, inp["extr_q"], inp["extr_k"], inp["coord_q"], inp["coord_k"])
    extras = {"se3rep_q": torch.linalg.inv(inp["extr_q"]).cuda(), "se3rep_k": r["se3_k"].cuda(),
              "inv_se3rep_q": r["se3_qinv"].cuda()}
    if cfg.so3:
        extras.update({"so3rep_q": [r["so3_d1_q"].cuda(), r["so3_d2_q"].cuda()],
                       "so3rep_k": [r["so3_d1_k"].cuda(), r["so3_d2_k"].cuda()]})
    if cfg.so2:
        extras.update({"so2rep_q": tp.so2_mats(r["so2_th_q"]).cuda(), "so2rep_k": tp.so2_mats(r["so2_th_k"]).cuda()})
    if cfg.t2_dim():        # reference-format [B,T,3,3] matrices (make_T2mats, encoder.py:208-215)
        extras.update({"t2rep_q": r["t2_q"].cuda(), "t2rep_k": r["t2_k"].cuda(), "inv_t2rep_q": r["t2_qinv"].cuda()})

    class AttnFn:
        scale = cfg.head_dim ** -0.5
    q, k, v = (inp[n].cuda().clone().requires_grad_(True) for n in "qkv")
    tcp = torch.nn.Parameter(torch.tensor([tc], device="cuda"))
    out, _ = fast.multihead_geometric_transform_attention(q, k, v, AttnFn(), cfg.f_dims, extras, trans_coeff=tcp,
                                                          v_transform=vt, euclid=cfg.euclid)
    assert out.requires_grad
    out.backward(dout.cuda())
    torch.cuda.synchronize()
    # (euclid_sim in bf16 math: logits carry -|k'|^2/2 terms of tens of units, bound relative as in the forward tests)
    _check(out.detach().float().cpu().numpy(), ref_out, tc, BF16_TOL, "forward under autograd", rel=cfg.euclid)
    for name, got, ref in (("dq", q.grad, rq), ("dk", k.grad, rk), ("dv", v.grad, rv)):
        got = got.float().cpu().numpy()
        assert np.isfinite(got).all(), name
        err, scale_ = np.abs(got - ref).max(), np.abs(ref).max()
        assert err < 2e-2 * scale_ + 1e-3, (name, err, scale_)
    got_tc = float(tcp.grad.float().cpu())
    rtc, tc_scale = rtc
    assert abs(got_tc - float(rtc[0])) < 1e-2 * max(1.0, tc_scale), (got_tc, float(rtc[0]), tc_scale)


def test_backward_abi_direct_and_linearity():
    """ops.gta_attention_bwd called directly: gradients are linear in dout, and dtrans_coeff is accumulated per call."""
    ops = _ops()
    cfg = GtaConfig(**MSN_SO3, n_q_views=2, n_k_views=2)
    inp = make_inputs(cfg, 1, 100, 100, cross=False, seed=43, dtype=torch.bfloat16)
    reps = _dev_reps(cfg, inp)
    q, k, v = (inp[n].cuda() for n in "qkv")
    tc = torch.tensor([0.2], device="cuda")
    out, lse = ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, return_lse=True)
    dout = torch.randn(out.shape, device="cuda").bfloat16()
    g1 = ops.gta_attention_bwd(dout, q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc)
    g2 = ops.gta_attention_bwd(dout * 2, q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc)
    for a, b in zip(g1[:3], g2[:3]):
        assert a.shape == q.shape
        assert (2 * a.float() - b.float()).abs().max() <= 2e-2 * b.float().abs().max()
    assert abs(2 * float(g1[3]) - float(g2[3])) <= 2e-2 * abs(float(g2[3])) + 1e-4


@pytest.mark.parametrize("case", [
    (CFG1_B, 2, 2, 16, 16, False, 2), (MSN_SO3, 3, 2, 40, 64, True, 2), (MSN_SO3, 5, 5, 256, 256, False, 1),
    (CLEVR, 3, 2, 171, 300, True, 1), (CLEVR, 3, 2, 853, 300, True, 1), (MSN_SO3, 2, 5, 256, 256, True, 1),
    (CLEVR_T2, 3, 2, 171, 300, True, 1), (CLEVR_EUCLID, 2, 2, 150, 150, False, 1),     # generic path around both cores
], ids=["cfg1b", "msn_cross", "msn_enc", "clevr_dec", "clevr_dec_full", "msn_more_keys", "clevr_t2", "clevr_euclid"])
def test_backward_fused_kernel_matches_kernel_pair(case):
    """The fused backward kernel (dK, dV and bulk-reduced dQ partial sums in one launch; default for head dims <= 96) against
    the dK/dV + dQ kernel pair (GTA_FLAG_BWD_SPLIT) on the same call: same bf16 operands, different summation order."""
    from gta_b200 import _lib
    ops = _ops()
    base, nq, nk, tq, tk, cross, B = case
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk)
    inp = make_inputs(cfg, B, tq, tk, cross=cross, seed=47, dtype=torch.bfloat16)
    reps = _dev_reps(cfg, inp)
    q, k, v = (inp[n].cuda() for n in "qkv")
    tc = torch.tensor([0.3], device="cuda")
    eu = dict(euclid=cfg.euclid)
    out, lse = ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, return_lse=True, **eu)
    dout = torch.randn(out.shape, device="cuda").bfloat16()
    force = _lib.GTA_FLAG_SINGLE_LAUNCH      # (calls of less than one wave of key tiles default to the kernel pair)
    fused = ops.gta_attention_bwd(dout, q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc, flags=force, **eu)
    again = ops.gta_attention_bwd(dout, q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc, flags=force, **eu)
    pair = ops.gta_attention_bwd(dout, q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc, flags=_lib.GTA_FLAG_BWD_SPLIT, **eu)
    # the run-time-layout epilogue (head layouts without a compile-time instantiation) on the same call
    rt = ops.gta_attention_bwd(dout, q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc, flags=force | _lib.GTA_FLAG_RUNTIME_LAYOUT, **eu)
    torch.cuda.synchronize()
    for name, a, b in zip(("dq", "dk", "dv"), fused[:3], rt[:3]):
        if name == "dq":
            assert float((a.float() - b.float()).abs().max()) <= 1e-2 * float(b.float().abs().max()) + 1e-4
        else:
            assert torch.equal(a, b), name + " (run-time layout epilogue)"
    if fused[3] is not None:      # (the compile-time epilogue takes the raw w components from the bf16 K'/V' images)
        assert abs(float(fused[3]) - float(rt[3])) <= 1e-2 * max(1.0, abs(float(rt[3])))
    for name, a, a2, b in zip(("dq", "dk", "dv"), fused[:3], again[:3], pair[:3]):
        assert torch.isfinite(a.float()).all(), name
        scale_ = float(b.float().abs().max())
        err = float((a.float() - b.float()).abs().max())
        print(f"{name}: fused vs pair max-abs {err:.3e} (scale {scale_:.3e})")
        assert err <= 1e-2 * scale_ + 1e-4, (name, err, scale_)
        if name != "dq":                # dK / dV are accumulated in a fixed order; the dQ partial sums are added in arrival order
            assert torch.equal(a, a2), name
    if fused[3] is not None:
        assert abs(float(fused[3]) - float(pair[3])) <= 1e-2 * max(1.0, abs(float(pair[3])))


def test_backward_headline_batch_and_long_sequence():
    """The backward at benchmark sizes through size-independent properties: (i) MSN encoder shape at B = 64 (5 120 CTAs of the
    fused kernel): the gradients of a batch element do not depend on the rest of the batch — three elements are recomputed at
    B = 1 (dK / dV bit-equal, dQ up to the order of its fp32 partial sums); (ii) L = 8 192 (64 query tiles per key tile: ring
    slots and barrier parities over many pairs): fused kernel against the kernel pair."""
    from gta_b200 import _lib
    ops = _ops()
    cfg = GtaConfig(**MSN_SO3, n_q_views=5, n_k_views=5)
    B = 64
    inp = make_inputs(cfg, B, 256, 256, cross=False, seed=71, dtype=torch.bfloat16)
    reps = _dev_reps(cfg, inp)
    q, k, v = (inp[n].cuda() for n in "qkv")
    tc = torch.tensor([0.01], device="cuda")
    out, lse = ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, return_lse=True)
    dout = torch.randn(out.shape, device="cuda").bfloat16()
    dq, dk, dv, _ = ops.gta_attention_bwd(dout, q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc)
    assert all(torch.isfinite(t.float()).all() for t in (dq, dk, dv))
    for b in (0, 31, 63):
        one = {n: (t[b:b + 1] if torch.is_tensor(t) and t.shape[0] == B else t) for n, t in inp.items()}
        one["extr_q"], one["coord_q"] = one["extr_k"], one["coord_k"]
        reps1 = _dev_reps(cfg, one)
        q1, k1, v1 = (x[b:b + 1] for x in (q, k, v))
        o1, l1 = ops.gta_attention_fwd(q1, k1, v1, reps1, cfg.f_dims, trans_coeff=tc, return_lse=True)
        g1 = ops.gta_attention_bwd(dout[b:b + 1], q1, k1, v1, o1, l1, reps1, cfg.f_dims, trans_coeff=tc, flags=_lib.GTA_FLAG_SINGLE_LAUNCH)
        assert torch.equal(o1, out[b:b + 1])
        assert torch.equal(g1[1], dk[b:b + 1]) and torch.equal(g1[2], dv[b:b + 1]), b
        err = float((g1[0].float() - dq[b:b + 1].float()).abs().max())
        assert err <= 2 ** -7 * float(dq[b:b + 1].float().abs().max()), (b, err)        # one bf16 ulp of the largest entry
    del q, k, v, out, lse, dout, dq, dk, dv
    cfg = GtaConfig(**MSN_SO3, n_q_views=2, n_k_views=2)
    inp = make_inputs(cfg, 1, 4096, 4096, cross=False, seed=72, dtype=torch.bfloat16)
    reps = _dev_reps(cfg, inp)
    q, k, v = (inp[n].cuda() for n in "qkv")
    out, lse = ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, return_lse=True)
    dout = torch.randn(out.shape, device="cuda").bfloat16()
    fused = ops.gta_attention_bwd(dout, q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc)
    pair = ops.gta_attention_bwd(dout, q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc, flags=_lib.GTA_FLAG_BWD_SPLIT)
    for name, a, b_ in zip(("dq", "dk", "dv"), fused[:3], pair[:3]):
        scale_ = float(b_.float().abs().max())
        err = float((a.float() - b_.float()).abs().max())
        print(f"L=8192 {name}: fused vs pair max-abs {err:.3e} (scale {scale_:.3e})")
        assert err <= 1e-2 * scale_ + 1e-4, (name, err, scale_)
    assert abs(float(fused[3]) - float(pair[3])) <= 1e-2 * max(1.0, abs(float(pair[3])))


def test_host_pipeline_matches_device_call():
    """gta_b200.host.HostStagedAttention (pinned host buffers, batch chunks over three streams) is bit-identical to the
    device-resident call on the same inputs, for self- and cross-attention."""
    from gta_b200.host import HostStagedAttention
    for cross in (False, True):
        cfg = GtaConfig(**MSN_SO3, n_q_views=3 if cross else 2, n_k_views=2)
        B, tq, tk = 6, (40 if cross else 64), 64
        inp = make_inputs(cfg, B, tq, tk, cross=cross, seed=51, dtype=torch.bfloat16)
        base = lambda t: t if t._base is None else base(t._base)
        bufs = ({"q": base(inp["q"]).pin_memory(), "kv": base(inp["k"]).pin_memory()} if cross
                else {"qkv": base(inp["q"]).pin_memory()})
        small = {n: inp[n].contiguous().pin_memory() for n in (("extr_k", "coord_k", "extr_q", "coord_q") if cross
                                                                  else ("extr_k", "coord_k"))}
        out_host = torch.empty(B, cfg.n_q_views * tq, cfg.heads, cfg.head_dim, dtype=torch.bfloat16).pin_memory()
        pipe = HostStagedAttention(cfg, bufs, small, out_host, torch.device("cuda", 0), chunks=4)
        pipe.run()
        torch.cuda.synchronize()
        ref = _run(cfg, inp)                                  # [B,H,Tq,D]
        assert np.array_equal(out_host.permute(0, 2, 1, 3).float().numpy(), ref)
        # back-to-back calls overlap (the input copies of a call do not wait for the previous call to drain): same result
        out_host.zero_()
        for _ in range(3):
            pipe.run()
        pipe.run(inputs_on_stream=True)
        torch.cuda.synchronize()
        assert np.array_equal(out_host.permute(0, 2, 1, 3).float().numpy(), ref)


@pytest.mark.parametrize("base", [MSN_SO3, CLEVR_T2], ids=["msn_so3", "clevr_t2"])
def test_attention_map_output(base):
    """The reference's second return value (`attn`, source/layers.py:207-211) through gta_attn_probs and
    through the drop-in with RETURN_ATTENTION_MAP."""
    from gta_b200 import gta as fast
    from oracle import torch_port as tp
    ops = _ops()
    cfg = GtaConfig(**base, n_q_views=3, n_k_views=2)
    inp = make_inputs(cfg, 2, 30, 45, cross=True, seed=61, dtype=torch.bfloat16)
    reps = _dev_reps(cfg, inp)
    q, k, v = (inp[n].cuda() for n in "qkv")
    tc = torch.tensor([0.3], device="cuda")
    out, lse = ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, return_lse=True)
    attn = ops.gta_attention_probs(q, k, lse, reps, cfg.f_dims, trans_coeff=tc).cpu()
    r = tp.build_reps(cfg, inp["extr_q"], inp["extr_k"], inp["coord_q"], inp["coord_k"])
    qt, kt, vt = tp.transform_qkv(cfg, inp["q"].float(), inp["k"].float(), inp["v"].float(), r, 0.3)
    ref = torch.softmax(qt @ kt.transpose(-1, -2) * cfg.head_dim ** -0.5, -1)
    assert attn.shape == ref.shape
    assert (attn - ref).abs().max() < 1e-2
    assert (attn.sum(-1) - 1).abs().max() < 2e-2
    # attn @ v' reproduces the forward output (before the output rep): consistency of the map with the fused kernel
    if not cfg.t2_dim():
        extras = {"se3rep_q": torch.linalg.inv(inp["extr_q"]).cuda(), "se3rep_k": r["se3_k"].cuda(), "inv_se3rep_q": r["se3_qinv"].cuda(),
                  "so3rep_q": [r["so3_d1_q"].cuda(), r["so3_d2_q"].cuda()], "so3rep_k": [r["so3_d1_k"].cuda(), r["so3_d2_k"].cuda()],
                  "so2rep_q": tp.so2_mats(r["so2_th_q"]).cuda(), "so2rep_k": tp.so2_mats(r["so2_th_k"]).cuda()}

        class AttnFn:
            scale = cfg.head_dim ** -0.5
        fast.RETURN_ATTENTION_MAP = True
        try:
            with torch.no_grad():
                o2, a2 = fast.multihead_geometric_transform_attention(q, k, v, AttnFn(), cfg.f_dims, extras, trans_coeff=tc)
        finally:
            fast.RETURN_ATTENTION_MAP = False
        assert a2 is not None and (a2.cpu() - ref).abs().max() < 1e-2
        # (the extras-based reps were built by torch on the host, `out` with the device rep builder: equal up to rounding)
        assert (o2.float() - out.float()).abs().max() < 1e-2
