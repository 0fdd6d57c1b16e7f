"""Small forward (+ backward) calls of every pipeline for compute-sanitizer (memcheck / racecheck / synccheck):
BASELINE config 1 shape at B = 1 and the ragged CLEVR encoder / decoder shapes at B = 1.

    compute-sanitizer --tool racecheck python tools/sanitizer_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gta_b200 import _lib, ops  # noqa: E402
from gta_b200.synth import CFG1_A, CLEVR, CLEVR_EUCLID, CLEVR_T2, MSN_SO3, MSN_SO3_EUCLID, GtaConfig, make_inputs  # noqa: E402

CASES = [("cfg1", CFG1_A, 2, 2, 256, 256, False), ("clevr_enc", CLEVR, 2, 2, 300, 300, False),
         ("clevr_dec", CLEVR, 3, 2, 171, 300, True), ("msn_small", MSN_SO3, 5, 5, 64, 64, False)]
PIPES = [("two_launch", _lib.GTA_FLAG_TWO_LAUNCH), ("single_launch", _lib.GTA_FLAG_SINGLE_LAUNCH),
         ("v4_streaming", _lib.GTA_FLAG_V4_PIPELINE), ("v5_spare_p", _lib.GTA_FLAG_V5_PIPELINE)]
only = sys.argv[1:] or None
for name, base, nq, nk, tq, tk, cross in CASES:
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk)
    inp = make_inputs(cfg, 1, tq, tk, cross=cross, seed=3, dtype=torch.bfloat16)
    ek, ck = inp["extr_k"].cuda(), inp["coord_k"].cuda()
    eq = inp["extr_q"].cuda() if cross else ek
    cq = inp["coord_q"].cuda() if cross else ck
    reps = ops.build_reps(eq, ek, cq, ck, so2_nfreqs=cfg.so2, so3_maxdeg=cfg.so3)
    q, k, v = (inp[n].cuda() for n in "qkv")
    tc = torch.tensor([0.01], device="cuda")
    ref = None
    for pname, fl in PIPES:
        if only and pname not in only:
            continue
        out, lse = ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, flags=fl, return_lse=True)
        torch.cuda.synchronize()
        ref = out if ref is None else ref
        print("%-10s %-14s max |out - first pipeline| = %.3e" % (name, pname, float((out.float() - ref.float()).abs().max())), flush=True)
    if not only or "backward" in only:
        out, lse = ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, return_lse=True)
        dout = torch.randn_like(out)
        for bname, bfl in (("fused kernel", _lib.GTA_FLAG_SINGLE_LAUNCH), ("fused kernel, run-time layout", _lib.GTA_FLAG_SINGLE_LAUNCH | _lib.GTA_FLAG_RUNTIME_LAYOUT), ("kernel pair", _lib.GTA_FLAG_BWD_SPLIT)):
            g = ops.gta_attention_bwd(dout, q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc, flags=bfl)
            torch.cuda.synchronize()
            print("%-10s backward ok (%s), |dq|max %.3f" % (name, bname, float(g[0].float().abs().max())), flush=True)

# generic-path layouts (t2 block, euclid_sim): forward + backward through the element-wise rep passes
if not only or "generic" in only:
    for name, base, nq, nk, tq, tk, cross in [("clevr_t2", CLEVR_T2, 3, 2, 57, 100, True), ("clevr_euclid", CLEVR_EUCLID, 2, 2, 75, 75, False),
                                              ("msn_so3_euclid", MSN_SO3_EUCLID, 2, 2, 40, 40, False)]:
        cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk)
        inp = make_inputs(cfg, 1, tq, tk, cross=cross, seed=5, dtype=torch.bfloat16)
        ek, ck = inp["extr_k"].cuda(), inp["coord_k"].cuda()
        eq = inp["extr_q"].cuda() if cross else ek
        cq = inp["coord_q"].cuda() if cross else ck
        reps = ops.build_reps(eq, ek, cq, ck, so2_nfreqs=cfg.so2, so3_maxdeg=cfg.so3, t2=bool(cfg.t2_dim()), euclid=cfg.euclid)
        q, k, v = (inp[n].cuda() for n in "qkv")
        tc = torch.tensor([0.3], device="cuda")
        out, lse = ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, return_lse=True, euclid=cfg.euclid)
        g = ops.gta_attention_bwd(torch.randn_like(out), q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc, euclid=cfg.euclid)
        torch.cuda.synchronize()
        print("%-14s generic forward + backward ok, |dq|max %.3f d(tc) %.4f" % (name, float(g[0].float().abs().max()), float(g[3])), flush=True)
