"""Per-CTA phase timing of the attention kernel from its optional clock64 stamps (GtaAttnParams.debug_clocks).
Run on the GPU box: python tools/phase_timing.py [workload] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from gta_b200 import _lib, ops  # noqa: E402
from gta_b200.synth import GtaConfig, make_inputs  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "msn_enc"
    base, nq, nk, tq, tk, cross, B, _ = WORKLOADS[name]
    if len(sys.argv) > 2:
        B = int(sys.argv[2])
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk)
    inp = make_inputs(cfg, B, tq, tk, cross=cross, seed=0, dtype=torch.bfloat16)
    dev = torch.device("cuda")
    ek, ck = inp["extr_k"].to(dev), inp["coord_k"].to(dev)
    eq = inp["extr_q"].to(dev) if cross else ek
    cq = inp["coord_q"].to(dev) if cross else ck
    reps = ops.build_reps(eq, ek, cq, ck, so2_nfreqs=cfg.so2, so3_maxdeg=cfg.so3)
    q, k, v = (inp[n].to(dev) for n in "qkv")
    tc = torch.tensor([0.01], device=dev)
    nct = B * cfg.heads * ((nq * tq + 255) // 256)
    for _ in range(3):
        ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc)
    dbg = torch.zeros(nct, 8, dtype=torch.int64, device=dev)
    ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, debug_clocks=dbg, flags=_lib.GTA_FLAG_V1_PIPELINE)
    torch.cuda.synchronize()
    d = dbg.cpu().double()
    t0 = d[:, 0].min()
    names = ["prologue(Q stage)", "first S latency", "main loop", "tail wait (o_final)", "epilogue"]
    ph = [d[:, i + 1] - d[:, i] for i in range(5)]
    tot = d[:, 5] - d[:, 0]
    print(f"{name} B={B}: {nct} CTAs, kernel span {(d[:,5].max()-t0):.0f} clk (clock64 is per-SM; span is approximate)")
    for nme, x in zip(names, ph):
        print(f"  {nme:22s} mean {x.mean():9.0f}  p50 {x.median():9.0f}  max {x.max():9.0f} clk   ({100*x.mean()/tot.mean():5.1f}% of CTA time)")
    print(f"  {'s_full wait in loop':22s} mean {d[:,6].mean():9.0f} clk ({100*d[:,6].mean()/tot.mean():5.1f}%)  per tile {(d[:,6]/max(1,(nk*tk+127)//128-1)).mean():.0f}")
    print(f"  CTA total mean {tot.mean():.0f} clk; per key tile in main loop {(ph[2]/((nk*tk+127)//128)).mean():.0f} clk (tensor work per tile-pair = {2*2*128*128*cfg.head_dim*2/8192:.0f} clk)")
    smid = d[:, 7].long()
    per_sm = torch.bincount(smid, minlength=148)
    print(f"  CTAs per SM: min {per_sm.min().item()} max {per_sm.max().item()}")


if __name__ == "__main__":
    main()
