"""Phase clocks of the fused backward kernel (thread 0 of every CTA; gta_attn_bwd2.cu): per query tile, the waits and the
two SIMT phases of the compute warpgroups; plus the epilogue.   usage: bwd2_phase.py [workload] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from gta_b200 import ops  # noqa: E402
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
    out, lse = ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, return_lse=True)
    dout = torch.randn(out.shape, device=dev).bfloat16()
    H = cfg.heads
    ntk = (nk * tk + 127) // 128
    for _ in range(2):
        ops.gta_attention_bwd(dout, q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc)
    dbg = torch.zeros(B * H * ntk, 16, dtype=torch.int64, device=dev)
    ops.gta_attention_bwd(dout, q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc, debug_clocks=dbg)
    torch.cuda.synchronize()
    d = dbg.cpu().double()
    n = d[:, 6]
    D = cfg.head_dim
    print(f"{name} B={B} fused backward kernel: {len(d)} CTAs x {n.mean():.0f} query tiles, CTA span {d[:, 0].mean():.0f} clk;"
          f" loop {(d[:, 8] / n).mean():.0f} clk per tile (tensor work per tile {2 * (D // 16) * 64 + 24 * (D * 51 // 96)} clk)")
    for i, lab in ((1, "wait for S^T"), (12, "  of the next: tcgen05.ld S"), (2, "tcgen05.ld S + exp2 + pack P"), (3, "wait P^T/dS^T free + tcgen05.st P"),
                   (4, "wait for dP^T"), (5, "tcgen05.ld dP + dS + st.shared + arrive"), (7, "next-tile statistics + barrier")):
        print(f"    {lab:42s} {(d[:, i] / n).mean():7.0f} clk per tile")
    print(f"    epilogue: operand prefetch {d[:, 13].mean():.0f} + wait for the last MMAs {(d[:, 9] - d[:, 13]).mean():.0f}, drain to shared memory {d[:, 10].mean():.0f}, rotate + store {d[:, 11].mean():.0f} clk per CTA")


if __name__ == "__main__":
    main()
