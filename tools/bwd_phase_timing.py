"""Phase clocks of the backward kernels (thread 0 of the first compute warpgroup of every CTA): per streamed tile,
the wait for S/dP, the TMEM load, the P/dS computation, the TMEM store + hand-off; plus the epilogue.
usage: bwd_phase_timing.py [workload] [B]"""
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
    ntq, ntk = (nq * tq + 127) // 128, (nk * tk + 127) // 128
    for _ in range(2):
        ops.gta_attention_bwd(dout, q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc)
    dbg = torch.zeros(B * H * (ntk + ntq), 16, dtype=torch.int64, device=dev)
    ops.gta_attention_bwd(dout, q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc, debug_clocks=dbg)
    torch.cuda.synchronize()
    d = dbg.cpu().double()
    for nme, part in (("dKV kernel (CTA per key tile)", d[:B * H * ntk]), ("dQ kernel (CTA per query tile)", d[B * H * ntk:])):
        n = part[:, 6]
        print(f"{name} B={B} {nme}: {len(part)} CTAs x {n.mean():.0f} tiles, CTA span {part[:, 0].mean():.0f} clk = {(part[:, 0] / n).mean():.0f} per tile"
              f" (tensor work per tile {(12 * 64 + (16 if 'dKV' in nme else 8) * 51)} clk)")
        for i, lab in ((1, "wait for S/dP (MMA + wake-up)"), (2, "tcgen05.ld 4 x 32 columns"), (3, "P / dS computation"), (4, "tcgen05.st + fence + arrive")):
            print(f"    {lab:32s} {(part[:, i] / n).mean():7.0f} clk per tile")
        print(f"    {'next-tile statistics + barrier':32s} {(part[:, 12] / n).mean():7.0f} clk per tile;  whole loop {(part[:, 13] / n).mean():.0f} clk per tile")
        print(f"    epilogue parts: setup+chunk loop {part[:, 11].mean():.0f} (tcgen05.wait::ld {part[:, 8].mean():.0f}, rotate+stage {part[:, 9].mean():.0f}), barrier + coalesced store {part[:, 10].mean():.0f}")
        print(f"    {'epilogue (after the last tile)':32s} {part[:, 5].mean():7.0f} clk per CTA, of which {part[:, 7].mean():.0f} waiting for the last MMAs")


if __name__ == "__main__":
    main()
