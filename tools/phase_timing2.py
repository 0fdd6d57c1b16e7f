"""Phase clocks of the persistent attention kernel (v2): softmax warpgroup A thread 0 and the UMMA issuer lane
accumulate clock64 deltas over the CTA's items into GtaAttnParams.debug_clocks [grid][16]."""
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
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk, v_transform=os.environ.get('V_TRANSFORM', '1') == '1')
    inp = make_inputs(cfg, B, tq, tk, cross=cross, seed=0, dtype=torch.bfloat16)
    dev = torch.device("cuda")
    ek, ck = inp["extr_k"].to(dev), inp["coord_k"].to(dev)
    eq = inp["extr_q"].to(dev) if cross else ek
    cq = inp["coord_q"].to(dev) if cross else ck
    reps = ops.build_reps(eq, ek, cq, ck, so2_nfreqs=cfg.so2, so3_maxdeg=cfg.so3)
    q, k, v = (inp[n].to(dev) for n in "qkv")
    tc = torch.tensor([0.01], device=dev)
    flags = int(os.environ.get('GTA_FLAGS', '0'))
    for _ in range(3):
        ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, v_transform=cfg.v_transform, flags=int(os.environ.get('GTA_FLAGS', '0')))
    dbg = torch.zeros(2 * 148, 16, dtype=torch.int64, device=dev)     # [softmax/issuer plane | staging-warp plane]
    ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, v_transform=cfg.v_transform, debug_clocks=dbg, flags=int(os.environ.get('GTA_FLAGS', '0')))
    torch.cuda.synchronize()
    d_all = dbg.cpu().double()
    grid_n = int((d_all[:148, 5] > 0).sum())
    sd = d_all[grid_n:2 * grid_n]
    d = d_all[:148]
    d = d[d[:, 5] > 0]
    items = d[:, 5]
    ntile = (nk * tk + 127) // 128
    print(f"{name} B={B} lib={os.environ.get('GTA_B200_LIB','default')}: {len(d)} CTAs, items/CTA {items.min():.0f}..{items.max():.0f}")
    print(f"  CTA span (softmax A)   mean {d[:,0].mean():10.0f} clk   per item {(d[:,0]/items).mean():8.0f}")
    for i, nme in ((1, "main loops"), (2, "epilogues (after o_final)"), (3, "  s_full waits (in loops)"), (4, "o_final waits")):
        print(f"  {nme:28s} per item {(d[:,i]/items).mean():8.0f} clk  ({100*(d[:,i]/d[:,0]).mean():5.1f}% of span)")
    print(f"  per key tile: loop {(d[:,1]/items/ntile).mean():.0f} clk, of which s_full wait {(d[:,3]/items/ntile).mean():.0f};"
          f" tensor work per tile-pair {2*2*128*128*cfg.head_dim*2/8192:.0f} clk")
    if True:
        for i, nme in ((6, "setup (1/l, first tmem ld)"), (7, "triv+se3 chunks"), (13, "so3 chunks"), (14, "so2 chunks + o_free + lse")):
            print(f"  epilogue part {nme:30s} per item {(d[:,i]/items).mean():7.0f} clk")
    for i, nme in ((8, "k_full"), (9, "v_full"), (10, "p_full"), (11, "o_free"), (12, "q_full")):
        print(f"  UMMA issuer wait on {nme:7s} per item {(d[:,i]/items).mean():8.0f} clk ({100*(d[:,i]/d[:,0]).mean():5.1f}% of span)")
    if sd[:, 0].sum() > 0:
        it = items.mean()
        print(f"  staging warps: span {sd[:,0].mean():.0f} clk; per item: K'/V' units {(sd[:,1]).mean()/it:.0f} clk "
              f"({sd[:,3].mean()/it:.2f} units, {(sd[:,1]/sd[:,3].clamp(min=1)).mean():.0f} clk each), unit barrier {(sd[:,2]).mean()/it:.0f}, "
              f"Q' staging {(sd[:,4]).mean()/it:.0f}, q_free wait {(sd[:,5]).mean()/it:.0f}")


if __name__ == "__main__":
    main()
