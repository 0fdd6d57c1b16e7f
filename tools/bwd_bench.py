"""Timing of the fused backward (gta_attn_bwd) next to the forward at a bench.py workload shape.
usage: bwd_bench.py [workload] [B]   -> one JSON line (CUDA events, 20 timed calls after 3 warm-ups)."""
import json
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

    def timed(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    ms_f = timed(lambda: ops.gta_attention_fwd(q, k, v, reps, cfg.f_dims, trans_coeff=tc, return_lse=True))
    flags = int(os.environ.get("GTA_BWD_FLAGS", "0"))      # 4096 = GTA_FLAG_BWD_SPLIT (dK/dV kernel + dQ kernel)
    ms_b = timed(lambda: ops.gta_attention_bwd(dout, q, k, v, out, lse, reps, cfg.f_dims, trans_coeff=tc, flags=flags))
    H, D, Tq, Tk = cfg.heads, cfg.head_dim, nq * tq, nk * tk
    f_fwd = 4.0 * B * H * Tq * Tk * D
    print(json.dumps({"workload": name, "batch": B, "bwd_flags": flags, "fwd_ms": ms_f, "bwd_ms": ms_b,
                      "fwd_tflops": f_fwd / ms_f / 1e9, "bwd_tflops_algorithmic(2.5x fwd flops)": 2.5 * f_fwd / ms_b / 1e9,
                      "bwd_tflops_executed(3.5x: S and dP recomputed in both kernels)": 3.5 * f_fwd / ms_b / 1e9,
                      "fwd_bwd_Mtok_s": B * Tq / (ms_f + ms_b) / 1e3}))


if __name__ == "__main__":
    main()
