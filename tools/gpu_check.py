"""Bring-up checks run on the GPU box (each step in its own process so a sticky CUDA error or a hang in one
step does not hide the others).  Usage: python tools/gpu_check.py [step ...]; no args = all steps."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STEPS = ["probe", "reps", "rotate", "attn_small", "attn_shapes"]


def step_probe():
    import torch
    from gta_b200 import ops
    torch.manual_seed(0)
    for D in (32, 64, 96, 128):
        A = torch.randn(128, D, device="cuda").bfloat16()
        Bm = torch.randn(128, D, device="cuda").bfloat16()
        P = torch.rand(128, 128, device="cuda").bfloat16()
        V = torch.randn(128, D, device="cuda").bfloat16()
        refS = A.float() @ Bm.float().T
        refO = P.float() @ V.float()
        for tm in (False, True):
            S, O = ops.umma_probe(A, Bm, P, V, tm)
            torch.cuda.synchronize()
            print(f"probe D={D} p_in_tmem={tm}: errS={(S-refS).abs().max().item():.3e} errO={(O-refO).abs().max().item():.3e}"
                  f" (|S|max={refS.abs().max().item():.1f} |O|max={refO.abs().max().item():.1f})", flush=True)


def _case(base, nq, nk, tq, tk, cross, B, seed=0, dtype=None, vt=True):
    import torch
    from gta_b200.synth import GtaConfig, make_inputs
    cfg = GtaConfig(**base, n_q_views=nq, n_k_views=nk, v_transform=vt)
    inp = make_inputs(cfg, B, tq, tk, cross=cross, seed=seed, dtype=dtype or torch.float32)
    return cfg, inp


def _dev_reps(cfg, inp):
    from gta_b200 import ops
    c = lambda t: t.cuda()
    eq, ek = c(inp["extr_q"]), c(inp["extr_k"])
    cq, ck = c(inp["coord_q"]), c(inp["coord_k"])
    if inp["extr_q"] is inp["extr_k"]:
        eq = ek
    if inp["coord_q"] is inp["coord_k"]:
        cq = ck
    return ops.build_reps(eq, ek, cq, ck, so2_nfreqs=cfg.so2, so3_maxdeg=cfg.so3, max_freq_h=cfg.max_freq_h,
                          max_freq_w=cfg.max_freq_w, shared_freqs=cfg.shared_freqs)


def step_reps():
    import numpy as np
    from gta_b200.synth import MSN_SO3
    from oracle import c_oracle
    cfg, inp = _case(MSN_SO3, 3, 5, 16, 64, True, 2)
    r = _dev_reps(cfg, inp)
    o = c_oracle.build_reps(cfg, inp["extr_q"], inp["extr_k"], inp["coord_q"], inp["coord_k"])
    for k in ("se3_q", "se3_k", "so3_q", "so3_k", "so2_q", "so2_k"):
        print("reps", k, float(np.abs(getattr(r, k).cpu().numpy() - o[k]).max()), flush=True)


def step_rotate():
    import numpy as np
    import torch
    from gta_b200 import ops
    from gta_b200.synth import CLEVR, MSN_SO3
    from oracle import c_oracle
    for base, args in ((MSN_SO3, (3, 2, 8, 16, True, 2)), (CLEVR, (2, 2, 21, 21, False, 1))):
        for dt in (torch.float32, torch.bfloat16):
            cfg, inp = _case(base, *args, dtype=dt)
            r = _dev_reps(cfg, inp)
            tc = torch.tensor([0.3], device="cuda")
            qt, kt, vt = ops.rotate_debug(inp["q"].cuda(), inp["k"].cuda(), inp["v"].cuda(), r, cfg.f_dims, trans_coeff=tc)
            _, q2, k2, v2 = c_oracle.gta_attention(cfg, inp["q"].float(), inp["k"].float(), inp["v"].float(), inp["extr_q"],
                                                   inp["extr_k"], inp["coord_q"], inp["coord_k"], trans_coeff=0.3,
                                                   return_rotated=True)
            print("rotate", cfg.head_dim, dt, float(np.abs(qt.cpu().numpy() - q2).max()),
                  float(np.abs(kt.cpu().numpy() - k2).max()), float(np.abs(vt.cpu().numpy() - v2).max()), flush=True)


def _attn(base, args, dt, flags, tc=0.01, vt=True):
    import numpy as np
    import torch
    from gta_b200 import ops
    from oracle import c_oracle
    cfg, inp = _case(base, *args, dtype=dt, vt=vt)
    r = _dev_reps(cfg, inp)
    tct = torch.tensor([tc], device="cuda")
    out = ops.gta_attention_fwd(inp["q"].cuda(), inp["k"].cuda(), inp["v"].cuda(), r, cfg.f_dims, trans_coeff=tct,
                                v_transform=vt, flags=flags)
    torch.cuda.synchronize()
    ref = c_oracle.gta_attention(cfg, inp["q"].float(), inp["k"].float(), inp["v"].float(), inp["extr_q"], inp["extr_k"],
                                 inp["coord_q"], inp["coord_k"], trans_coeff=tc)
    err = float(np.abs(out.float().cpu().numpy() - ref).max())
    print(f"attn D={cfg.head_dim} args={args} dt={dt} flags={flags} tc={tc} vt={vt}: max-abs err {err:.3e} "
          f"(|ref| max {np.abs(ref).max():.2f})", flush=True)
    return err


def step_attn_small():
    import torch
    from gta_b200.synth import CFG1_A, CFG1_B, CLEVR, MSN_SO3
    for flags in (0, 16):
        _attn(CFG1_A, (2, 2, 64, 64, False, 1), torch.bfloat16, flags)           # exactly one tile
        _attn(CFG1_B, (2, 2, 16, 16, False, 2), torch.float32, flags)            # partial tile
        _attn(MSN_SO3, (5, 5, 64, 64, False, 1), torch.bfloat16, flags)          # 320 tokens: 3 tiles, tail 64
        _attn(CLEVR, (3, 2, 100, 150, True, 2), torch.bfloat16, flags, tc=1.0)   # cross, ragged


def step_attn_shapes():
    import torch
    from gta_b200.synth import CLEVR, MSN_SO3
    for flags in (0, 16):
        _attn(MSN_SO3, (5, 5, 256, 256, False, 1), torch.bfloat16, flags)
        _attn(MSN_SO3, (5, 5, 512, 256, True, 1), torch.bfloat16, flags)
        _attn(CLEVR, (2, 2, 300, 300, False, 2), torch.bfloat16, flags)
        _attn(CLEVR, (3, 2, 853, 300, True, 1), torch.float32, flags)
        _attn(CLEVR, (3, 2, 853, 300, True, 1), torch.bfloat16, flags, vt=False)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        globals()["step_" + sys.argv[2]]()
        sys.exit(0)
    steps = sys.argv[1:] or STEPS
    for s in steps:
        t = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", s], timeout=240,
                               capture_output=True, text=True)
            print(f"=== {s}: rc={p.returncode} ({time.time()-t:.1f}s)\n{p.stdout}{p.stderr[-3000:]}", flush=True)
        except subprocess.TimeoutExpired as e:
            print(f"=== {s}: TIMEOUT\n{(e.stdout or b'').decode() if isinstance(e.stdout, bytes) else e.stdout}", flush=True)
