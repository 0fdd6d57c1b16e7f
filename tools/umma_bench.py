"""tcgen05.mma throughput of the exact shapes/layouts the attention kernels issue (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gta_b200 import _lib

l = _lib.dev_lib()
names = {0: "QK SS N=128 (D/16 MMAs)", 1: "QK SS N=64 (D/16 MMAs)", 2: "PV TS N=D (8 MMAs)", 3: "PV SS N=D (8 MMAs)",
         4: "PV TS N=D (4 MMAs)"}
for D in (96, 64, 128):
    for grid in (1, 148):
        for mode in range(5):
            reps = 200
            out = torch.zeros(grid, 2, dtype=torch.int64, device="cuda")
            for _ in range(2):
                _lib.check_dev(l.gta_dev_umma_bench(D, mode, reps, grid, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            o = out.double().cpu()
            nm = (D // 16) if mode < 2 else (8 if mode < 4 else 4)
            flop = 2 * 128 * (128 if mode == 0 else 64 if mode == 1 else D) * 16
            tot = o[:, 1].mean().item() / (reps * nm)
            print(f"D={D} grid={grid:3d} {names[mode]:26s}: issue {o[:,0].mean().item()/(reps*nm):6.1f} clk/MMA, "
                  f"total {tot:6.1f} clk/MMA  ({flop/tot:6.0f} FLOP/clk/SM; spec 8192)", flush=True)
