"""Throughput of the softmax exp2/sum/pack phase alone (no tensor work): MUFU vs polynomial mix."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gta_b200 import _lib
l = _lib.dev_lib()
inp = torch.randn(1024, device="cuda")
for warps in (4, 8):
    for num, den in ((0, 4), (1, 4), (1, 3), (1, 2), (2, 3), (1, 1)):
        reps = 50
        out = torch.zeros(148 * warps * 32, device="cuda")
        clk = torch.zeros(148, dtype=torch.int64, device="cuda")
        for _ in range(2):
            _lib.check_dev(l.gta_dev_softmax_bench(num, den, warps, reps, 148, inp.data_ptr(), out.data_ptr(), clk.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        c = clk.double().mean().item() / reps
        print(f"warps/CTA={warps} ({warps//4}/SMSP) poly {num}/{den}: {c:7.0f} clk per 128-col row-tile per warp "
              f"-> {c * (1 if warps == 4 else 1):7.0f} clk for {warps//4} tile(s) per SMSP", flush=True)
