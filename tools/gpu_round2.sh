#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_check.py attn_small attn_shapes > gpurun_out/check2.log 2>&1; tail -c 2500 gpurun_out/check2.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for wl in msn_enc msn_dec clevr_enc clevr_dec; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-e2e > gpurun_out/bench2_$wl.json 2>gpurun_out/bench2_$wl.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench2_$wl.json")); r=d["roofline"]; print("$wl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],3), "attn_ms", round(r["kernel_ms"],3), "stage_ms", round(r["stage_kernel_ms"],3), "frac", round(r["frac"],3), d["clocks"])
except Exception as e: print("$wl failed", e); print(open("gpurun_out/bench2_$wl.err").read()[-1500:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2 -s 3 -c 1 -f -o gpurun_out/prof_attn2 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --batch 16 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out | tail -5
