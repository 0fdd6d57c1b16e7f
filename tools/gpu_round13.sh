#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for wl in msn_enc msn_dec clevr_enc clevr_dec; do
  timeout 120 python bench.py --workload $wl --no-cpu --no-e2e --steps 50 > gpurun_out/bench13_${wl}.json 2>gpurun_out/bench13_${wl}.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench13_${wl}.json")); r=d["roofline"]; print("$wl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],3), "attn_ms", round(r["kernel_ms"],3), "stage_ms", round(r["stage_kernel_ms"],3), "frac", round(r["frac"],3))
except Exception as e: print("$wl failed", e); print(open("gpurun_out/bench13_${wl}.err").read()[-800:])
PY
done
