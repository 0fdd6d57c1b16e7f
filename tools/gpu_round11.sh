#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
GTA_FLAGS=64 timeout 200 python tools/phase_timing2.py msn_enc 64 2>&1 | head -12
for fl in 64 0; do
for wl in msn_enc msn_dec clevr_dec; do
  timeout 120 python bench.py --workload $wl --no-cpu --no-e2e --steps 50 --flags $fl > gpurun_out/bench11_${wl}_$fl.json 2>gpurun_out/bench11_${wl}_$fl.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench11_${wl}_$fl.json")); r=d["roofline"]; print("$wl flags=$fl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],3), "attn_ms", round(r["kernel_ms"],3), "stage_ms", round(r["stage_kernel_ms"],3), "frac", round(r["frac"],3))
except Exception as e: print("$wl flags=$fl failed", e); print(open("gpurun_out/bench11_${wl}_$fl.err").read()[-800:])
PY
done
done
