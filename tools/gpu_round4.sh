#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for wl in msn_enc msn_dec clevr_enc clevr_dec; do
  for fl in 0 16; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --steps 50 --flags $fl > gpurun_out/bench4_${wl}_$fl.json 2>gpurun_out/bench4_${wl}_$fl.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench4_${wl}_$fl.json")); r=d["roofline"]; print("$wl flags=$fl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],3), "attn_ms", round(r["kernel_ms"],3), "stage_ms", round(r["stage_kernel_ms"],3), "frac", round(r["frac"],3), d["clocks"])
except Exception as e: print("$wl failed", e); print(open("gpurun_out/bench4_${wl}_$fl.err").read()[-1500:])
PY
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd3 -s 3 -c 1 -f -o gpurun_out/prof_attn3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --batch 16 > gpurun_out/ncu_full3.log 2>&1
ls gpurun_out | tail -3
