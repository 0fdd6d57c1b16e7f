#!/bin/bash
# session 3, call 3: polynomial exp2 share on the D=64 (CLEVR) shapes, where both softmax warpgroups are always in their
# exponential phase (tensor work 1024 clk per key-tile pair against 2048 clk of MUFU)
mkdir -p gpurun_out
for wl in clevr_dec clevr_enc msn_dec; do
  for lib in libgta_b200.so libgta_b200_p14.so libgta_b200_p13.so libgta_b200_p12.so; do
    for fl in 1024 32; do
      GTA_B200_LIB=$PWD/gta_b200/$lib timeout 300 python bench.py --no-cpu --no-e2e --no-info --no-backward --steps 50 --flags $fl --workload $wl > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err
      python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_q.json")); r=d["roofline"]; print("$wl $lib flags=$fl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],4), "dom_kernel_ms", round(r["kernel_ms"],4), "frac", round(r["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("failed", e); print(open("gpurun_out/bench_q.err").read()[-1500:])
PY
    done
  done
done
timeout 300 python -m pytest tests -m gpu -q -x -k "golden or pipelines" 2>&1 | tail -3
