#!/bin/bash
# session 3, call 7: same-box A/B of the SO(2) staging rows, both pipelines (interleaved twice to see the run-to-run spread)
mkdir -p gpurun_out
for rep in 1 2; do
for wl in msn_dec clevr_dec msn_enc clevr_enc; do
  for lib in libgta_b200.so libgta_b200_noso2.so; do
    for fl in 32 1024; do
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
done
