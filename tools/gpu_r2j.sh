#!/bin/bash
# session 3, call 6: SO(2) staging rows in both attention kernels -- full forward parity, default pipeline of every workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "not backward and not bwd and not train" 2>&1 | tail -3
for wl in msn_enc msn_dec clevr_dec clevr_enc cfg1 sweep2; do
  for fl in 0 ${FL2:-}; do
    timeout 300 python bench.py --no-cpu --no-e2e --no-info --no-backward --steps 50 --flags $fl --workload $wl > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_q.json")); r=d["roofline"]; print("$wl flags=$fl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],4), "dom_kernel_ms", round(r["kernel_ms"],4), "frac", round(r["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("failed", e); print(open("gpurun_out/bench_q.err").read()[-1500:])
PY
  done
done
