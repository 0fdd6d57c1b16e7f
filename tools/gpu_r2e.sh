#!/bin/bash
# session 3, call 1: merged rep-construction launch (parity + headline step), phase clocks of the CLEVR shapes
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "reps or so2 or golden or dropin or headline" 2>&1 | tail -3
for wl in clevr_dec clevr_enc msn_dec; do
  for fl in 0 1024; do
    GTA_FLAGS=$fl timeout 200 python tools/phase_timing2.py $wl 2>&1 | tail -20
  done
done
for wl in msn_enc clevr_dec clevr_enc; do
  timeout 300 python bench.py --no-cpu --no-e2e --no-info --no-backward --steps 50 --workload $wl > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_q.json")); r=d["roofline"]; print("$wl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],4), "dom_kernel_ms", round(r["kernel_ms"],4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], d.get("gpu_launches"))
except Exception as e: print("failed", e); print(open("gpurun_out/bench_q.err").read()[-1500:])
PY
done
