#!/bin/bash
# session 3, call 4: L1 prefetch of the next Q row in the stagers (D <= 64) -- A/B against a build without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "golden or pipelines or clevr or ragged or cfg1 or config1 or fused_attention" 2>&1 | tail -3
for lib in libgta_b200.so libgta_b200_nopf.so; do
  GTA_B200_LIB=$PWD/gta_b200/$lib GTA_FLAGS=1024 timeout 200 python tools/phase_timing2.py clevr_dec 2>&1 | grep -E 'per item|per key tile' | head -8
done
for wl in clevr_dec clevr_enc cfg1; do
  for lib in libgta_b200.so libgta_b200_nopf.so; do
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
