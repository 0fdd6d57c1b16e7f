#!/bin/bash
# session 3, call 5: SO(2) table rows of the epilogue staged in shared memory (attn_fwd3_kernel) -- parity, phase clocks, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "not backward and not bwd and not train and not reference_modules" 2>&1 | tail -3
for lib in libgta_b200.so libgta_b200_noso2.so; do
  GTA_B200_LIB=$PWD/gta_b200/$lib GTA_FLAGS=1024 timeout 200 python tools/phase_timing2.py msn_enc 2>&1 | grep -E 'per item|per key tile' | head -10
  GTA_B200_LIB=$PWD/gta_b200/$lib GTA_FLAGS=1024 timeout 200 python tools/phase_timing2.py clevr_dec 2>&1 | grep -E 'per item|per key tile' | head -10
done
for wl in msn_enc msn_dec clevr_dec clevr_enc; do
  for lib in libgta_b200.so libgta_b200_noso2.so; do
    for fl in 1024; do
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
