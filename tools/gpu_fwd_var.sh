#!/bin/bash
mkdir -p gpurun_out
for v in "" _p14 _hoist; do
  for wl in msn_enc msn_dec clevr_dec; do
    GTA_B200_LIB=$PWD/gta_b200/libgta_b200$v.so timeout 200 python bench.py --no-cpu --no-e2e --no-info --no-backward --steps 50 --workload $wl > gpurun_out/bench_var.json 2>gpurun_out/bench_var.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_var.json")); r=d["roofline"]; print("lib$v $wl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],4), "dom_kernel_ms", round(r["kernel_ms"],4), "frac", round(r["frac"],3), "step_frac", round(r["step_frac"],3))
except Exception as e: print("lib$v $wl failed", e); print(open("gpurun_out/bench_var.err").read()[-500:])
PY
  done
done
GTA_B200_LIB=$PWD/gta_b200/libgta_b200_p14.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or msn or sweep" 2>&1 | tail -2
