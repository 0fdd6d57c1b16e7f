#!/bin/bash
mkdir -p gpurun_out
for v in "" _r6 _r8 _r7u2 _u4; do
  for wl in msn_enc clevr_dec; do
    GTA_B200_LIB=$PWD/gta_b200/libgta_b200$v.so timeout 200 python bench.py --no-cpu --no-e2e --no-info --no-backward --steps 50 --flags 1024 --workload $wl > gpurun_out/bench_var.json 2>gpurun_out/bench_var.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_var.json")); r=d["roofline"]; print("lib$v $wl step_ms", round(d["ms_per_step"],4), "attn_ms", round(r["kernel_ms"],4), "stage_ms", round(r["staging_kernel_ms"],4))
except Exception as e: print("lib$v $wl failed", e); print(open("gpurun_out/bench_var.err").read()[-500:])
PY
  done
done
