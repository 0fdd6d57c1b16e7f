#!/bin/bash
# session 3, call 11: same-box bisection of the 4 % the final library lost against the session's first commit
mkdir -p gpurun_out
for rep in 1 2; do
for wl in msn_enc msn_dec clevr_dec; do
  for lib in libgta_b200_base.so libgta_b200_4cd65ef.so libgta_b200_d639384.so libgta_b200.so; do
      GTA_B200_LIB=$PWD/gta_b200/$lib timeout 300 python bench.py --no-cpu --no-e2e --no-info --no-backward --steps 50 --workload $wl > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err
      python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_q.json")); r=d["roofline"]; print("$wl $lib", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],4), "dom_kernel_ms", round(r["kernel_ms"],4), "call_ms", round(r.get("library_call_ms") or 0,4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("failed", e); print(open("gpurun_out/bench_q.err").read()[-1500:])
PY
  done
done
done
