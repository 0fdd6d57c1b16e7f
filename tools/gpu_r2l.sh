#!/bin/bash
# session 3, call 10: final library vs the library of the session's first commit (22b6cff) on the SAME box, default pipelines,
# interleaved three times (boxes of this pool differ by ~4 % at the headline shape)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,serial,power.limit,clocks.max.sm --format=csv,noheader
for rep in 1 2 3; do
for wl in msn_enc msn_dec clevr_dec clevr_enc; do
  for lib in libgta_b200.so libgta_b200_base.so; do
      GTA_B200_LIB=$PWD/gta_b200/$lib timeout 300 python bench.py --no-cpu --no-e2e --no-info --no-backward --steps 50 --workload $wl > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err
      python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_q.json")); r=d["roofline"]; print("$wl $lib", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],4), "dom_kernel_ms", round(r["kernel_ms"],4), "frac", round(r["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("failed", e); print(open("gpurun_out/bench_q.err").read()[-1500:])
PY
  done
done
done
