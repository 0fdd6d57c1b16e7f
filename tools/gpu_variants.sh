#!/bin/bash
mkdir -p gpurun_out
for lib in $LIBS; do for wl in ${WLS:-msn_enc}; do
    GTA_B200_LIB=$PWD/gta_b200/$lib timeout 200 python bench.py --no-cpu --no-e2e --steps 50 --flags ${FL:-0} --workload $wl > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_q.json")); r=d["roofline"]; print("$lib $wl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("$lib $wl failed", e); print(open("gpurun_out/bench_q.err").read()[-800:])
PY
done; done
