#!/bin/bash
mkdir -p gpurun_out
timeout ${PT:-240} python -m pytest tests -m gpu -q -x -k "${PYTEST_K:-v4_streaming}" 2>&1 | tail -3
for lib in ${LIBS:-libgta_b200.so}; do for wl in ${WLS:-msn_enc}; do for fl in ${FLAGS:-256 32}; do
    GTA_B200_LIB=$PWD/gta_b200/$lib timeout 200 python bench.py --no-cpu --no-e2e --no-info --no-backward --steps 50 --flags $fl --workload $wl > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_q.json")); r=d["roofline"]; print("$lib $wl flags=$fl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],4), "dom_kernel_ms", round(r["kernel_ms"],4), "frac", round(r["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("$lib $wl flags=$fl failed", e); print(open("gpurun_out/bench_q.err").read()[-800:])
PY
done; done; done
