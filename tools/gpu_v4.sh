#!/bin/bash
# v4 pipeline (streaming softmax + epilogue warpgroup): parity, then A/B bench against the other pipelines
mkdir -p gpurun_out
timeout ${PT:-240} python -m pytest tests -m gpu -q -x -k "${PYTEST_K:-v4_streaming}" 2>&1 | tail -5
[ "${PIPESTATUS[0]}" = "124" ] && { echo "PYTEST TIMED OUT (hang?)"; exit 1; }
for wl in ${WLS:-msn_enc}; do for fl in ${FLAGS:-256 32 0}; do
    timeout 200 python bench.py --no-cpu --no-e2e --steps 50 --flags $fl --workload $wl > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_q.json")); r=d["roofline"]; print("$wl flags=$fl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],4), "dom_kernel_ms", round(r["kernel_ms"],4), "attn2L_ms", round(r["two_launch_attention_kernel_ms"],4), "frac", round(r["frac"],3), "stage_ms", round(r["staging_kernel_ms"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("$wl flags=$fl failed", e); print(open("gpurun_out/bench_q.err").read()[-800:])
PY
done; done
