#!/bin/bash
# whole GPU suite (with measured errors) + per-workload A/B of the attention pipelines
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 1500 python -m pytest tests -m gpu -q -rA > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_full.log; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_full.log | head -20
for wl in ${WLS:-msn_enc msn_dec clevr_enc clevr_dec cfg1}; do for fl in ${FLAGS:-0 256 32}; do
    timeout 200 python bench.py --no-cpu --no-e2e --no-info --no-backward --steps 50 --flags $fl --workload $wl > gpurun_out/bench_ab_${wl}_$fl.json 2>gpurun_out/bench_q.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_ab_${wl}_$fl.json")); r=d["roofline"]; print("$wl flags=$fl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],4), "dom_kernel_ms", round(r["kernel_ms"],4), "frac", round(r["frac"],3), "step_frac", round(r["step_frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("$wl flags=$fl failed", e); print(open("gpurun_out/bench_q.err").read()[-800:])
PY
done; done
