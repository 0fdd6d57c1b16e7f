#!/bin/bash
# quick loop: parity subset of the fused kernel, phase clocks, headline bench (fused and two-launch)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "${PYTEST_K:-fused_attention_matches_oracle or bf16_pipelines or many_items or headline}" 2>&1 | tail -3
for lib in ${LIBS:-libgta_b200.so}; do
  GTA_B200_LIB=$PWD/gta_b200/$lib timeout 200 python tools/phase_timing2.py ${WL:-msn_enc} ${B:-64} 2>&1 | tail -19
  for fl in 0 ${FLAGS2:-32}; do
    GTA_B200_LIB=$PWD/gta_b200/$lib timeout 300 python bench.py --no-cpu --no-e2e --steps 50 --flags $fl --workload ${WL:-msn_enc} > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_q.json")); r=d["roofline"]; print("$lib flags=$fl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],4), "dom_kernel_ms", round(r["kernel_ms"],4), "attn2L_ms", round(r["two_launch_attention_kernel_ms"],4), "stage_ms", round(r["staging_kernel_ms"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("failed", e); print(open("gpurun_out/bench_q.err").read()[-1500:])
PY
  done
done
