#!/bin/bash
# Fused single-launch kernel: parity subset, then bench lines for register-split variants and the two-launch baseline.
mkdir -p gpurun_out
set -o pipefail
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -4 || { echo 'SMOKE FAILED'; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -x -rA -k "${PYTEST_K:-fused or golden or sweep or headline or many_items or long_sequence}" > gpurun_out/pytest_fused.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_fused.log
bench_one() {  # name, lib, extra args
  local name=$1 lib=$2; shift 2
  GTA_B200_LIB=$lib timeout 300 python bench.py --no-cpu --no-e2e --steps 50 "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); r=d["roofline"]; print("$name", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],4), "dom_kernel_ms", round(r["kernel_ms"],4), "attn2L_ms", round(r["two_launch_attention_kernel_ms"],4), "stage_ms", round(r["staging_kernel_ms"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/bench_$name.err").read()[-1500:])
PY
}
L=$PWD/gta_b200
bench_one fused_default $L/libgta_b200.so
bench_one two_launch $L/libgta_b200.so --flags 32
for v in $VARIANTS; do [ -f $L/libgta_b200_$v.so ] && bench_one fused_$v $L/libgta_b200_$v.so; done
timeout 200 python tools/phase_timing2.py msn_enc 64 2>&1 | tee gpurun_out/phase_fused.log | tail -20
for wl in $WORKLOADS_EXTRA; do bench_one $wl $L/libgta_b200.so --workload $wl; bench_one ${wl}_2l $L/libgta_b200.so --workload $wl --flags 32; done
