#!/bin/bash
# One GPU session: smoke (fail fast), parity suite, phase clocks, bench lines.  usage: gpu_run.sh [quick]
mkdir -p gpurun_out
set -o pipefail
timeout 150 python __graft_entry__.py smoke 2>&1 | tail -3 || { echo 'SMOKE FAILED - stopping'; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log || { echo "PYTEST FAILED - stopping"; exit 1; }
timeout 200 python tools/phase_timing2.py msn_enc 64 2>&1 | tee gpurun_out/phase_msn_enc.log
bench_one() {  # name, lib, extra args
  local name=$1 lib=$2; shift 2
  GTA_B200_LIB=$lib timeout 300 python bench.py --no-cpu --no-e2e --steps 50 "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); r=d["roofline"]; print("$name", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],3), "attn_ms", round(r["kernel_ms"],4), "stage_ms", round(r["stage_kernel_ms"],3), "frac", round(r["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("$name failed", e); print(open("gpurun_out/bench_$name.err").read()[-1500:])
PY
}
L=$PWD/gta_b200
bench_one default $L/libgta_b200.so
for v in lds p13 lds_p14; do [ -f $L/libgta_b200_$v.so ] && bench_one $v $L/libgta_b200_$v.so; done
[ "$1" = "quick" ] && exit 0
for wl in msn_dec clevr_enc clevr_dec cfg1 sweep2; do bench_one $wl $L/libgta_b200.so --workload $wl; done
