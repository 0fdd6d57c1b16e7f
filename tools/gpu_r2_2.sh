#!/bin/bash
# v5 pipeline (four softmax warpgroups, column-split score tiles) bring-up
mkdir -p gpurun_out
set -o pipefail
timeout 150 python __graft_entry__.py smoke 2>&1 | tail -3 || { echo 'SMOKE FAILED - stopping'; exit 1; }
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "v5 or fp32 or strided or lse or identity or invariance or long_sequence or dropin or ablation or generic" 2>&1 | tail -15 | tee gpurun_out/pytest_v5.log || { echo "PYTEST FAILED - stopping"; exit 1; }
timeout 200 python tools/phase_timing5.py msn_enc 64 2>&1 | tee gpurun_out/phase_v5_msn_enc.log
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
bench_one v5_default $L/libgta_b200.so
bench_one v2_flag32 $L/libgta_b200.so --flags 32
for v in p14 p13 p12; do [ -f $L/libgta_b200_$v.so ] && bench_one v5_$v $L/libgta_b200_$v.so; done
bench_one v5_msn_dec $L/libgta_b200.so --workload msn_dec
bench_one v5_clevr_enc $L/libgta_b200.so --workload clevr_enc
bench_one v5_clevr_dec $L/libgta_b200.so --workload clevr_dec
bench_one v5_sweep2 $L/libgta_b200.so --workload sweep2 --steps 10
