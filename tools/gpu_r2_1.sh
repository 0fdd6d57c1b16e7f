#!/bin/bash
# v3 pipeline (pre/post warpgroup) bring-up: smoke, targeted parity, phase clocks, variant benches.
mkdir -p gpurun_out
timeout 240 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "v3 or fp32 or strided or lse or identity or invariance or long_sequence or dropin" 2>&1 | tail -6 | tee gpurun_out/pytest_v3.log
timeout 200 python tools/phase_timing3.py msn_enc 64 2>&1 | tee gpurun_out/phase_v3_msn_enc.log
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
bench_one v3_default $L/libgta_b200.so
bench_one v2_flag32 $L/libgta_b200.so --flags 32
bench_one v3_nosplit $L/libgta_b200_nosplit.so
bench_one v3_p14 $L/libgta_b200_p14.so
bench_one v3_p13 $L/libgta_b200_p13.so
bench_one v3_p12 $L/libgta_b200_p12.so
bench_one v3_msn_dec $L/libgta_b200.so --workload msn_dec
bench_one v3_clevr_enc $L/libgta_b200.so --workload clevr_enc
bench_one v3_clevr_dec $L/libgta_b200.so --workload clevr_dec
GTA_B200_LIB=$L/libgta_b200_p13.so timeout 200 python tools/phase_timing3.py msn_enc 64 2>&1 | tee gpurun_out/phase_v3_p13_msn_enc.log
