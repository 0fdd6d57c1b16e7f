#!/bin/bash
# session 3, call 12: reference point d639384 (before the SO(2) staging rows) vs the restructured staging rows (nothing live
# across the main loop; compiled out of the single-launch kernel at D = 96) vs the same sources with -DGTA_SO2_STAGE=0;
# both new libraries carry the one-warp-per-block view part of build_reps_kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "reps or golden or pipelines" 2>&1 | tail -2
for rep in 1 2; do
for wl in msn_enc msn_dec clevr_dec clevr_enc; do
  for lib in libgta_b200_d639384.so libgta_b200_noso2.so libgta_b200.so; do
    for fl in 0 ${FL2:-}; do
      GTA_B200_LIB=$PWD/gta_b200/$lib timeout 300 python bench.py --no-cpu --no-e2e --no-info --no-backward --steps 50 --flags $fl --workload $wl > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err
      python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_q.json")); r=d["roofline"]; print("$wl $lib flags=$fl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],4), "dom_kernel_ms", round(r["kernel_ms"],4), "call_ms", round(r.get("library_call_ms") or 0,4), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e: print("failed", e); print(open("gpurun_out/bench_q.err").read()[-1500:])
PY
    done
  done
done
done
