#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/phase_timing2.py msn_enc 64 2>&1 | tee gpurun_out/phase5_msn_enc.log
for wl in msn_enc msn_dec clevr_enc clevr_dec; do
  timeout 300 python bench.py --workload $wl --no-cpu --no-e2e --steps 50 > gpurun_out/bench7_${wl}.json 2>gpurun_out/bench7_${wl}.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench7_${wl}.json")); r=d["roofline"]; print("$wl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],3), "attn_ms", round(r["kernel_ms"],3), "stage_ms", round(r["stage_kernel_ms"],3), "frac", round(r["frac"],3))
except Exception as e: print("$wl failed", e); print(open("gpurun_out/bench7_${wl}.err").read()[-1500:])
PY
done
for lib in p14 p13 p12; do
  GTA_B200_LIB=$PWD/gta_b200/libgta_b200_$lib.so timeout 300 python bench.py --no-cpu --no-e2e --steps 50 > gpurun_out/bench7_$lib.json 2>gpurun_out/bench7_$lib.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench7_$lib.json")); r=d["roofline"]; print("$lib", round(d["value"],1), "Mtok/s attn_ms", round(r["kernel_ms"],3), "frac", round(r["frac"],3))
except Exception as e: print("$lib failed", e); print(open("gpurun_out/bench7_$lib.err").read()[-1500:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd5 -s 3 -c 1 -f -o gpurun_out/prof_attn5 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full5.log 2>&1
