#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/phase_timing2.py msn_enc 64 2>&1 | tee gpurun_out/phase2_msn_enc.log
for lib in p14 p13 p12; do
  GTA_B200_LIB=$PWD/gta_b200/libgta_b200_$lib.so timeout 300 python tools/phase_timing2.py msn_enc 64 2>&1 | tee gpurun_out/phase2_msn_enc_$lib.log
  GTA_B200_LIB=$PWD/gta_b200/libgta_b200_$lib.so timeout 300 python bench.py --no-cpu --no-e2e --steps 50 > gpurun_out/bench5_$lib.json 2>gpurun_out/bench5_$lib.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench5_$lib.json")); r=d["roofline"]; print("$lib", round(d["value"],1), "Mtok/s attn_ms", round(r["kernel_ms"],3), "frac", round(r["frac"],3))
except Exception as e: print("$lib failed", e); print(open("gpurun_out/bench5_$lib.err").read()[-1500:])
PY
done
GTA_B200_LIB=$PWD/gta_b200/libgta_b200_p13.so timeout 300 python tools/gpu_check.py attn_shapes 2>&1 | tail -12
