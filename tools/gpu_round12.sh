#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 1500 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; tail -c 600 gpurun_out/bench_2gpu_ref.json; tail -3 gpurun_out/bench_2gpu_ref.err
for lib in p14 p13; do
  GTA_B200_LIB=$PWD/gta_b200/libgta_b200_$lib.so timeout 300 python bench.py --no-cpu --no-e2e --steps 50 > gpurun_out/bench12_$lib.json 2>gpurun_out/bench12_$lib.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench12_$lib.json")); r=d["roofline"]; print("$lib", round(d["value"],1), "Mtok/s attn_ms", round(r["kernel_ms"],3), "frac", round(r["frac"],3))
except Exception as e: print("$lib failed", e); print(open("gpurun_out/bench12_$lib.err").read()[-1500:])
PY
done
