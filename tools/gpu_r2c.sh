#!/bin/bash
# train-step lines with the fused backward, backward tests, compute-sanitizer on the backward kernels
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_reference_modules.py -m gpu -q -k "backward or train or module" 2>&1 | tail -3
for w in train_step_msn train_step_clevr; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 2 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w exit $?"; tail -c 900 gpurun_out/bench_$w.json; tail -2 gpurun_out/bench_$w.err
done
ARGS=backward TOOLS="racecheck synccheck memcheck" ST=500 bash tools/gpu_sanitizer.sh
