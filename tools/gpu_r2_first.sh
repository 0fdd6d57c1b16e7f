#!/bin/bash
# Round-2 first GPU session: smoke, the whole GPU parity suite (with the measured errors), baseline bench lines.
mkdir -p gpurun_out
set -o pipefail
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -5 || { echo 'SMOKE FAILED'; exit 1; }
timeout 1500 python -m pytest tests -m gpu -q -rA > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_gpu_full.log
grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_full.log | head -40
timeout 300 python bench.py --steps 50 --no-cpu > gpurun_out/bench_r2_base.json 2> gpurun_out/bench_r2_base.err; tail -c 1500 gpurun_out/bench_r2_base.json
timeout 200 python tools/phase_timing2.py msn_enc 64 2>&1 | tee gpurun_out/phase_msn_enc_r2base.log | tail -20
