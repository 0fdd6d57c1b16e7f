#!/bin/bash
# default bench line (all legs), the reference arm, and the train-step workload on one GPU
mkdir -p gpurun_out
timeout 600 python bench.py --steps ${STEPS:-20} --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default exit $?"; tail -c 3000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"; tail -c 600 gpurun_out/bench_reference.json
for w in ${TRAIN:-train_step_msn train_step_clevr}; do timeout 600 python bench.py --workload $w --steps 5 --warmup 2 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w exit $?"; tail -c 1500 gpurun_out/bench_$w.json; tail -5 gpurun_out/bench_$w.err; done
