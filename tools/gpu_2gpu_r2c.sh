#!/bin/bash
# 2-GPU confirmation of the final library: default bench line + reference arm under torchrun
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-info > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"; tail -c 1500 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
timeout 200 $TR bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref exit $?"; tail -c 300 gpurun_out/bench_ref_n$N.json
