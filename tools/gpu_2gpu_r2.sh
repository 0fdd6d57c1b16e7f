#!/bin/bash
mkdir -p gpurun_out
N=${N:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-info > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"; tail -c 1200 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
timeout 600 $TR bench.py --gpus $N --workload train_step_msn --steps 5 --warmup 2 > gpurun_out/bench_train_msn_n$N.json 2> gpurun_out/bench_train_msn_n$N.err; echo "train exit $?"; tail -c 1500 gpurun_out/bench_train_msn_n$N.json; tail -3 gpurun_out/bench_train_msn_n$N.err
timeout 300 $TR bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref exit $?"; tail -c 300 gpurun_out/bench_ref_n$N.json
