#!/bin/bash
# one ncu --set full capture of a kernel (regex $KERNEL) from a short bench run; report -> gpurun_out/$OUT.ncu-rep
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-attn_fwd4} -s ${SKIP:-2} -c 1 -f -o gpurun_out/${OUT:-prof} \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --workload ${WL:-msn_enc} --flags ${FLAGS:-0} > gpurun_out/ncu_${OUT:-prof}.log 2>&1
tail -3 gpurun_out/ncu_${OUT:-prof}.log; ls -la gpurun_out/*.ncu-rep
