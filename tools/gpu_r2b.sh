#!/bin/bash
# round-2 (second session): whole GPU suite, headline bench line with the fused backward, backward timings of every workload,
# launch list + ncu --set full capture of the fused backward kernel
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q -rA > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_full.log; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_full.log | head -20
timeout 600 python bench.py > gpurun_out/bench_msn_enc_r2b.json 2> gpurun_out/bench_msn_enc_r2b.err; tail -c 1500 gpurun_out/bench_msn_enc_r2b.json
rm -f gpurun_out/bwd2_bench.jsonl
for w in "msn_enc 64" "msn_dec 64" "clevr_enc 32" "clevr_dec 32" "cfg1 2" "sweep2 1"; do
  set -- $w
  for f in 0 4096; do
    GTA_BWD_FLAGS=$f timeout 300 python tools/bwd_bench.py $1 $2 | tee -a gpurun_out/bwd2_bench.jsonl | cut -c1-130
  done
done
timeout 200 python tools/bwd2_phase.py msn_enc 64 | tee gpurun_out/bwd2_phase_msn_enc.txt
timeout 200 python tools/bwd2_phase.py clevr_dec 32 | tee gpurun_out/bwd2_phase_clevr_dec.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bwd|rotate|delta" -c 40 --csv --log-file gpurun_out/launches_bwd2.csv python tools/bwd_bench.py msn_enc 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_fused -s 2 -c 1 -f -o gpurun_out/r02_bwd_fused python tools/bwd_bench.py msn_enc 64 > gpurun_out/ncu_bwd_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
