mkdir -p gpurun_out
timeout 200 python tools/bwd_bench.py msn_enc 64 | tee gpurun_out/bwd_bench_msn_enc.json
timeout 200 python tools/bwd_bench.py clevr_dec 32 | tee gpurun_out/bwd_bench_clevr_dec.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_bwd.csv python tools/bwd_bench.py msn_enc 16 > gpurun_out/ncu_launches_bwd.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 2 -c 2 -f -o gpurun_out/prof_bwd python tools/bwd_bench.py msn_enc 16 > gpurun_out/ncu_full_bwd.log 2>&1
ls -la gpurun_out | tail -5
