mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "backward" 2>&1 | tail -4 || exit 1
timeout 200 python tools/bwd_bench.py msn_enc 64 | tee gpurun_out/bwd_bench_msn_enc.json | cut -c1-120
timeout 200 python tools/bwd_bench.py clevr_dec 32 | tee gpurun_out/bwd_bench_clevr_dec.json | cut -c1-120
