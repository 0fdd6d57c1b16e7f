mkdir -p gpurun_out
timeout 150 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "backward" 2>&1 | tail -25
