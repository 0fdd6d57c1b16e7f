#!/bin/bash
# Quick verification: smoke (forward + backward vs the oracle) and the whole GPU parity suite.
mkdir -p gpurun_out
set -o pipefail
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -5 || { echo 'SMOKE FAILED'; exit 1; }
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
