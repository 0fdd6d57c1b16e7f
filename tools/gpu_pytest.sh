#!/bin/bash
mkdir -p gpurun_out
timeout ${PT:-600} python -m pytest tests -m gpu -q -rA -k "${PYTEST_K}" > gpurun_out/pytest_sel.log 2>&1; echo "exit $?"
grep -E "^(PASSED|FAILED|ERROR)" gpurun_out/pytest_sel.log | sed 's/ - .*//' | head -${NL:-80}; tail -2 gpurun_out/pytest_sel.log
