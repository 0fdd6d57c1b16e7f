#!/bin/bash
GTA_B200_LIB=$PWD/gta_b200/libgta_b200_dbg.so timeout 200 python -m pytest tests -m gpu -q -x -k "${PYTEST_K}" > gpurun_out/dbg.log 2>&1
grep -E "passed|failed" gpurun_out/dbg.log | tail -2
grep "timed out" gpurun_out/dbg.log | awk '{print "block",$7,"warp",int($9/32),"bar",$10,"parity",$12}' | sort | uniq -c | sort -k3n -k5n | head -${NL:-60}
