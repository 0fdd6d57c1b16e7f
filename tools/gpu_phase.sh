#!/bin/bash
mkdir -p gpurun_out
for fl in ${FLAGS:-512 32}; do echo "== flags $fl"; GTA_FLAGS=$fl timeout 200 python tools/phase_timing2.py ${WL:-msn_enc} ${B:-64} 2>&1 | grep -v "epilogue part" | tail -14; done
