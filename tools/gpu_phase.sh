#!/bin/bash
mkdir -p gpurun_out
for lib in $LIBS; do GTA_B200_LIB=$PWD/gta_b200/$lib timeout 200 python tools/phase_timing2.py ${WL:-msn_enc} ${B:-64} 2>&1 | tail -22; done
