#!/bin/bash
# A/B of build variants of the fused backward: bwd_bench + phase clocks per library
for v in ${VARIANTS:-main}; do
  [ "$v" = "main" ] && v=""
  echo "== variant libgta_b200$v.so"
  GTA_B200_LIB=$PWD/gta_b200/libgta_b200$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fused_kernel_matches_kernel_pair or (backward and msn_enc)" 2>&1 | tail -1
  for w in ${WL:-msn_enc:64 clevr_dec:32}; do
    GTA_B200_LIB=$PWD/gta_b200/libgta_b200$v.so timeout 200 python tools/bwd_bench.py ${w%%:*} ${w##*:} | cut -c1-110
    GTA_B200_LIB=$PWD/gta_b200/libgta_b200$v.so timeout 200 python tools/bwd2_phase.py ${w%%:*} ${w##*:} | head -5
  done
done
