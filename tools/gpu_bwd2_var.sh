#!/bin/bash
mkdir -p gpurun_out
GTA_B200_LIB=$PWD/gta_b200/libgta_b200_dbg.so timeout 250 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "fused_kernel_matches_kernel_pair" > gpurun_out/bwd2_dbg.log 2>&1
grep -E "passed|failed|error" gpurun_out/bwd2_dbg.log | tail -3
grep "timed out" gpurun_out/bwd2_dbg.log | awk '{print "block",$7,"warp",int($9/32),"bar",$10,"parity",$12}' | sort | uniq -c | sort -k3n -k5n | head -40
if grep -q "failed\|error\|timed out" gpurun_out/bwd2_dbg.log; then grep -E "Error|assert" gpurun_out/bwd2_dbg.log | head; exit 1; fi
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "backward" 2>&1 | tail -2
for v in ${VARIANTS:-""}; do
  [ "$v" = "main" ] && v=""
  echo "== variant libgta_b200$v.so"
  for w in ${WL:-msn_enc:64 clevr_dec:32}; do
    GTA_B200_LIB=$PWD/gta_b200/libgta_b200$v.so timeout 200 python tools/bwd_bench.py ${w%%:*} ${w##*:} | cut -c1-110
    GTA_B200_LIB=$PWD/gta_b200/libgta_b200$v.so timeout 200 python tools/bwd2_phase.py ${w%%:*} ${w##*:}
  done
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bwd|rotate|delta" -c 40 --csv --log-file gpurun_out/launches_bwd2.csv python tools/bwd_bench.py msn_enc 64 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [l for l in open('gpurun_out/launches_bwd2.csv') if l.startswith('"')]
r = list(csv.reader(rows)); h = r[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
d = collections.defaultdict(list)
for x in r[1:]:
    try: d[x[ki][:70]].append(float(x[vi].replace(',', '')))
    except Exception: pass
for k, v in d.items(): print(f"{k:72s} n={len(v):3d} avg {sum(v)/len(v)/1e3:9.1f} us")
PY
