#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_reference_modules.py -m gpu -q -k "backward or train or module" 2>&1 | tail -3
for w in "msn_enc 64" "clevr_dec 32" "cfg1 2"; do
  set -- $w
  GTA_BWD_FLAGS=0 timeout 200 python tools/bwd_bench.py $1 $2 | cut -c1-130
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
