#!/bin/bash
# launch list of the default command, every launch of the run (no skip), + BASELINE config 1 line again
mkdir -p gpurun_out/art3
O=gpurun_out/art3
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_msn_enc.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-info --no-backward > /dev/null 2>&1
cut -d, -f5 $O/launches_msn_enc.csv | sort | uniq -c | sort -rn | head -12
timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu --no-e2e > $O/bench_cfg1.json 2> $O/bench_cfg1.err
timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu --no-e2e > $O/bench_cfg1_b.json 2> $O/bench_cfg1.err
python - <<'PY'
import json
for f in ("gpurun_out/art3/bench_cfg1.json","gpurun_out/art3/bench_cfg1_b.json"):
    d=json.load(open(f)); print(f, round(d["value"],2), round(d["ms_per_step"],4), (d.get("backward") or {}).get("ms"))
PY
