#!/bin/bash
# host-staged pipeline with overlap across calls: parity test + default bench line (e2e leg)
mkdir -p gpurun_out/art3
timeout 300 python -m pytest tests -m gpu -q -x -k "host_pipeline or headline or golden" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/art3/bench_msn_enc.json 2> gpurun_out/art3/bench_msn_enc.err; echo "default exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/art3/bench_msn_enc.json")); print(round(d["value"],2), round(d["ms_per_step"],4), d["e2e"], d["clocks"])
PY
