#!/bin/bash
# One GPU session: parity tests, bench lines for the 4 config shapes, ncu launch list + one full capture.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | tail -5
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_msn_enc.json 2> gpurun_out/bench_msn_enc.err; tail -c 3000 gpurun_out/bench_msn_enc.json; tail -3 gpurun_out/bench_msn_enc.err
for wl in msn_dec clevr_enc clevr_dec; do
  python bench.py --workload $wl --no-cpu --no-e2e > gpurun_out/bench_$wl.json 2>gpurun_out/bench_$wl.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$wl.json")); print("$wl", d["value"], d["ms_per_step"], d["roofline"])
except Exception as e: print("$wl failed", e)
PY
done
python bench.py --flags 1 --no-cpu --no-e2e > gpurun_out/bench_msn_enc_ptmem.json 2>&1; python -c "
import json; d=json.load(open('gpurun_out/bench_msn_enc_ptmem.json')); print('ptmem', d['value'], d['roofline'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -f -o gpurun_out/prof_attn python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --batch 16 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
