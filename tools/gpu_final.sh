#!/bin/bash
# Round deliverables: parity suite, smoke, bench lines (all shapes), reference arm, backward timing, ncu launch list + full captures.
mkdir -p gpurun_out
set -o pipefail
timeout 150 python __graft_entry__.py smoke 2>&1 | tail -2 || { echo 'SMOKE FAILED - stopping'; exit 1; }
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --backward > gpurun_out/bench_msn_enc.json 2> gpurun_out/bench_msn_enc.err; tail -c 1200 gpurun_out/bench_msn_enc.json; tail -2 gpurun_out/bench_msn_enc.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err; cut -c1-200 gpurun_out/bench_reference.json
for wl in msn_dec clevr_enc clevr_dec cfg1 sweep2 sweep5 sweep10 sweep20; do
  steps=20; [ $wl = sweep10 ] && steps=5; [ $wl = sweep20 ] && steps=3
  timeout 500 python bench.py --workload $wl --no-cpu --no-e2e --backward --steps $steps > gpurun_out/bench_$wl.json 2>gpurun_out/bench_$wl.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$wl.json")); r=d["roofline"]; print("$wl", round(d["value"],1), "Mtok/s step_ms", round(d["ms_per_step"],3), "attn_ms", round(r["kernel_ms"],3), "stage_ms", round(r["stage_kernel_ms"],3), "frac", round(r["frac"],3), "bwd_ms", round(d["backward"]["ms"],3), d["clocks"])
except Exception as e: print("$wl failed", e); print(open("gpurun_out/bench_$wl.err").read()[-800:])
PY
done
timeout 200 python tools/bwd_phase_timing.py msn_enc 16 > gpurun_out/bwd_phase_msn_enc.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd3 -s 3 -c 1 -f -o gpurun_out/prof_attn_final python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 2 -c 2 -f -o gpurun_out/prof_bwd python tools/bwd_bench.py msn_enc 16 > gpurun_out/ncu_full_bwd.log 2>&1
ls -la gpurun_out | tail -6
