#!/bin/bash
# Final round-2 artefacts (third session, final code): smoke, whole GPU suite, bench lines, reference arm, launch list, ncu capture
mkdir -p gpurun_out/art3
O=gpurun_out/art3
rm -f $O/*
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q -rA > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_full.log; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_full.log | head -20
timeout 900 python bench.py > $O/bench_msn_enc.json 2> $O/bench_msn_enc.err; echo "default exit $?"
for wl in msn_dec clevr_enc clevr_dec cfg1 sweep2; do timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-e2e > $O/bench_$wl.json 2> $O/bench_$wl.err; echo "$wl exit $?"; done
timeout 600 python bench.py --steps 20 --warmup 5 --flags 32 --no-cpu --no-e2e --no-info > $O/bench_msn_enc_single_launch.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>/dev/null; echo "reference exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file $O/launches_msn_enc.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-info --no-backward > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd3 -s 4 -c 1 -f -o $O/ncu_attn_fwd3_msn_enc python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-info --no-backward > /dev/null 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/art3/bench_*.json")):
    try:
        d=json.load(open(f)); r=d.get("roofline",{}); b=d.get("backward") or {}
        print(f.split("/")[-1], round(d["value"],2), d["unit"], "ms", round(d["ms_per_step"],4), "frac", r.get("frac") and round(r["frac"],3), "step_frac", r.get("step_frac") and round(r["step_frac"],3), "bwd_ms", b.get("ms") and round(b["ms"],3), "launches", d.get("gpu_launches"), d.get("clocks",{}).get("sm_mhz"), d.get("clocks",{}).get("reasons"))
    except Exception as e: print(f, "failed", e)
PY
