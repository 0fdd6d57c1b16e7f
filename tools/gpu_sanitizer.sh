#!/bin/bash
mkdir -p gpurun_out
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  echo "=== compute-sanitizer --tool $tool"
  timeout ${ST:-400} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_check.py $ARGS > gpurun_out/sanitizer_$tool.log 2>&1; echo "exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|max \|out|backward ok|Error|error" gpurun_out/sanitizer_$tool.log | sort | uniq -c | sort -rn | head -30
done
