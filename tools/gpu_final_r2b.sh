#!/bin/bash
# Final round-2 artefacts (third session): sanitizer on the forward pipelines (new cp.async staging rows), then gpu_final_r2.sh
mkdir -p gpurun_out
TOOLS="racecheck memcheck" ARGS="two_launch single_launch" ST=300 bash tools/gpu_sanitizer.sh
bash tools/gpu_final_r2.sh
