#!/bin/bash
# compute-sanitizer over one small pass through every kernel family (tools/gpu_sanity.py); logs -> gpurun_out/
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python tools/gpu_sanity.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2_sanitizer_$tool.log
  tail -4 gpurun_out/r2_sanitizer_$tool.log
done
