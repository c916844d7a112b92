#!/bin/bash
# compute-sanitizer over every kernel path (tools/sanitize_all.py); logs -> gpurun_out/
export HY_CUDA_JIT_CACHE=$PWD/gpurun_out/jit_cache
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_all.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?"; tail -4 gpurun_out/r02_sanitizer_$tool.log
done
