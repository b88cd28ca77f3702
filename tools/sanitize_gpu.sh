#!/bin/bash
# compute-sanitizer over tools/sanitize_batch.py (every kernel of libdynprog_cuda, every dpc_solve path); logs -> $1
out=${1:-gpurun_out/sanitizer}
mkdir -p $out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 77 python tools/sanitize_batch.py > $out/r2_sanitizer_$tool.log 2>&1
  echo "compute-sanitizer --tool $tool: exit code $?" >> $out/r2_sanitizer_$tool.log
  tail -4 $out/r2_sanitizer_$tool.log
done
