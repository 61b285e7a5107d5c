#!/bin/bash
# compute-sanitizer over a small pass through every kernel (scripts/sanitize_workload.py checks every result against the
# oracle as well).  memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards between the warps and
# lanes that share scratch; synccheck: invalid __syncwarp / barrier use; initcheck: reads of uninitialised device memory.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  n=3000; [ $tool = memcheck ] && n=6000
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python scripts/sanitize_workload.py $n > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
  grep -E "ok$" gpurun_out/sanitize_$tool.log | tr '\n' ' '; echo
done
