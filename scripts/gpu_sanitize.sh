#!/bin/bash
# compute-sanitizer hygiene of the library's kernels on a small workload that touches every
# path (scripts/sanitize_workload.py): memcheck + racecheck with a kernel filter on the iglu namespace; initcheck
# WITHOUT a filter (torch's own fills have to be tracked, or every freshly zeroed tensor reads as uninitialised).
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --kernel-regex kns=iglu python scripts/sanitize_workload.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done" gpurun_out/sanitizer_$tool.log | tail -3
done
timeout 1500 compute-sanitizer --tool initcheck python scripts/sanitize_workload.py > gpurun_out/sanitizer_initcheck.log 2>&1
echo "== initcheck exit $?"; grep -E "ERROR SUMMARY|done" gpurun_out/sanitizer_initcheck.log | tail -3
grep -B2 -A12 "Uninitialized" gpurun_out/sanitizer_initcheck.log | grep -E "Uninitialized|iglu|at 0x|in /" | head -40
