#!/bin/bash
# compute-sanitizer hygiene: memcheck + racecheck of the library's kernels on a small workload
# that touches every path (scripts/sanitize_workload.py). (initcheck is not run: torch's own fills are not tracked
# under a kernel filter and show up as false positives.)
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --kernel-regex kns=iglu python scripts/sanitize_workload.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|done" gpurun_out/sanitizer_$tool.log | tail -3
done
