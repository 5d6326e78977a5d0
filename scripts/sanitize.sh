#!/bin/bash
# compute-sanitizer (memcheck, racecheck, initcheck, synccheck) over scripts/sanitize_target.py; summaries -> gpurun_out/sanitizer_<tool>.txt
# Usage: bash scripts/sanitize.sh [tools...]     (default: all four, run concurrently)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS="${*:-memcheck racecheck initcheck synccheck}"
for t in $TOOLS; do
  (
    timeout 900 compute-sanitizer --tool $t --print-limit 20 --error-exitcode 77 \
      python scripts/sanitize_target.py all > gpurun_out/sanitizer_$t.full.log 2>&1
    rc=$?
    { echo "== compute-sanitizer --tool $t python scripts/sanitize_target.py all  (exit $rc; 77 = errors reported, 124 = timeout)";
      grep -E "ok:|ok$|ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard|Invalid|Uninitialized" gpurun_out/sanitizer_$t.full.log | head -60; } > gpurun_out/sanitizer_$t.txt
  ) &
done
wait
for t in $TOOLS; do tail -n 8 gpurun_out/sanitizer_$t.txt; done
