#!/bin/bash
# compute-sanitizer over the default engine (ring vote, slow-column queue) on a few parity cases
mkdir -p gpurun_out
for TOOL in memcheck racecheck synccheck; do
  echo "== $TOOL"
  timeout 600 compute-sanitizer --tool $TOOL --print-limit 5 python scripts/sanitize_cases.py cfg2_1500 ragged_duplex_2 edge_strict cfg3_1500 > gpurun_out/sanitize_$TOOL.log 2>&1
  grep -E "ERROR SUMMARY|^ok|RACECHECK SUMMARY|Error" gpurun_out/sanitize_$TOOL.log | head -12
done
