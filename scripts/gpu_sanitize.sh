#!/bin/bash
# compute-sanitizer over the default engine: small parity cases with every tool, and batches large enough that every CTA of
# the ring kernel goes around its ring (more tiles than 16 stages x 148 SMs) under memcheck and racecheck.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_sanitize.sh <tag>'
tag=${1:-s}
out=gpurun_out/sanitizer_$tag.txt
mkdir -p gpurun_out; : > $out
small="cfg2_1500 ragged_duplex_2 edge_strict cfg3_1500 golden_cfg4_600"
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool python scripts/sanitize_cases.py $small" >> $out
  timeout 600 compute-sanitizer --tool $tool python scripts/sanitize_cases.py $small 2>&1 | grep -E "^ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -20 >> $out
done
big="cfg2:200000 cfg4:60000 cfg5:100000"
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool python scripts/sanitize_cases.py $big" >> $out
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_cases.py $big 2>&1 | grep -E "^ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -20 >> $out
done
cat $out
