#!/bin/bash
# One gpurun call: stage times of the cfg2 shape (and others given as extra specs) for every library under gencore_b200/csrc/variants/.
#   gpurun --timeout 900 -- 'bash scripts/gpu_variants.sh <tag> [spec ...]'
tag=${1:-v}; shift
specs=${@:-cfg2}
mkdir -p gpurun_out
for lib in gencore_b200/csrc/variants/lib_*.so; do
  echo "== $lib" | tee -a gpurun_out/variants_$tag.log
  GENCORE_B200_LIB=$PWD/$lib timeout 300 python scripts/shape_perf.py $specs 2>&1 | tail -n +1 | tee -a gpurun_out/variants_$tag.log
done
