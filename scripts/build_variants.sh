#!/bin/bash
# Tuning aid: builds libgencore_b200 with different compile-time knobs of the ring kernel into gencore_b200/csrc/variants/
# (git-ignored); scripts/gpu_variants.sh measures them on the GPU box through GENCORE_B200_LIB.
#   bash scripts/build_variants.sh name1:"-DFLAG=1 ..." name2:"..."
cd "$(dirname "$0")/.."
mkdir -p gencore_b200/csrc/variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC $flags -Xptxas -v -I include \
    -o gencore_b200/csrc/variants/lib_$name.so gencore_b200/csrc/gencore_b200.cu 2> /tmp/variant_$name.log || { echo "$name: build failed"; tail -5 /tmp/variant_$name.log; continue; }
  echo "$name: $(grep -A2 'vote_ring_kernel' /tmp/variant_$name.log | grep -E 'registers' | head -1 | sed 's/ptxas info    : //') $(grep -A1 'vote_ring_kernel' /tmp/variant_$name.log | grep -oE '[0-9]+ bytes stack' | head -1)"
done
