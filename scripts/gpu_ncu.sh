#!/bin/bash
# one full ncu capture of one kernel (regex $2) of the bench step
TAG=${1:-n}
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:$2 -s ${3:-3} -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; tail -2 gpurun_out/ncu_full_$TAG.log
