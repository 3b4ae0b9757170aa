#!/bin/bash
TAG=${1:-p}
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 1200 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
for K in select_template umi_group; do
echo "== ncu $K"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_${K}_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${K}_$TAG.log 2>&1; tail -1 gpurun_out/ncu_${K}_$TAG.log
done
