#!/bin/bash
TAG=${1:-fin}
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/pytest_gpu_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bam bench"; timeout 900 python scripts/bam_bench.py 300000 > gpurun_out/bam_bench_$TAG.json 2> gpurun_out/bam_bench_$TAG.err; tail -c 800 gpurun_out/bam_bench_$TAG.json; tail -3 gpurun_out/bam_bench_$TAG.err
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 2600 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
