#!/bin/bash
# One GPU-box visit: smoke, gpu parity tests, bench, ncu launch list, one full ncu capture of the vote kernel.
# usage: gpurun --timeout 2400 -- 'bash scripts/gpu_check.sh [tag]'
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 3000 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -8 gpurun_out/pytest_gpu_$TAG.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1; tail -2 gpurun_out/ncu_bench_$TAG.log
echo "== ncu full (vote_tiled_kernel)"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:vote_tiled -s 3 -c 1 -f -o gpurun_out/vote_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; tail -2 gpurun_out/ncu_full_$TAG.log
echo "== memcheck (smoke under compute-sanitizer)"; timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_$TAG.log 2>&1; tail -4 gpurun_out/memcheck_$TAG.log
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; tail -c 1500 gpurun_out/bench_ref_$TAG.json
ls -la gpurun_out
