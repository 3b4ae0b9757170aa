#!/bin/bash
# round-1 second half: parity, sweep of vote modes / thread counts, launch list, one full capture of the staged vote kernel
TAG=${1:-r2a}
SWEEP=${2:-0:256,2:256,2:192,2:128}
KERNEL=${3:-vote_staged}
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== pytest gpu parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
fi
echo "== bench + sweep"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --sweep "$SWEEP" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 2600 gpurun_out/bench_$TAG.json; grep sweep gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1; grep -c gpu__time gpurun_out/launches_$TAG.csv
echo "== ncu full"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:$KERNEL -s 3 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; tail -2 gpurun_out/ncu_full_$TAG.log
fi
