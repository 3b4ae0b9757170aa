#!/bin/bash
TAG=${1:-pp}
mkdir -p gpurun_out
echo "== pytest gpu (pipelined only, guarded)"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "pipelined" > gpurun_out/pytest_pipe_$TAG.log 2>&1; tail -5 gpurun_out/pytest_pipe_$TAG.log
for M in 0 1; do
echo "== bench vote-mode $M"; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --vote-mode $M > gpurun_out/bench_${TAG}_m$M.json 2> gpurun_out/bench_${TAG}_m$M.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${TAG}_m$M.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['config']['stage_ms'], d['roofline']['frac'], d['e2e']['value'])
except Exception as e:
    print("bench failed", e); print(open('gpurun_out/bench_${TAG}_m$M.err').read()[-1500:])
PY
done
echo "== ncu full pipe"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vote_pipe -s 3 -c 1 -f -o gpurun_out/prof_pipe_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --vote-mode 1 > gpurun_out/ncu_pipe_$TAG.log 2>&1; tail -2 gpurun_out/ncu_pipe_$TAG.log
