#!/bin/bash
# tuning aid: vote stage times of several settings (mode:threads[:ablate[:window_shift]]) in one process
TAG=${1:-sw}
mkdir -p gpurun_out
GCB_PROFILING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sweep-only --sweep "$2" > gpurun_out/sweep_$TAG.out 2> gpurun_out/sweep_$TAG.err; grep sweep gpurun_out/sweep_$TAG.err; tail -3 gpurun_out/sweep_$TAG.err | grep -v sweep
