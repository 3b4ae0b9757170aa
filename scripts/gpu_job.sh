#!/bin/bash
# One gpurun call of round 2: parity tests, the bench line, other shapes, a launch list, ncu captures of the main kernels, the BAM-to-BAM tools.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_job.sh <tag> [tests|notests] [full]'
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.txt 2>&1
if [ "${2:-tests}" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
  tail -5 gpurun_out/pytest_$tag.log
fi
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -c 3000 gpurun_out/bench_$tag.json
timeout 900 python scripts/shape_perf.py cfg1 cfg3 cfg4 cfg5 cfg2:1000000:14 deep:30:0.01 deep:30:0.003 deep:50:0.003 > gpurun_out/shapes_$tag.log 2>&1
cat gpurun_out/shapes_$tag.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 3 --inner 1 --no-cpu-baseline --no-configs --no-strong --no-bam > gpurun_out/launches_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vote_ring -s 3 -c 1 -o gpurun_out/prof_ring_$tag \
  python bench.py --steps 1 --warmup 3 --inner 1 --no-cpu-baseline --no-configs --no-strong --no-bam > gpurun_out/ncu_ring_$tag.log 2>&1
if [ "${3:-}" = "full" ]; then
  for k in select_template slow_columns umi_group duplex_kernel tile_prep2; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/prof_${k}_$tag \
      python bench.py --steps 1 --warmup 3 --inner 1 --no-cpu-baseline --no-configs --no-strong --no-bam > gpurun_out/ncu_${k}_$tag.log 2>&1
  done
  timeout 600 python scripts/bam_bench.py 300000 > gpurun_out/bam_bench_$tag.json 2> gpurun_out/bam_bench_$tag.err
  cat gpurun_out/bam_bench_$tag.json
fi
ls -la gpurun_out | tail -5
