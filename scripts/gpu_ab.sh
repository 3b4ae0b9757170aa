#!/bin/bash
# One gpurun call: A/B of the library builds under gencore_b200/csrc/variants/ on the cfg2 / cfg3 / cfg5 shapes (same batches, one
# process: scripts/shape_perf.py with GENCORE_B200_LIBS), then the GPU tests and the bench line with the build that won
# (2 x log of the cfg2 pass + log of the cfg3 pass + log of the cfg5 pass, smallest wins).
#   gpurun --timeout 640 -- 'bash scripts/gpu_ab.sh <tag>'
tag=${1:-ab}
mkdir -p gpurun_out
libs=$(ls $PWD/gencore_b200/csrc/variants/lib_*.so | paste -sd, -)
GENCORE_B200_LIBS=$libs timeout 240 python scripts/shape_perf.py cfg2 cfg3 cfg5 > gpurun_out/ab_$tag.log 2>&1
cat gpurun_out/ab_$tag.log
win=$(python - <<'P' gpurun_out/ab_$tag.log
import math, re, sys, collections
score = collections.defaultdict(float); seen = collections.defaultdict(set)
for line in open(sys.argv[1]):
    m = re.match(r"\[(lib_\w+\.so)\] (cfg\d): .*one call per pass ([0-9.]+)", line)
    if m:
        score[m.group(1)] += (2.0 if m.group(2) == "cfg2" else 1.0) * math.log(float(m.group(3)))
        seen[m.group(1)].add(m.group(2))
full = [k for k in score if len(seen[k]) == 3]
print(min(full, key=lambda k: score[k]) if full else "")
P
)
echo "winner: $win" | tee -a gpurun_out/ab_$tag.log
if [ -n "$win" ]; then export GENCORE_B200_LIB=$PWD/gencore_b200/csrc/variants/$win; fi
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -c 600 gpurun_out/bench_$tag.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 3 --inner 1 --no-cpu-baseline --no-configs --no-strong --no-bam > gpurun_out/launches_$tag.log 2>&1
