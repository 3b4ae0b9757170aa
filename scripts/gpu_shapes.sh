#!/bin/bash
# exploration: the hot path on the other fixed-length shapes of BASELINE.json (1M pairs each), one line per shape
mkdir -p gpurun_out
for C in cfg1 cfg3 cfg4; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $C > gpurun_out/bench_shape_$C.json 2> gpurun_out/bench_shape_$C.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_shape_$C.json"))
    print("$C value %.4g ms/step %.4f stage_ms %s ring frac %.3f whole vote frac %.3f e2e %.4g" % (d["value"], d["ms_per_step"], {k: round(v,4) for k,v in d["config"]["stage_ms"].items()}, d["roofline"]["frac"], d["roofline"]["whole_vote"]["frac"], d["e2e"]["value"]))
except Exception as e:
    print("$C failed", e); print(open("gpurun_out/bench_shape_$C.err").read()[-800:])
PY
done
