#!/bin/bash
# bench line + per-kernel launch list (no tests, no full capture)
TAG=${1:-l}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline ${2:+--sweep $2} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value %.4g ms/step %.4f stage_ms %s frac %.3f e2e %.4g" % (d["value"], d["ms_per_step"], {k: round(v,4) for k,v in d["config"]["stage_ms"].items()}, d["roofline"]["frac"], d["e2e"]["value"]))
PY
grep sweep gpurun_out/bench_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/launches_$TAG.csv")))
hdr=None
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if 10 < int(d["ID"]) < 21: print(d["ID"], d["Kernel Name"][:44], d["Metric Value"])
PY
