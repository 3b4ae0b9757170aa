#!/bin/bash
# per-kernel launch list of one bench step on another shape ($1) and vote mode ($2)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file gpurun_out/launches_$1_$2.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --config $1 --vote-mode $2 > gpurun_out/ncu_list_$1_$2.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/launches_$1_$2.csv")))
hdr=None
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if 0 < int(d["ID"]) < 12: print(d["ID"], d["Kernel Name"][:44], d["Metric Value"])
PY
