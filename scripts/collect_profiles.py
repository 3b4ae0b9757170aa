#!/usr/bin/env python
"""Copies what a gpu_job.sh run left under gpurun_out/ (scratch) into profiles/ (tracked) as text: the bench line, the launch
list, the ncu details page of every captured kernel, the DRAM traffic of the ring kernel (profiles/traffic.json).
    python scripts/collect_profiles.py <tag> [<prefix, default r03>]"""
import json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
pre = sys.argv[2] if len(sys.argv) > 2 else "r03"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
line = [l for l in open(os.path.join(G, f"bench_{tag}.json")) if l.startswith("{")][-1]
open(os.path.join(P, f"{pre}_bench_n1.json"), "w").write(line)
shutil.copy(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"{pre}_launches.csv"))
for f in (f"shapes_{tag}.log", f"bam_bench_{tag}.json"):
    if os.path.exists(os.path.join(G, f)):
        shutil.copy(os.path.join(G, f), os.path.join(P, f"{pre}_{f.replace('_' + tag, '')}"))
for k in ("ring", "select_template", "slow_columns", "umi_group", "duplex_kernel", "tile_prep2"):
    rep = os.path.join(G, f"prof_{k}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    name = {"ring": "vote_ring", "duplex_kernel": "duplex"}.get(k, k)
    open(os.path.join(P, f"{pre}_{name}_details.txt"), "w").write(det)
    if k == "ring":
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw"], capture_output=True, text=True).stdout
        rd = wr = None
        for l in raw.splitlines():
            f = l.split()
            if len(f) >= 3 and f[0] == "dram__bytes_read.sum":
                rd = float(f[-1].replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[f[1]]
            if len(f) >= 3 and f[0] == "dram__bytes_write.sum":
                wr = float(f[-1].replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[f[1]]
        if rd and wr:
            json.dump({"kernel": "vote_ring_kernel", "traffic": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                       "source": f"profiles/{pre}_vote_ring_details.txt (ncu --set full --clock-control none, one launch, cfg2 1 M pairs)"},
                      open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(sorted(f for f in os.listdir(P) if f.startswith(pre)))
