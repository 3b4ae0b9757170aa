#!/usr/bin/env python
"""One sorted BAM over N processes / GPUs: shard i of N runs `gencore_b200 --shard i/N --device <i mod GPUs>` on the
whole input and keeps the clusters of its coordinate window; `gencore_b200 --merge` joins the outputs in the reference's order.
    python scripts/sharded_bam.py -n 8 -i in.bam -o out.bam -r ref.fa [--gpus 8] [--engine lib.so] [reference flags ...]
Prints one JSON line with the wall times.  (Every process still inflates and parses the whole input: the host side of the tool
is what bounds it)"""
import argparse, json, os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gencore_b200 import build as gbuild

ap = argparse.ArgumentParser()
ap.add_argument("-n", type=int, default=2)
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("-i", required=True)
ap.add_argument("-o", required=True)
ap.add_argument("-r", required=True)
ap.add_argument("--engine", default=None)
args, flags = ap.parse_known_args()
cli = gbuild.build_cli()
with tempfile.TemporaryDirectory() as tmp:
    parts = [os.path.join(tmp, f"shard{i}.bam") for i in range(args.n)]
    t0 = time.perf_counter()
    procs = []
    for i, part in enumerate(parts):
        cmd = [cli, "-i", args.i, "-o", part, "-r", args.r, "--shard", f"{i}/{args.n}", "--device", str(i % max(args.gpus, 1))] + flags
        if args.engine:
            cmd += ["--engine", args.engine]
        procs.append(subprocess.Popen(cmd, stderr=subprocess.PIPE, text=True))
    for p in procs:
        err = p.communicate()[1]
        if p.returncode != 0:
            sys.exit("shard failed: " + err[-2000:])
    t1 = time.perf_counter()
    subprocess.run([cli, "-o", args.o, "--merge"] + parts, check=True)
    t2 = time.perf_counter()
print(json.dumps({"shards": args.n, "gpus": args.gpus, "shards_s": t1 - t0, "merge_s": t2 - t1, "total_s": t2 - t0}))
