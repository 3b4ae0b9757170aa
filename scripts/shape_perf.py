#!/usr/bin/env python
"""Exploration: stage times of the hot path on a fixed-length duplex shape of a given depth and error rate (not a bench line).
    python scripts/shape_perf.py <depth> <err> [pairs]"""
import dataclasses, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gencore_b200 import synth
from gencore_b200.abi import (STAGE_ALL, STAGE_DUPLEX, STAGE_SELECT_TEMPLATE, STAGE_UMI_GROUP, STAGE_VOTE_FAST_ONLY, STAGE_VOTE_PREP_ONLY,
                              STAGE_VOTE_REST_ONLY, Options)
from gencore_b200.device import DeviceBatch, DeviceResult
from gencore_b200.engine import ConsensusEngine

depth, err = float(sys.argv[1]), float(sys.argv[2])
pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 500_000
cfg = dataclasses.replace(synth.CONFIGS["cfg4"], depth=depth, err=err, n_contigs=2, contig_len=10_000_000)
batch, genome, _ = synth.make_fixed_batch(cfg, seed=77, n_pairs=pairs, with_qnames=False)
print("depth %g err %g: %d clusters, %d pairs, %.0f MB payload" % (depth, err, batch.n_clusters, batch.n_pairs, len(batch.payload) / 1e6))
dev = torch.device("cuda:0")
for mode in (4, 0):
    with ConsensusEngine(Options.default(), 0) as eng:
        eng.set_reference(genome)
        eng.set_vote_mode(mode)
        db = DeviceBatch.from_host(batch, dev)
        dr = DeviceResult.allocate(batch.n_pairs, batch.n_clusters, len(batch.payload), dev)
        ts = torch.cuda.Stream(device=dev)
        stages = [STAGE_UMI_GROUP, STAGE_SELECT_TEMPLATE, STAGE_VOTE_PREP_ONLY, STAGE_VOTE_FAST_ONLY, STAGE_VOTE_REST_ONLY, STAGE_DUPLEX]
        for _ in range(3):
            eng.cluster_by_umi_device(db.struct, dr.struct, STAGE_ALL, ts.cuda_stream)
        torch.cuda.synchronize()
        n = 5
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)] for _ in range(n)]
        for k in range(n):
            for q, st in enumerate(stages):
                ev[k][q].record(ts)
                eng.cluster_by_umi_device(db.struct, dr.struct, st, ts.cuda_stream)
            ev[k][len(stages)].record(ts)
        torch.cuda.synchronize()
        ms = [float(np.mean([ev[k][q].elapsed_time(ev[k][q + 1]) for k in range(n)])) for q in range(len(stages))]
        print("mode %d stage ms (umi, select, prep, vote, rest, duplex): %s  total %.3f  -> %.3g pairs/s" %
              (mode, ["%.3f" % x for x in ms], sum(ms), batch.n_pairs / (sum(ms) * 1e-3)))
