#!/usr/bin/env python
"""Exploration (not a bench line): stage times of the hot path on other shapes, one process, one GPU.
    python scripts/shape_perf.py <spec> [<spec> ...]     spec = cfgN[:pairs[:window_shift]] | deep:<depth>:<err>[:pairs[:window_shift]]
GENCORE_B200_LIBS=a.so,b.so: every shape is measured with each of these builds of the library in turn (same batch, same process)."""
import dataclasses, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gencore_b200 import synth
from gencore_b200.abi import (STAGE_ALL, STAGE_DUPLEX, STAGE_SELECT_TEMPLATE, STAGE_UMI_GROUP, STAGE_VOTE_FAST_ONLY, STAGE_VOTE_PREP_ONLY,
                              STAGE_VOTE_REST_ONLY, Options)
from gencore_b200.device import DeviceBatch, DeviceResult
from gencore_b200.engine import ConsensusEngine

dev = torch.device("cuda:0")
for spec in sys.argv[1:]:
    f = spec.split(":")
    if f[0] == "deep":
        depth, err = float(f[1]), float(f[2])
        pairs = int(f[3]) if len(f) > 3 else 500_000
        shift = int(f[4]) if len(f) > 4 else 0
        cfg = dataclasses.replace(synth.CONFIGS["cfg4"], depth=depth, err=err, n_contigs=2, contig_len=10_000_000)
    else:
        pairs = int(f[1]) if len(f) > 1 else 1_000_000
        shift = int(f[2]) if len(f) > 2 else 0
        cfg = synth.CONFIGS[f[0]]
        cfg = dataclasses.replace(cfg, n_contigs=min(cfg.n_contigs, 2), contig_len=min(cfg.contig_len, 20_000_000))
    batch, genome, _ = synth.make_batch(cfg, seed=77, n_pairs=pairs, with_qnames=False)
    opt = Options.default(cluster_size_req=cfg.supporting_reads)
    for lib_path in (os.environ.get("GENCORE_B200_LIBS", "").split(",") if os.environ.get("GENCORE_B200_LIBS") else [None]):
      if lib_path:
          print("[%s]" % os.path.basename(lib_path), end=" ")
      with ConsensusEngine(opt, 0, lib_path=lib_path) as eng:
          eng.set_reference(genome)
          if shift:
              eng.set_debug(2, shift)
          db = DeviceBatch.from_host(batch, dev)
          dr = DeviceResult.allocate(batch.n_pairs, batch.n_clusters, len(batch.payload), dev)
          ts = torch.cuda.Stream(device=dev)
          stages = [STAGE_UMI_GROUP, STAGE_SELECT_TEMPLATE, STAGE_VOTE_PREP_ONLY, STAGE_VOTE_FAST_ONLY, STAGE_VOTE_REST_ONLY, STAGE_DUPLEX]
          for _ in range(3):
              eng.cluster_by_umi_device(db.struct, dr.struct, STAGE_ALL, ts.cuda_stream)
          torch.cuda.synchronize()
          assert eng.batch_status() == 0
          n = 10
          ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)] for _ in range(n)]
          for k in range(n):
              for q, st in enumerate(stages):
                  ev[k][q].record(ts)
                  eng.cluster_by_umi_device(db.struct, dr.struct, st, ts.cuda_stream)
              ev[k][len(stages)].record(ts)
          torch.cuda.synchronize()
          ms = [float(np.mean([ev[k][q].elapsed_time(ev[k][q + 1]) for k in range(n)])) for q in range(len(stages))]
          # the whole path as a caller issues it: one call per pass, no events in between
          p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
          p0.record(ts)
          for k in range(2 * n):
              eng.cluster_by_umi_device(db.struct, dr.struct, STAGE_ALL, ts.cuda_stream)
          p1.record(ts)
          torch.cuda.synchronize()
          pass_ms = p0.elapsed_time(p1) / (2 * n)
          print("%s: %d clusters, %d pairs, %.0f MB payload, max cluster %d KB; stage ms (umi, select, prep, ring, rest, duplex): %s  total %.3f, one call per pass %.3f -> %.3g pairs/s, "
                "vote %.2f TB/s of payload" % (spec, batch.n_clusters, batch.n_pairs, len(batch.payload) / 1e6, batch.max_cluster_bytes() >> 10,
                                               ["%.3f" % x for x in ms], sum(ms), pass_ms, batch.n_pairs / (pass_ms * 1e-3), len(batch.payload) / (sum(ms[2:5]) * 1e-3) / 1e12),
                flush=True)
    del db, dr
