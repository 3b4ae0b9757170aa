#!/usr/bin/env python
"""Measurement aid: the timeline of gcb_consensus_batch (gcb_set_debug key 6) on the cfg2 shape, for a few chunk sizes.
    python scripts/e2e_trace.py [pairs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gencore_b200 import synth
from gencore_b200.abi import Options
from gencore_b200.device import pinned_copy, pinned_result
from gencore_b200.engine import ConsensusEngine

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cfg = synth.CONFIGS["cfg2"]
batch, genome, _ = synth.make_batch(cfg, seed=5, n_pairs=pairs, with_qnames=False)
with ConsensusEngine(Options.default(), 0) as eng:
    eng.set_reference(genome)
    pb, pr = pinned_copy(batch), pinned_result(batch, len(batch.payload) // 4 + 4096)
    for chunk in (48 << 20, 24 << 20, 12 << 20, 96 << 20):
        eng.set_chunk_bytes(chunk)
        eng.set_debug(6, 0)
        for _ in range(2):
            eng.cluster_by_umi(pb, pr)
        eng.set_debug(6, 1)
        t = time.perf_counter()
        for _ in range(3):
            eng.cluster_by_umi(pb, pr)
        print("chunk %d MB: %.3f ms per call" % (chunk >> 20, (time.perf_counter() - t) / 3 * 1e3), file=sys.stderr, flush=True)
