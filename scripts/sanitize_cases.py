#!/usr/bin/env python
"""A few parity cases through the default engine, for compute-sanitizer (scripts/gpu_sanitize.sh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_cases
from parity import assert_results_equal
from gencore_b200.engine import ConsensusEngine
from oracle.pyoracle import Oracle

cases = dict(parity_cases.gpu_cases())
orc = Oracle()
def big(name):
    """cfgN:<pairs>: a batch with more tiles than ring stages per CTA x SMs (every CTA goes around its ring and its arena)."""
    import dataclasses
    from gencore_b200 import synth
    from gencore_b200.abi import Options
    cfg_name, pairs = name.split(":")
    cfg = synth.CONFIGS[cfg_name]
    cfg = dataclasses.replace(cfg, n_contigs=min(cfg.n_contigs, 2), contig_len=min(cfg.contig_len, 20_000_000))
    batch, genome, _ = synth.make_batch(cfg, seed=4242, n_pairs=int(pairs), with_qnames=False)
    return batch, genome, Options.default(cluster_size_req=cfg.supporting_reads)


for name in sys.argv[1:] or ["cfg2_1500", "ragged_duplex_2", "edge_strict", "cfg3_1500"]:
    batch, genome, opt = big(name) if ":" in name else cases[name]()
    with ConsensusEngine(opt, 0) as eng:
        eng.set_reference(genome)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, orc.consensus(batch, genome, opt), name)
    print("ok", name, batch.n_pairs, "pairs")
