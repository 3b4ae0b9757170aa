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
for name in sys.argv[1:] or ["cfg2_1500", "ragged_duplex_2", "edge_strict", "cfg3_1500"]:
    batch, genome, opt = cases[name]()
    with ConsensusEngine(opt, 0) as eng:
        eng.set_reference(genome)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, orc.consensus(batch, genome, opt), name)
    print("ok", name, batch.n_pairs, "pairs")
