"""GPU-side debugging aid: run named parity cases several times and report where they differ from the oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity_cases
from gencore_b200.engine import ConsensusEngine
from gencore_b200.hoststats import group_slots
from oracle.pyoracle import Oracle

names = sys.argv[1:] or ["golden_cfg1_600"]
cases = dict(parity_cases.gpu_cases())
orc = Oracle()
for name in names:
    batch, genome, opt = cases[name]()
    want = orc.consensus(batch, genome, opt)
    for rep in range(3):
        with ConsensusEngine(opt, 0) as eng:
            eng.set_reference(genome)
            res = eng.cluster_by_umi(batch)
            res2 = eng.cluster_by_umi(batch)
        for tag, r in (("first", res), ("second", res2)):
            slots = group_slots(batch, r)
            bad = [n for n in r.groups.dtype.names if not np.array_equal(r.groups[slots][n], want.groups[slots][n])]
            n = int(want.out_bytes[0])
            nb = int((r.out_payload[:n] != want.out_payload[:n]).sum())
            where = np.flatnonzero(r.out_payload[:n] != want.out_payload[:n])[:6]
            print(name, rep, tag, "bad fields", bad, "bad bytes", nb, where, flush=True)
