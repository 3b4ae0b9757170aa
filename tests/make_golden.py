"""Regenerates tests/golden/*.npz: packed input batches together with what the REFERENCE ITSELF
(oracle/_ref/libgencore_ref.so = /root/reference/src compiled unchanged + oracle/ref_harness.cpp)
returned for them.  Run in the build container (needs /root/reference):  python tests/make_golden.py
The fixtures let the oracle and the CUDA path be pinned to the reference on machines without it."""
import dataclasses
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import cases  # noqa: E402
from gencore_b200 import synth  # noqa: E402
from gencore_b200.abi import Options  # noqa: E402
from oracle import pyoracle  # noqa: E402
from parity import save_golden  # noqa: E402


def main():
    pyoracle.build()
    out = os.path.join(HERE, "golden")
    os.makedirs(out, exist_ok=True)
    todo = []
    b, g, c = cases.edge_batch()
    for name in ("default", "strict", "loose", "s2_x"):
        todo.append((f"edge_{name}", b, g, c, cases.OPTION_SETS[name]))
    for i, umi in enumerate(["none", "single", "duplex"]):
        b, g, c = synth.make_ragged_batch(900 + i, n_clusters=40, umi=umi, contig_len=30_000, err=0.02)
        todo.append((f"ragged_{umi}", b, g, c, Options.default()))
    for name in ("cfg1", "cfg2", "cfg3", "cfg4"):
        cfg = dataclasses.replace(synth.CONFIGS[name], contig_len=30_000, n_contigs=min(synth.CONFIGS[name].n_contigs, 2))
        b, g, c = synth.make_fixed_batch(cfg, seed=20261017, n_pairs=600)
        todo.append((f"{name}_600", b, g, c, Options.default(cluster_size_req=cfg.supporting_reads)))
    for name, b, g, c, opt in todo:
        ref = pyoracle.Reference(opt, b.umi_prefix, g, c)
        pairs, outp, stats, _secs, n = ref.consensus(b)
        ref.close()
        path = os.path.join(out, name + ".npz")
        save_golden(path, b, g, c, opt, pairs, outp, stats, n)
        print(f"{name}: {b.n_pairs} pairs, {b.n_clusters} clusters, {n} reference pairs, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
