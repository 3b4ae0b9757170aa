"""Writes tests/golden_fasta/fasta.npz: what the UNMODIFIED reference's FastaReader (oracle/_ref, built from /root/reference)
makes of every text of tests/fasta_cases.py.  Run in the build container:  python tests/make_golden_fasta.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from fasta_cases import fasta_cases  # noqa: E402
from oracle.pyoracle import reference_fasta_load  # noqa: E402

out = {}
for name, text in fasta_cases().items():
    ids, sizes, offs, packed = reference_fasta_load(text, max_contigs=128)
    out[name + "/text"] = np.frombuffer(text, np.uint8)
    out[name + "/ids"] = np.asarray(ids, dtype="S") if ids else np.zeros(0, "S1")
    out[name + "/sizes"] = sizes
    out[name + "/offs"] = offs
    out[name + "/packed"] = packed[:int(offs[-1] + (sizes[-1] + 1) // 2) if len(sizes) else 0]
np.savez_compressed(os.path.join(HERE, "golden_fasta", "fasta.npz"), **out)
print("wrote", len(fasta_cases()), "cases")
