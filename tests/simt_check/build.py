"""Builds the kernels a second time with g++ against the SIMT interpreter (simt_check.h) so the parity
tests can execute the real kernel source on a CPU-only box.  TEST INFRASTRUCTURE ONLY: the output
(tests/simt_check/_build/libgencore_b200_simt.so) is loaded by tests/ alone."""
from __future__ import annotations

import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "gencore_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libgencore_b200_simt.so")


def build(force: bool = False) -> str:
    deps = glob.glob(os.path.join(CSRC, "*.cu*")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        [os.path.join(HERE, "simt_check.h"), os.path.join(ROOT, "include", "gencore_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-DGCB_SIMT_CHECK", "-x", "c++", "-I", os.path.join(ROOT, "include"),
           "-I", HERE, "-I", CSRC, "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-variable", "-Wno-maybe-uninitialized",
           "-o", OUT, os.path.join(CSRC, "gencore_b200.cu")]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("simt_check build failed:\n" + proc.stdout + proc.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
