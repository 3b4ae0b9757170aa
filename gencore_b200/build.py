"""Builds gencore_b200/csrc/libgencore_b200.so for sm_100a with nvcc (cross-compiles without a GPU).

    python -m gencore_b200.build [--verbose]

The library is kept in-tree (git-ignored) so that it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libgencore_b200.so")
SOURCES = ["gencore_b200.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "gencore_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", os.path.join(ROOT, "include"), "-o", LIB] + \
        [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    return LIB


CLI_SRC = os.path.join(CSRC, "host", "gencore_b200_cli.cpp")
CLI_BIN = os.path.join(HERE, "bin", "gencore_b200")


def build_cli(force: bool = False) -> str:
    """The host pipeline (sorted BAM in, consensus BAM out) that drives the engine through its C ABI."""
    deps = [CLI_SRC, os.path.join(ROOT, "include", "gencore_b200.h")]
    if not force and os.path.exists(CLI_BIN) and all(os.path.getmtime(d) <= os.path.getmtime(CLI_BIN) for d in deps):
        return CLI_BIN
    os.makedirs(os.path.dirname(CLI_BIN), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", CLI_BIN, CLI_SRC, "-lz", "-ldl", "-lpthread"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("g++ failed:\n" + proc.stdout + proc.stderr)
    return CLI_BIN


if __name__ == "__main__":
    print(build_cli(force=True))
    print(build(force=True, verbose="--verbose" in sys.argv))
