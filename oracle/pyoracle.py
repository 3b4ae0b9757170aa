"""ctypes loaders for the two checkers under oracle/.  TEST INFRASTRUCTURE ONLY.

  * `Oracle`     -> oracle/_build/libgencore_oracle.so  (plain-C restatement, gencore_oracle.c)
  * `Reference`  -> oracle/_ref/libgencore_ref.so       (the reference's own classes + ref_harness.cpp)

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from typing import List, Optional

import numpy as np

from gencore_b200.abi import Batch, BatchStruct, Genome, Options, Result, ResultStruct

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libgencore_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libgencore_ref.so")
REF_BIN = os.path.join(HERE, "_ref", "gencore")


def build(port: bool = True, ref: bool = True) -> None:
    """make -C oracle (the ref target is a no-op when /root/reference is absent)."""
    targets = (["port"] if port else []) + (["ref", "bridge"] if ref else [])  # bridge: the reference bound to the C ABI (integration/)
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


class GenomeStruct(C.Structure):
    _fields_ = [("packed4", C.c_void_p), ("contig_off", C.c_void_p), ("contig_len", C.c_void_p), ("n_contigs", C.c_int32)]


class Oracle:
    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(port=True, ref=False)
        self.lib = C.CDLL(PORT_SO)
        self.lib.gco_consensus_batch.restype = C.c_int
        self.lib.gco_consensus_batch.argtypes = [C.c_void_p] * 4
        self.lib.gco_umi_diff.argtypes = [C.c_char_p, C.c_char_p]
        self.lib.gco_is_duplex.argtypes = [C.c_char_p, C.c_char_p]
        self.lib.gco_get_umi.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        self.lib.gco_encode_umi.argtypes = [C.c_char_p, C.c_void_p, C.c_int]

    def consensus(self, batch: Batch, genome: Optional[Genome], opt: Optional[Options] = None) -> Result:
        opt = opt or Options.default()
        res = Result.allocate(batch)
        bs, rs = batch.as_struct(), res.as_struct()
        gp = None
        if genome is not None:
            gs = GenomeStruct(genome.packed4.ctypes.data, genome.contig_off.ctypes.data, genome.contig_len.ctypes.data,
                              len(genome.contig_len))
            gp = C.byref(gs)
        rc = self.lib.gco_consensus_batch(C.byref(opt), gp, C.byref(bs), C.byref(rs))
        if rc != 0:
            raise RuntimeError(f"gco_consensus_batch -> {rc}")
        return res

    def umi_diff(self, a: str, b: str) -> int:
        return self.lib.gco_umi_diff(a.encode(), b.encode())

    def is_duplex(self, a: str, b: str) -> bool:
        return bool(self.lib.gco_is_duplex(a.encode(), b.encode()))

    def get_umi(self, qname: str, prefix: str) -> str:
        buf = C.create_string_buffer(512)
        n = self.lib.gco_get_umi(qname.encode(), prefix.encode(), buf, 512)
        assert n >= 0
        return buf.value.decode()

    def encode_umi(self, umi: str, words: int) -> np.ndarray:
        out = np.zeros(words, np.uint64)
        n = self.lib.gco_encode_umi(umi.encode(), out.ctypes.data, words)
        assert n >= 0
        return out


REF_PAIR_RESULT = np.dtype(
    [
        ("cluster", "<i4"), ("slot", "<i4", (2,)), ("name_slot", "<i4", (2,)), ("l_qname", "<i4", (2,)),
        ("diff", "<i4", (2,)), ("nm", "<i4", (2,)), ("fr", "<i4", (2,)), ("rr", "<i4", (2,)),
        ("merge_reads", "<i4"), ("reverse_merge_reads", "<i4"), ("is_duplex", "<i4"), ("reserved", "<i4"), ("pad", "<i4"),
        ("out_off", "<i8", (2,)),
    ]
)
assert REF_PAIR_RESULT.itemsize == 96

REF_STATS = np.dtype(
    [
        ("pre_cluster", "<i8"), ("pre_multi_cluster", "<i8"), ("pre_molecule", "<i8"), ("pre_molecule_se", "<i8"),
        ("pre_molecule_pe", "<i8"), ("pre_uncounted", "<i8"), ("pre_hist", "<i8", (100,)),
        ("post_cluster", "<i8"), ("post_multi_cluster", "<i8"), ("post_sscs", "<i8"), ("post_dcs", "<i8"),
    ]
)


def write_fasta(path: str, contigs: List[np.ndarray], names: List[str]) -> None:
    """60-column FASTA without blank lines (fastareader.cpp:70-95 would inject a raw newline otherwise)."""
    with open(path, "wb") as f:
        for name, c in zip(names, contigs):
            f.write(b">" + name.encode() + b"\n")
            n = len(c)
            full = (n // 60) * 60
            if full:
                block = np.empty((full // 60, 61), np.uint8)
                block[:, :60] = c[:full].reshape(-1, 60)
                block[:, 60] = 10
                f.write(block.tobytes())
            if n > full:
                f.write(c[full:].tobytes() + b"\n")


def reference_available() -> bool:
    return os.path.exists(REF_SO)


class Reference:
    """The reference's own Cluster::clusterByUMI over a packed batch (oracle/ref_harness.cpp)."""

    def __init__(self, opt: Optional[Options], umi_prefix: str, genome: Optional[Genome], contigs: Optional[List[np.ndarray]]):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        self.lib = C.CDLL(REF_SO)
        self.lib.gcr_create.restype = C.c_void_p
        self.lib.gcr_create.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.gcr_destroy.argtypes = [C.c_void_p]
        self.lib.gcr_consensus_batch.restype = C.c_int64
        self.lib.gcr_consensus_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.gcr_umi_diff.argtypes = [C.c_char_p, C.c_char_p]
        self.lib.gcr_is_duplex.argtypes = [C.c_char_p, C.c_char_p]
        self.lib.gcr_get_umi.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        self.opt = opt or Options.default()
        self._tmp = None
        fasta = b""
        names: List[str] = []
        lens = np.zeros(0, np.int64)
        if genome is not None and contigs is not None:
            self._tmp = tempfile.NamedTemporaryFile(suffix=".fa", delete=False)
            self._tmp.close()
            write_fasta(self._tmp.name, contigs, genome.names)
            fasta = self._tmp.name.encode()
            names = genome.names
            lens = np.ascontiguousarray(genome.contig_len, np.int64)
        arr = (C.c_char_p * max(len(names), 1))(*[n.encode() for n in names])
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(2)
        if not os.environ.get("GCR_VERBOSE"): os.dup2(devnull, 2)  # the FASTA loader prints the contig list to stderr (fastareader.cpp:161,168)
        try:
            self.ctx = self.lib.gcr_create(C.byref(self.opt), umi_prefix.encode(), fasta, len(names), arr,
                                           lens.ctypes.data if len(lens) else None)
        finally:
            os.dup2(saved, 2)
            os.close(saved)
            os.close(devnull)

    def close(self):
        if self.ctx:
            self.lib.gcr_destroy(self.ctx)
            self.ctx = None
        if self._tmp is not None:
            try:
                os.unlink(self._tmp.name)
            except OSError:
                pass
            self._tmp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def consensus(self, batch: Batch, want_results: bool = True):
        """Returns (pair results, out payload, stats, seconds inside clusterByUMI)."""
        assert batch.qnames is not None, "the reference harness needs qnames"
        names = [bytes(q) for q in batch.qnames]
        blob = b"\0".join(names) + b"\0"
        lens = np.fromiter((len(q) + 1 for q in names), np.int64, len(names))
        off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64) if len(names) else np.zeros(0, np.int64)
        cap = batch.n_pairs if want_results else 0
        res = np.zeros(max(cap, 1), REF_PAIR_RESULT)
        out = np.zeros(len(batch.payload) if want_results else 16, np.uint8)
        out_bytes = np.zeros(1, np.int64)
        stats = np.zeros(1, REF_STATS)
        secs = C.c_double(0)
        bs = batch.as_struct()
        nm = batch.nm if batch.nm is not None else np.zeros(2 * batch.n_pairs, np.uint8)
        n = self.lib.gcr_consensus_batch(self.ctx, C.byref(bs), blob, off.ctypes.data, nm.ctypes.data,
                                         res.ctypes.data if want_results else None, cap,
                                         out.ctypes.data if want_results else None, len(out) if want_results else 0,
                                         out_bytes.ctypes.data, stats.ctypes.data, C.byref(secs))
        if n < 0:
            raise RuntimeError(f"gcr_consensus_batch -> {n}")
        return res[:min(n, cap)], out[:int(out_bytes[0])], stats[0], secs.value, int(n)

    def umi_diff(self, a, b):
        return self.lib.gcr_umi_diff(a.encode(), b.encode())

    def is_duplex(self, a, b):
        return bool(self.lib.gcr_is_duplex(a.encode(), b.encode()))

    def get_umi(self, qname, prefix):
        buf = C.create_string_buffer(512)
        n = self.lib.gcr_get_umi(qname.encode(), prefix.encode(), buf, 512)
        assert n >= 0
        return buf.value.decode()


def reference_fasta_load(text: bytes, max_contigs: int = 64):
    """The reference's own FastaReader over `text`: (ids, sizes, offsets, packed bytes) in file order."""
    lib = C.CDLL(REF_SO)
    lib.gcr_fasta_load.restype = C.c_int
    lib.gcr_fasta_load.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    ids = np.zeros(256 * max_contigs, np.uint8)
    sizes, offs = np.zeros(max_contigs, np.int64), np.zeros(max_contigs, np.int64)
    cap = len(text) // 2 + 16 * max_contigs + 16
    packed = np.zeros(cap, np.uint8)
    with tempfile.NamedTemporaryFile(suffix=".fa", delete=False) as f:
        f.write(text)
        path = f.name
    try:
        n = lib.gcr_fasta_load(path.encode(), max_contigs, ids.ctypes.data, sizes.ctypes.data, offs.ctypes.data, packed.ctypes.data, cap)
    finally:
        os.unlink(path)
    if n < 0:
        raise RuntimeError("gcr_fasta_load failed")
    names = [bytes(ids[256 * i:256 * (i + 1)]).split(b"\0")[0] for i in range(n)]
    return names, sizes[:n].copy(), offs[:n].copy(), packed
