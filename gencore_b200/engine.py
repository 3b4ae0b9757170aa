"""Host-side mirror of the reference's seam for the hot path.

The reference's in-process seam is `Cluster::clusterByUMI(umiDiffThreshold, preStats, postStats,
crossContig)` called cluster by cluster from gencore.cpp:355 and gencore.cpp:409.  `ConsensusEngine`
is that call batched: `cluster_by_umi(batch)` hands a packed batch of clusters to the CUDA library
(libgencore_b200.so, C ABI in include/gencore_b200.h) and returns what the reference's Pair objects
would hold.  There is no CPU path: without the compiled library or without an sm_100 device the
constructor raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .abi import (GCB_ABI_VERSION, GCB_ERR_NO_DEVICE, GCB_OK, STAGE_ALL, Batch, BatchStruct, Genome, Options, Result,
                  ResultStruct)

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "csrc", "libgencore_b200.so")

# every symbol include/gencore_b200.h declares
ABI_SYMBOLS = ["gcb_abi_version", "gcb_default_options", "gcb_create", "gcb_destroy", "gcb_last_error", "gcb_set_reference",
               "gcb_set_reference_device", "gcb_consensus_batch", "gcb_consensus_batch_device", "gcb_batch_status",
               "gcb_launch_count", "gcb_set_chunk_bytes", "gcb_extract_umi", "gcb_set_debug", "gcb_set_slow_queue_bytes", "gcb_get_cluster_stats", "gcb_stat_depth", "gcb_pack_fasta", "gcb_host_alloc", "gcb_host_free"]


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gencore_b200: status {code}: {msg}")
        self.code = code


def load_library(path: Optional[str] = None) -> C.CDLL:
    path = path or os.environ.get("GENCORE_B200_LIB") or DEFAULT_LIB  # (the variable: measuring a differently built library)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is missing: build it with `python -m gencore_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.gcb_abi_version.restype = C.c_int
    lib.gcb_default_options.argtypes = [C.c_void_p]
    lib.gcb_create.restype = C.c_int
    lib.gcb_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.gcb_destroy.argtypes = [C.c_void_p]
    lib.gcb_last_error.restype = C.c_char_p
    lib.gcb_last_error.argtypes = [C.c_void_p]
    lib.gcb_set_reference.restype = C.c_int
    lib.gcb_set_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]
    lib.gcb_set_reference_device.restype = C.c_int
    lib.gcb_set_reference_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]
    lib.gcb_consensus_batch.restype = C.c_int
    lib.gcb_consensus_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gcb_consensus_batch_device.restype = C.c_int
    lib.gcb_consensus_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.gcb_batch_status.restype = C.c_int
    lib.gcb_batch_status.argtypes = [C.c_void_p, C.c_void_p]
    lib.gcb_extract_umi.restype = C.c_int
    lib.gcb_extract_umi.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_char_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.gcb_set_slow_queue_bytes.restype = C.c_int
    lib.gcb_set_slow_queue_bytes.argtypes = [C.c_void_p, C.c_int64]
    lib.gcb_get_cluster_stats.restype = C.c_int
    lib.gcb_get_cluster_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.gcb_stat_depth.restype = C.c_int
    lib.gcb_stat_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
    lib.gcb_pack_fasta.restype = C.c_int
    lib.gcb_pack_fasta.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]
    lib.gcb_set_chunk_bytes.restype = C.c_int
    lib.gcb_host_alloc.restype = C.c_void_p
    lib.gcb_host_alloc.argtypes = [C.c_size_t, C.c_int]
    lib.gcb_host_free.restype = None
    lib.gcb_host_free.argtypes = [C.c_void_p]
    lib.gcb_set_debug.restype = C.c_int
    lib.gcb_set_debug.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.gcb_set_chunk_bytes.argtypes = [C.c_void_p, C.c_int64]
    lib.gcb_launch_count.restype = C.c_int64
    lib.gcb_launch_count.argtypes = [C.c_void_p]
    if lib.gcb_abi_version() != GCB_ABI_VERSION:
        raise RuntimeError(f"{path}: ABI version {lib.gcb_abi_version()} != {GCB_ABI_VERSION}")
    return lib


class ConsensusEngine:
    """One context on one GPU.  `options` are the Options fields the hot path reads."""

    def __init__(self, options: Optional[Options] = None, device: int = 0, lib_path: Optional[str] = None):
        self.lib = load_library(lib_path)
        self.options = options or Options.default()
        self.device = device
        self._ctx = C.c_void_p()
        rc = self.lib.gcb_create(C.byref(self.options), device, C.byref(self._ctx))
        if rc == GCB_ERR_NO_DEVICE:
            raise EngineError(rc, "no sm_100 (B200) device: the consensus engine has no CPU path")
        if rc != GCB_OK:
            raise EngineError(rc, "gcb_create failed")
        self._genome_keepalive = None

    # -- lifecycle
    def close(self) -> None:
        if self._ctx:
            self.lib.gcb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc != GCB_OK:
            raise EngineError(rc, (self.lib.gcb_last_error(self._ctx) or b"").decode())

    # -- Reference::instance()->getData (reference.cpp:33-71)
    def set_reference(self, genome: Optional[Genome]) -> None:
        if genome is None:
            self._check(self.lib.gcb_set_reference(self._ctx, None, 0, None, None, 0))
            return
        self._check(self.lib.gcb_set_reference(self._ctx, genome.packed4.ctypes.data, len(genome.packed4), genome.contig_off.ctypes.data,
                                               genome.contig_len.ctypes.data, len(genome.contig_len)))

    def set_reference_device(self, packed4_ptr: int, packed_bytes: int, contig_off: np.ndarray, contig_len: np.ndarray, keepalive=None) -> None:
        """The packed genome already lives on the device (e.g. after an NCCL broadcast)."""
        self._genome_keepalive = keepalive
        self._check(self.lib.gcb_set_reference_device(self._ctx, packed4_ptr, packed_bytes, contig_off.ctypes.data,
                                                      contig_len.ctypes.data, len(contig_len)))

    # -- Cluster::clusterByUMI over a batch of clusters, host buffers in and out
    def cluster_by_umi(self, batch: Batch, result: Optional[Result] = None) -> Result:
        res = result or Result.allocate(batch)
        bs, rs = batch.as_struct(), res.as_struct()
        self._check(self.lib.gcb_consensus_batch(self._ctx, C.byref(bs), C.byref(rs)))
        return res

    # -- same with every buffer resident on the device (pointers are integers, e.g. torch .data_ptr())
    def cluster_by_umi_device(self, bs: BatchStruct, rs: ResultStruct, stages: int = STAGE_ALL, stream: int = 0) -> None:
        self._check(self.lib.gcb_consensus_batch_device(self._ctx, C.byref(bs), C.byref(rs), stages, C.c_void_p(stream)))

    # -- BamUtil::getUMI over a list of names (bamutil.cpp:40-112), encoded for Batch.umi
    def extract_umi(self, names, prefix: str, umi_words: int = 1):
        """names: list of bytes (qname, or the MI:Z value when the record has one).  Returns (uint64 [n, umi_words], uint8 status [n])."""
        blob = b"".join(names)
        off = np.zeros(len(names) + 1, np.int64)
        np.cumsum([len(x) for x in names], out=off[1:])
        buf = np.frombuffer(blob, np.uint8) if blob else np.zeros(1, np.uint8)
        out = np.zeros((len(names), umi_words), np.uint64)
        status = np.zeros(max(len(names), 1), np.uint8)
        self._check(self.lib.gcb_extract_umi(self._ctx, buf.ctypes.data, off.ctypes.data, len(names), prefix.encode(), umi_words,
                                             out.ctypes.data, status.ctypes.data))
        return out, status[:len(names)]

    # -- FastaReader::readAll + to4bits (fastareader.cpp:58-152): FASTA text -> Genome
    def pack_fasta(self, text: bytes, max_contigs: int = 4096) -> Genome:
        buf = np.frombuffer(text, np.uint8) if len(text) else np.zeros(1, np.uint8)
        cap = len(text) // 2 + 16 * max_contigs + 16
        out = np.zeros(cap, np.uint8)
        off, ln, noff = (np.zeros(max(max_contigs, 1), np.int64) for _ in range(3))
        nlen = np.zeros(max(max_contigs, 1), np.int32)
        nc, nb = C.c_int32(0), C.c_int64(0)
        self._check(self.lib.gcb_pack_fasta(self._ctx, buf.ctypes.data, len(text), max_contigs, out.ctypes.data, cap, off.ctypes.data, ln.ctypes.data,
                                            noff.ctypes.data, nlen.ctypes.data, C.byref(nc), C.byref(nb)))
        n = nc.value
        names = [text[int(noff[i]):int(noff[i]) + int(nlen[i])].decode("latin-1") for i in range(n)]
        return Genome(np.ascontiguousarray(out[:max(nb.value, 16)]), off[:n].copy(), ln[:n].copy(), names)

    def set_chunk_bytes(self, nbytes: int) -> None:
        """Payload bytes per pipeline chunk of cluster_by_umi (tuning only: results do not depend on it)."""
        self._check(self.lib.gcb_set_chunk_bytes(self._ctx, nbytes))

    def set_debug(self, key: int, value: int) -> None:
        """Tuning / test aid (see gcb_set_debug): 2 = tile window shift, 3 = lanes per cluster, 5 = generic kernel only."""
        self._check(self.lib.gcb_set_debug(self._ctx, key, value))

    def set_slow_queue_bytes(self, nbytes: int) -> None:
        """Bytes of the slow-column queue (tuning / tests: tiles whose columns do not fit are voted by the generic kernel)."""
        self._check(self.lib.gcb_set_slow_queue_bytes(self._ctx, nbytes))

    def cluster_stats(self, reset: bool = True):
        """What Cluster::clusterByUMI added to preStats / postStats for every cluster processed since the last reset
        (gcb_cluster_stats, summed on the device): a hoststats.ClusterStats."""
        from .hoststats import MAX_SUPPORTING_READS, ClusterStats
        raw = np.zeros(10 + MAX_SUPPORTING_READS, np.int64)
        self._check(self.lib.gcb_get_cluster_stats(self._ctx, raw.ctypes.data, 1 if reset else 0))
        v = [int(x) for x in raw[:10]]
        return ClusterStats(pre_cluster=v[0], pre_multi_cluster=v[1], pre_molecule=v[2], pre_molecule_se=v[3], pre_molecule_pe=v[4],
                            pre_uncounted=v[5], pre_hist=raw[10:].copy(), post_cluster=v[6], post_multi_cluster=v[7], post_sscs=v[8],
                            post_dcs=v[9])

    def stat_depth(self, tid, pos, l_qseq, target_len, coverage_step: int = 10000, depth=None):
        """Stats::statDepth over reads (stats.cpp:56-83): returns the concatenated per-contig depth bins (int64), added to `depth`
        when one is given."""
        tid, pos, l_qseq = (np.ascontiguousarray(a, np.int32) for a in (tid, pos, l_qseq))
        target_len = np.ascontiguousarray(target_len, np.int64)
        bins = int((1 + target_len // coverage_step).sum())
        out = np.zeros(bins, np.int64) if depth is None else np.ascontiguousarray(depth, np.int64)
        assert len(out) == bins
        self._check(self.lib.gcb_stat_depth(self._ctx, tid.ctypes.data, pos.ctypes.data, l_qseq.ctypes.data, len(tid), coverage_step,
                                            target_len.ctypes.data, len(target_len), out.ctypes.data))
        return out

    def batch_status(self, stream: int = 0) -> int:
        return self.lib.gcb_batch_status(self._ctx, C.c_void_p(stream))

    @property
    def launches(self) -> int:
        return int(self.lib.gcb_launch_count(self._ctx))
