"""Device residency for batches and results: torch is used for HBM allocations, pinned host memory and
streams only (plumbing); every kernel is launched through the C ABI."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from .abi import GROUP_RESULT, Batch, BatchStruct, Result, ResultStruct


def _torch():
    import torch
    return torch


def _as_u8(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a).reshape(-1).view(np.uint8)


@dataclass
class DeviceBatch:
    """A gcb_batch whose arrays live in HBM."""
    struct: BatchStruct
    tensors: Dict[str, object]
    n_pairs: int
    n_clusters: int
    payload_bytes: int

    @staticmethod
    def from_host(batch: Batch, device) -> "DeviceBatch":
        torch = _torch()
        t = {}
        for name in ("cluster_pair_off", "cluster_ref", "cluster_flags", "umi", "reads", "cigar", "payload"):
            host = torch.from_numpy(_as_u8(getattr(batch, name)))
            dev = torch.empty(max(host.numel(), 16), dtype=torch.uint8, device=device)
            dev[:host.numel()].copy_(host)
            t[name] = dev
        s = BatchStruct(batch.n_clusters, batch.n_pairs, batch.umi_words, batch.max_cluster_bytes(),
                        t["cluster_pair_off"].data_ptr(), t["cluster_ref"].data_ptr(), t["cluster_flags"].data_ptr(),
                        t["umi"].data_ptr(), t["reads"].data_ptr(), t["cigar"].data_ptr(), len(batch.cigar),
                        t["payload"].data_ptr(), len(batch.payload))
        return DeviceBatch(s, t, batch.n_pairs, batch.n_clusters, len(batch.payload))


@dataclass
class DeviceResult:
    struct: ResultStruct
    tensors: Dict[str, object]
    n_pairs: int
    n_clusters: int

    @staticmethod
    def allocate(n_pairs: int, n_clusters: int, out_capacity: int, device) -> "DeviceResult":
        torch = _torch()
        t = {
            "pair_group": torch.empty(max(n_pairs, 4), dtype=torch.int32, device=device),
            "cluster_n_groups": torch.empty(max(n_clusters, 4), dtype=torch.int32, device=device),
            "groups": torch.empty(max(n_pairs, 1) * GROUP_RESULT.itemsize, dtype=torch.uint8, device=device),
            "out_payload": torch.empty(max(out_capacity, 16), dtype=torch.uint8, device=device),
            "out_bytes": torch.zeros(1, dtype=torch.int64, device=device),
        }
        s = ResultStruct(t["pair_group"].data_ptr(), t["cluster_n_groups"].data_ptr(), t["groups"].data_ptr(),
                         t["out_payload"].data_ptr(), max(out_capacity, 16), t["out_bytes"].data_ptr())
        return DeviceResult(s, t, n_pairs, n_clusters)

    def to_host(self, out: Optional[Result] = None) -> Result:
        used = int(self.tensors["out_bytes"].cpu()[0])
        res = out or Result(np.empty(self.n_pairs, np.int32), np.empty(self.n_clusters, np.int32),
                            np.empty(self.n_pairs, GROUP_RESULT), np.zeros(max(used, 16), np.uint8), np.zeros(1, np.int64))
        res.pair_group[:] = self.tensors["pair_group"][:self.n_pairs].cpu().numpy()
        res.cluster_n_groups[:] = self.tensors["cluster_n_groups"][:self.n_clusters].cpu().numpy()
        res.groups[:] = self.tensors["groups"][:self.n_pairs * GROUP_RESULT.itemsize].cpu().numpy().view(GROUP_RESULT)
        res.out_payload[:used] = self.tensors["out_payload"][:used].cpu().numpy()
        res.out_bytes[0] = used
        return res


class _HostBlock:
    """Page-locked host memory from gcb_host_alloc, freed with the object."""

    def __init__(self, lib, nbytes: int, write_combined: bool):
        import ctypes
        self.lib = lib
        self.ptr = lib.gcb_host_alloc(max(nbytes, 16), 1 if write_combined else 0)
        if not self.ptr:
            raise MemoryError("gcb_host_alloc failed")
        self.array = np.ctypeslib.as_array((ctypes.c_uint8 * max(nbytes, 16)).from_address(self.ptr))

    def __del__(self):
        try:
            self.lib.gcb_host_free(self.ptr)
        except Exception:
            pass


def pinned_copy(batch: Batch, lib=None, write_combined: bool = False) -> Batch:
    """The same batch with every array in page-locked host memory (what gcb_consensus_batch recommends).  With `lib` (the
    loaded C library) the memory comes from gcb_host_alloc, optionally write-combined (a packed batch is only written by the host)."""
    torch = _torch()
    keep = []

    def pin(a: np.ndarray) -> np.ndarray:
        flat = _as_u8(a)
        if lib is not None:
            blk = _HostBlock(lib, flat.size, write_combined)
            blk.array[:flat.size] = flat
            keep.append(blk)
            return blk.array[:flat.size].view(a.dtype).reshape(a.shape)
        t = torch.empty(max(flat.size, 16), dtype=torch.uint8).pin_memory()
        t[:flat.size].copy_(torch.from_numpy(flat))
        keep.append(t)
        return t.numpy()[:flat.size].view(a.dtype).reshape(a.shape)

    b = Batch(pin(batch.cluster_pair_off), pin(batch.cluster_ref), pin(batch.cluster_flags), pin(batch.umi), pin(batch.reads),
              pin(batch.cigar), pin(batch.payload), batch.qnames, batch.nm, batch.umi_prefix)
    b._pinned = keep  # keep the torch storages alive
    return b


def pinned_result(batch: Batch, out_capacity: Optional[int] = None) -> Result:
    torch = _torch()
    keep = []

    def pin(shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        t = torch.zeros(max(n, 16), dtype=torch.uint8).pin_memory()
        keep.append(t)
        return t.numpy()[:n].view(dtype).reshape(shape)

    cap = len(batch.payload) if out_capacity is None else out_capacity
    r = Result(pin((batch.n_pairs,), np.int32), pin((batch.n_clusters,), np.int32), pin((batch.n_pairs,), GROUP_RESULT),
               pin((max(cap, 16),), np.uint8), pin((1,), np.int64))
    r._pinned = keep
    return r
