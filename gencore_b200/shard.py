"""Coordinate-window sharding of the hot path across the GPUs of one box.

Every Cluster (tid, left, right) is independent (gencore.cpp:76, cluster.cpp:55), and a batch holds its
clusters in coordinate order, so rank r owns one contiguous run of clusters — a genomic window — balanced
by payload bytes.  No data-path collective: the packed genome is broadcast once, the additive Stats
counters (stats.h:47-65) are summed once at the end.  torch.distributed is used as plumbing only.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .abi import Batch, Genome
from .hoststats import ClusterStats


def window_bounds(batch: Batch, world: int) -> np.ndarray:
    """int64 [world+1]: rank r owns clusters [b[r], b[r+1]); windows are balanced by payload bytes."""
    slab = batch.cluster_slab_bounds()
    total = int(slab[-1] - slab[0]) if batch.n_clusters else 0
    targets = slab[0] + (np.arange(1, world, dtype=np.int64) * total) // world
    inner = np.searchsorted(slab[:-1], targets, side="left") if batch.n_clusters else np.zeros(world - 1, np.int64)
    return np.concatenate([[0], inner, [batch.n_clusters]]).astype(np.int64)


def shard_batch(batch: Batch, world: int, rank: int) -> Tuple[Batch, Tuple[int, int]]:
    """The sub-batch of rank `rank` (offsets rebased so that it is a valid gcb_batch on its own)."""
    b = window_bounds(batch, world)
    c0, c1 = int(b[rank]), int(b[rank + 1])
    return slice_batch(batch, c0, c1), (c0, c1)


def slice_batch(batch: Batch, c0: int, c1: int) -> Batch:
    """Clusters [c0, c1) as a gcb_batch of their own."""
    p0, p1 = int(batch.cluster_pair_off[c0]), int(batch.cluster_pair_off[c1])
    slab = batch.cluster_slab_bounds()
    s0, s1 = int(slab[c0]), int(slab[c1])
    reads = batch.reads[2 * p0:2 * p1].copy()
    live = reads["l_qseq"] >= 0
    reads["data_off"][live] -= s0
    sub = Batch(
        (batch.cluster_pair_off[c0:c1 + 1] - p0).astype(np.int32), batch.cluster_ref[c0:c1].copy(), batch.cluster_flags[c0:c1].copy(),
        batch.umi[p0:p1].copy(), reads, batch.cigar, np.ascontiguousarray(batch.payload[s0:s1]),
        batch.qnames[p0:p1] if batch.qnames is not None else None, batch.nm[2 * p0:2 * p1] if batch.nm is not None else None,
        batch.umi_prefix)
    return sub


def split_batch(batch: Batch, parts: int) -> List[Batch]:
    """The batch cut into `parts` coordinate windows (every one a valid gcb_batch), in order."""
    return [shard_batch(batch, parts, r)[0] for r in range(parts)]


STATS_FIELDS = ("pre_cluster", "pre_multi_cluster", "pre_molecule", "pre_molecule_se", "pre_molecule_pe", "pre_uncounted",
                "post_cluster", "post_multi_cluster", "post_sscs", "post_dcs")


def stats_to_vector(st: ClusterStats) -> np.ndarray:
    return np.concatenate([np.asarray([getattr(st, f) for f in STATS_FIELDS], np.int64), st.pre_hist.astype(np.int64)])


def stats_from_vector(v: np.ndarray) -> ClusterStats:
    st = ClusterStats()
    for k, f in enumerate(STATS_FIELDS):
        setattr(st, f, int(v[k]))
    st.pre_hist = np.asarray(v[len(STATS_FIELDS):], np.int64).copy()
    return st


def broadcast_genome(genome: Genome, device, src: int = 0):
    """ONE broadcast of the packed reference from rank `src`; returns the device tensor every rank keeps."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(genome.packed4).to(device) if dist.get_rank() == src else torch.empty(len(genome.packed4), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=src)
    return t


def reduce_stats(st: ClusterStats, device) -> ClusterStats:
    """The final gather of per-rank Stats: every counter is additive, so it is one all-reduce(sum)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(stats_to_vector(st)).to(device)
    dist.all_reduce(t)
    return stats_from_vector(t.cpu().numpy())


def concat_results(parts: List[dict]) -> dict:
    """Ordered concatenation of per-rank results (window order = coordinate order)."""
    return {
        "cluster_n_groups": np.concatenate([p["cluster_n_groups"] for p in parts]),
        "pair_group": np.concatenate([p["pair_group"] for p in parts]),
        "out_payload": np.concatenate([p["out_payload"] for p in parts]),
    }
