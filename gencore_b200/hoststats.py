"""Host-side bookkeeping the reference does inside Cluster::clusterByUMI next to the arithmetic:
the Stats side effects (cluster.cpp:102,136,143,157,161,172,176,184-186) and the tag/NM values the
caller writes into the BAM records (pair.cpp:54-68, group.cpp:568-572), derived from gcb_result.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .abi import GROUP_DCS, GROUP_DUPLEX_PARTNER, GROUP_SSCS, Batch, Result

MAX_SUPPORTING_READS = 100  # stats.h:15


@dataclass
class ClusterStats:
    pre_cluster: int = 0
    pre_multi_cluster: int = 0
    pre_molecule: int = 0
    pre_molecule_se: int = 0
    pre_molecule_pe: int = 0
    pre_uncounted: int = 0
    pre_hist: np.ndarray = field(default_factory=lambda: np.zeros(MAX_SUPPORTING_READS, np.int64))
    post_cluster: int = 0
    post_multi_cluster: int = 0
    post_sscs: int = 0
    post_dcs: int = 0


def group_slots(batch: Batch, result: Result) -> np.ndarray:
    """Indices into result.groups of every live group slot (cluster_pair_off[c] + g, g < n_groups[c])."""
    ng = result.cluster_n_groups.astype(np.int64)
    base = np.repeat(batch.cluster_pair_off[:-1].astype(np.int64), ng)
    within = np.arange(int(ng.sum())) - np.repeat(np.cumsum(ng) - ng, ng)
    return base + within


def stats_from_result(batch: Batch, result: Result) -> ClusterStats:
    st = ClusterStats()
    ng = result.cluster_n_groups
    st.pre_cluster = int(batch.n_clusters)
    st.pre_multi_cluster = int((ng > 1).sum())
    slots = group_slots(batch, result)
    g = result.groups[slots]
    cluster_of = np.repeat(np.arange(batch.n_clusters), ng)
    mol = g["status"] != GROUP_DUPLEX_PARTNER  # every group is one addMolecule, except a consumed duplex partner
    supporting = g["merge_reads"].astype(np.int64).copy()
    has_partner = g["duplex_partner"] >= 0
    partner_slot = batch.cluster_pair_off[:-1].astype(np.int64)[cluster_of] + np.maximum(g["duplex_partner"], 0)
    supporting[has_partner] += result.groups["merge_reads"][partner_slot[has_partner]]
    pe = (g["tmpl_read"][:, 0] >= 0) & (g["tmpl_read"][:, 1] >= 0)
    st.pre_molecule = int(mol.sum())
    st.pre_molecule_pe = int((mol & pe).sum())
    st.pre_molecule_se = st.pre_molecule - st.pre_molecule_pe
    sup = supporting[mol]
    st.pre_uncounted = int((sup >= MAX_SUPPORTING_READS).sum())
    st.pre_hist = np.bincount(sup[sup < MAX_SUPPORTING_READS], minlength=MAX_SUPPORTING_READS).astype(np.int64)
    kept = (g["status"] == GROUP_SSCS) | (g["status"] == GROUP_DCS)
    st.post_sscs = int((g["status"] == GROUP_SSCS).sum())
    st.post_dcs = int((g["status"] == GROUP_DCS).sum())
    kept_per_cluster = np.bincount(cluster_of[kept], minlength=batch.n_clusters)
    st.post_cluster = int((kept_per_cluster > 0).sum())
    st.post_multi_cluster = int((kept_per_cluster > 1).sum())
    return st


def patched_nm(orig_nm: int, mismatch_inc: int) -> int:
    """group.cpp:527-572 for an NM:C tag: patched only when 0 < |inc|, inc <= 5 and the new value fits a byte."""
    if mismatch_inc == 0 or mismatch_inc > 5:
        return orig_nm
    new = orig_nm + mismatch_inc
    return new if 0 <= new <= 255 else orig_nm


def tag_byte(v: int) -> int:
    """pair.cpp:57-64: one byte appended from the address of an unsigned short => value mod 256."""
    return min(v, 65535) & 0xFF
