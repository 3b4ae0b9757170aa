"""Comparison of a result over a whole batch with a result computed over a window of its clusters (the same rows, rebased).
No arithmetic of the hot path lives here; bench.py and the tests hand it the reference result (from the oracle)."""
from __future__ import annotations

import numpy as np

from .abi import Batch, Result
from .hoststats import group_slots


def assert_window_equal(res: Result, part: Batch, ref: Result, c0: int, p0: int, what: str = "") -> int:
    """`ref` = the result of clusters [c0, c0 + part.n_clusters) alone (pairs from p0): every row and every consensus byte of
    `res` over that window must be identical once indices are rebased.  Returns the window's consensus bytes."""
    c1, p1 = c0 + part.n_clusters, p0 + part.n_pairs
    assert np.array_equal(res.cluster_n_groups[c0:c1], ref.cluster_n_groups), f"{what}: cluster_n_groups of clusters {c0}..{c1}"
    assert np.array_equal(res.pair_group[p0:p1], ref.pair_group), f"{what}: pair_group of pairs {p0}..{p1}"
    ps = group_slots(part, ref)
    if len(ps) == 0:
        return 0
    mine, theirs = res.groups[ps + p0], ref.groups[ps]
    have = theirs["tmpl_read"] >= 0
    # where the window's first consensus record lies in the whole output
    first = np.flatnonzero(have.reshape(-1))
    out0 = int(mine["out_off"].reshape(-1)[first[0]] - theirs["out_off"].reshape(-1)[first[0]]) if len(first) else 0
    for field in theirs.dtype.names:
        a, b = mine[field], theirs[field]
        if field in ("tmpl_read", "qname_donor"):
            b = np.where(b >= 0, b + 2 * p0, b)
        elif field == "umi_pair":
            b = np.where(b >= 0, b + p0, b)
        elif field == "out_off":
            b = np.where(have, b + out0, b)
            a = np.where(have, a, b)
        assert np.array_equal(a, b), f"{what}: groups[{field}] of pairs {p0}..{p1}"
    nb = int(ref.out_bytes[0])
    assert np.array_equal(res.out_payload[out0:out0 + nb], ref.out_payload[:nb]), f"{what}: consensus records of clusters {c0}..{c1}"
    return nb
