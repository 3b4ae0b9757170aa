"""Shared comparison helpers for the parity tests (imported by tests only)."""
from __future__ import annotations

import numpy as np

from gencore_b200.abi import GROUP_DCS, GROUP_SSCS, Batch, Result, align4
from gencore_b200.hoststats import group_slots, patched_nm, stats_from_result, tag_byte


def assert_results_equal(batch: Batch, a: Result, b: Result, what: str = "") -> None:
    """Bit-exact comparison of two gcb_result (e.g. CUDA vs oracle)."""
    np.testing.assert_array_equal(a.cluster_n_groups, b.cluster_n_groups, err_msg=f"{what} cluster_n_groups")
    np.testing.assert_array_equal(a.pair_group, b.pair_group, err_msg=f"{what} pair_group")
    slots = group_slots(batch, a)
    ga, gb = a.groups[slots], b.groups[slots]
    for name in ga.dtype.names:
        if not np.array_equal(ga[name], gb[name]):
            bad = np.flatnonzero((ga[name] != gb[name]).reshape(len(ga), -1).any(axis=1))
            raise AssertionError(f"{what} groups[{name}] differ at {len(bad)} slots, first slot {slots[bad[0]]}: "
                                 f"{ga[name][bad[0]]} vs {gb[name][bad[0]]}\n{ga[bad[0]]}\n{gb[bad[0]]}")
    assert int(a.out_bytes[0]) == int(b.out_bytes[0]), f"{what} out_bytes {a.out_bytes[0]} vs {b.out_bytes[0]}"
    n = int(a.out_bytes[0])
    if not np.array_equal(a.out_payload[:n], b.out_payload[:n]):
        bad = np.flatnonzero(a.out_payload[:n] != b.out_payload[:n])
        # locate the record
        for s in slots:
            for side in range(2):
                t = int(a.groups[s]["tmpl_read"][side])
                if t < 0:
                    continue
                l = int(batch.reads["l_qseq"][t])
                off = int(a.groups[s]["out_off"][side])
                if off <= bad[0] < off + align4(l) + align4((l + 1) // 2):
                    raise AssertionError(f"{what} out_payload differs at byte {bad[0]} (slot {s} side {side} rel {bad[0] - off}, "
                                         f"l_qseq {l}): {a.out_payload[bad[0]]} vs {b.out_payload[bad[0]]}; {len(bad)} bytes differ")
        raise AssertionError(f"{what} out_payload differs at byte {bad[0]} ({len(bad)} bytes)")


def assert_matches_reference(batch: Batch, res: Result, ref_pairs, ref_out, ref_stats, n_ref_pairs: int) -> None:
    """Compare a gcb_result with what the reference's own Cluster::clusterByUMI returned (ref_harness)."""
    slots = group_slots(batch, res)
    g = res.groups[slots]
    kept = (g["status"] == GROUP_SSCS) | (g["status"] == GROUP_DCS)
    assert int(kept.sum()) == n_ref_pairs, f"kept groups {int(kept.sum())} vs reference pairs {n_ref_pairs}"
    # index kept groups by their template slots
    by_tmpl = {}
    empties = {}
    cluster_of = np.repeat(np.arange(batch.n_clusters), res.cluster_n_groups)
    for s, gr, c in zip(slots[kept], g[kept], cluster_of[kept]):
        key = (int(gr["tmpl_read"][0]), int(gr["tmpl_read"][1]))
        if key == (-1, -1):
            empties.setdefault(int(c), []).append(int(s))
        else:
            assert key not in by_tmpl
            by_tmpl[key] = int(s)
    nm = batch.nm if batch.nm is not None else np.zeros(2 * batch.n_pairs, np.uint8)
    for rp in ref_pairs:
        key = (int(rp["slot"][0]), int(rp["slot"][1]))
        if key == (-1, -1):
            lst = empties.get(int(rp["cluster"]), [])
            assert lst, f"reference returned an empty Pair in cluster {rp['cluster']} that the result lacks"
            s = lst.pop()
        else:
            assert key in by_tmpl, f"reference returned templates {key} (cluster {rp['cluster']}); no kept group has them"
            s = by_tmpl.pop(key)
        gr = res.groups[s]
        ctx = f"slot {s} cluster {rp['cluster']}"
        assert int(gr["merge_reads"]) == int(rp["merge_reads"]), ctx
        assert (int(gr["status"]) == GROUP_DCS) == bool(rp["is_duplex"]), ctx
        if rp["is_duplex"]:
            assert int(gr["reverse_merge_reads"]) == int(rp["reverse_merge_reads"]), ctx
        for side in range(2):
            t = int(gr["tmpl_read"][side])
            if t < 0:
                continue
            assert int(gr["diff"][side]) == int(rp["diff"][side]), f"{ctx} side {side} diff"
            l = int(batch.reads["l_qseq"][t])
            off, roff = int(gr["out_off"][side]), int(rp["out_off"][side])
            mine = res.out_payload[off:off + align4(l) + align4((l + 1) // 2)]
            theirs = ref_out[roff:roff + align4(l) + align4((l + 1) // 2)]
            if not np.array_equal(mine, theirs):
                bad = np.flatnonzero(mine != theirs)
                raise AssertionError(f"{ctx} side {side}: consensus record differs at rel byte {bad[0]} of l_qseq {l} "
                                     f"({len(bad)} bytes): {mine[bad[0]]} vs {theirs[bad[0]]}")
            donor = int(gr["qname_donor"][side])
            name_read = donor if donor >= 0 else t
            assert name_read // 2 == int(rp["name_slot"][side]) // 2 or \
                bytes(batch.qnames[name_read // 2]) == bytes(batch.qnames[int(rp["name_slot"][side]) // 2]), f"{ctx} side {side} qname"
            assert int(rp["l_qname"][side]) == int(batch.reads["l_qname"][name_read]), f"{ctx} side {side} l_qname"
            assert int(rp["nm"][side]) == patched_nm(int(nm[t]), int(gr["mismatch_inc"][side])), f"{ctx} side {side} NM"
            assert int(rp["fr"][side]) == tag_byte(int(gr["merge_reads"])), f"{ctx} side {side} FR"
            if rp["is_duplex"]:
                assert int(rp["rr"][side]) == tag_byte(int(gr["reverse_merge_reads"])), f"{ctx} side {side} RR"
            else:
                assert int(rp["rr"][side]) == -1, f"{ctx} side {side} RR present"
    st = stats_from_result(batch, res)
    for name in ("pre_cluster", "pre_multi_cluster", "pre_molecule", "pre_molecule_se", "pre_molecule_pe", "pre_uncounted",
                 "post_cluster", "post_multi_cluster", "post_sscs", "post_dcs"):
        assert getattr(st, name) == int(ref_stats[name]), f"stats {name}: {getattr(st, name)} vs {int(ref_stats[name])}"
    np.testing.assert_array_equal(st.pre_hist, ref_stats["pre_hist"], err_msg="stats pre_hist")


# ----------------------------------------------------------------------------- golden fixtures (tests/golden/*.npz)

def save_golden(path, batch, genome, contigs, opt, ref_pairs, ref_out, ref_stats, n_ref):
    import ctypes
    qn = np.asarray([bytes(q) for q in batch.qnames], dtype="S")
    np.savez_compressed(
        path, cluster_pair_off=batch.cluster_pair_off, cluster_ref=batch.cluster_ref, cluster_flags=batch.cluster_flags,
        umi=batch.umi, reads=batch.reads, cigar=batch.cigar, payload=batch.payload, qnames=qn, nm=batch.nm,
        umi_prefix=np.asarray(batch.umi_prefix), packed4=genome.packed4, contig_off=genome.contig_off,
        contig_len=genome.contig_len, names=np.asarray(genome.names),
        opt=np.frombuffer(ctypes.string_at(ctypes.addressof(opt), ctypes.sizeof(opt)), np.uint8),
        ref_pairs=ref_pairs, ref_out=ref_out, ref_stats=np.asarray([ref_stats]), n_ref=np.asarray(n_ref))


def load_golden(path):
    from gencore_b200.abi import Genome, Options
    z = np.load(path, allow_pickle=False)
    batch = Batch(z["cluster_pair_off"], z["cluster_ref"], z["cluster_flags"], z["umi"], z["reads"], z["cigar"], z["payload"],
                  list(z["qnames"]), z["nm"], str(z["umi_prefix"]))
    genome = Genome(z["packed4"], z["contig_off"], z["contig_len"], [str(x) for x in z["names"]])
    opt = Options.from_buffer_copy(z["opt"].tobytes())
    return batch, genome, opt, z["ref_pairs"], z["ref_out"], z["ref_stats"][0], int(z["n_ref"])
