"""The oracle restatement (oracle/gencore_oracle.c) against the reference ITSELF (oracle/_ref: the
reference's unmodified sources + ref_harness.cpp) on randomised and hand-built batches.
Skipped when oracle/_ref is not built (the golden fixtures in test_oracle_golden.py cover that case)."""
import numpy as np
import pytest

import cases
from gencore_b200 import synth
from gencore_b200.abi import GROUP_DCS, GROUP_DUPLEX_DIFF, Options
from gencore_b200.hoststats import group_slots
from parity import assert_matches_reference


def _check(oracle, batch, genome, contigs, opt):
    from oracle.pyoracle import Reference
    res = oracle.consensus(batch, genome, opt)
    ref = Reference(opt, batch.umi_prefix, genome, contigs)
    try:
        pairs, out, stats, _secs, n = ref.consensus(batch)
    finally:
        ref.close()
    assert_matches_reference(batch, res, pairs, out, stats, n)
    return res


@pytest.fixture(autouse=True)
def _need_ref(have_reference):
    if not have_reference:
        pytest.skip("oracle/_ref not built")


@pytest.mark.parametrize("optname", list(cases.OPTION_SETS))
def test_edge_cases(oracle, optname):
    batch, genome, contigs = cases.edge_batch()
    res = _check(oracle, batch, genome, contigs, cases.OPTION_SETS[optname])
    if optname == "default":
        g = res.groups[group_slots(batch, res)]
        assert (g["mismatch_inc"] > 5).any(), "rollback case not exercised"
        assert ((g["mismatch_inc"] != 0) & (g["mismatch_inc"] <= 5)).any(), "NM patch case not exercised"
        assert (g["status"] == GROUP_DCS).any() and (g["status"] == GROUP_DUPLEX_DIFF).any()
        assert (g["qname_donor"] >= 0).any()
        assert (g["tmpl_read"] < 0).any()


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("umi", ["none", "single", "duplex"])
def test_ragged_random(oracle, seed, umi):
    batch, genome, contigs = synth.make_ragged_batch(100 + seed, n_clusters=60, umi=umi, err=0.02 if seed % 2 else 0.005)
    opt = list(cases.OPTION_SETS.values())[seed % len(cases.OPTION_SETS)]
    _check(oracle, batch, genome, contigs, opt)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4"])
def test_fixed_configs_small(oracle, name):
    cfg = synth.CONFIGS[name]
    import dataclasses
    small = dataclasses.replace(cfg, contig_len=200_000, n_contigs=min(cfg.n_contigs, 2))
    batch, genome, contigs = synth.make_fixed_batch(small, seed=20261017, n_pairs=3000)
    _check(oracle, batch, genome, contigs, Options.default(cluster_size_req=cfg.supporting_reads))


def test_no_reference_genome(oracle):
    batch, _genome, _contigs = synth.make_ragged_batch(5, n_clusters=40, umi="single")
    _check(oracle, batch, None, None, Options.default())


def test_deep_family_over_1000(oracle):
    batch, genome, contigs = cases.deep_batch()
    _check(oracle, batch, genome, contigs, Options.default())


def test_low_complexity_skip(oracle):
    batch, genome, contigs = cases.low_complexity_batch()
    res = _check(oracle, batch, genome, contigs, Options.default())
    assert (res.groups[0]["tmpl_read"] == -1).all()
