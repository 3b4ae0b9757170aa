"""Parity tests proper: the CUDA library (gencore_b200/csrc/libgencore_b200.so) through its C ABI on a
B200 against the oracle, bit-exact, on the shared case list; then the BASELINE.json shapes at bench size
(one million pairs each) compared record for record with the oracle, plus size-independent properties."""
import dataclasses
import os

import numpy as np
import pytest

import parity_cases
from gencore_b200.abi import GROUP_DCS, GROUP_SSCS, STAGE_ALL, Options, Result
from gencore_b200.hoststats import group_slots
from parity import assert_results_equal

pytestmark = pytest.mark.gpu

CASES = parity_cases.gpu_cases()


@pytest.fixture(scope="module")
def engine_cls():
    import torch
    assert torch.cuda.is_available(), "the gpu tests need a CUDA device"
    from gencore_b200.engine import ConsensusEngine
    return ConsensusEngine


@pytest.mark.parametrize("name,thunk", CASES, ids=[c[0] for c in CASES])
def test_cuda_matches_oracle_host_buffers(engine_cls, oracle, name, thunk):
    batch, genome, opt = thunk()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        res = eng.cluster_by_umi(batch)
        res2 = eng.cluster_by_umi(batch)  # the second run must give the same answer (no state leaks between batches)
        assert eng.launches > 0 or batch.n_clusters == 0
    expect = oracle.consensus(batch, genome, opt)
    assert_results_equal(batch, res, expect, name)
    assert_results_equal(batch, res2, expect, name + " (second run)")


@pytest.mark.parametrize("name", ["cfg2_40k", "ragged_duplex_5_big", "deep_1100", "cfg5_40k", "deep400_40k"])
def test_cuda_matches_oracle_device_buffers(engine_cls, oracle, name):
    import torch
    from gencore_b200.device import DeviceBatch, DeviceResult
    batch, genome, opt = dict(CASES)[name]()
    dev = torch.device("cuda:0")
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        db = DeviceBatch.from_host(batch, dev)
        dr = DeviceResult.allocate(batch.n_pairs, batch.n_clusters, len(batch.payload), dev)
        torch.cuda.synchronize()
        for _ in range(2):
            eng.cluster_by_umi_device(db.struct, dr.struct, STAGE_ALL, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert eng.batch_status() == 0
            res = dr.to_host()
            assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)


@pytest.mark.parametrize("name,thunk", CASES, ids=[c[0] for c in CASES])
def test_generic_vote_matches_oracle(engine_cls, oracle, name, thunk):
    """score_vote_kernel (no size limits; takes the tiles the ring kernel cannot stage) on every tile of every case."""
    batch, genome, opt = thunk()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        eng.set_debug(5, 1)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)


@pytest.mark.parametrize("name,thunk", CASES, ids=[c[0] for c in CASES])
def test_small_tile_window_matches_oracle(engine_cls, oracle, name, thunk):
    """The vote over 16 KB payload windows (twice the tiles, the voter warps in three groups) on every case."""
    batch, genome, opt = thunk()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        eng.set_debug(2, 14)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)


@pytest.mark.parametrize("lanes", [8, 16, 32])
@pytest.mark.parametrize("name", ["cfg2_40k", "cfg4_40k", "edge_default", "ragged_duplex_5_big", "ragged_single_4_big", "deep_1100", "low_complexity",
                                  "wide_umi_3", "cfg3_40k", "tiny_reads", "cfg5_40k", "cfg3_crowded"])
def test_lanes_per_cluster_do_not_change_results(engine_cls, oracle, name, lanes):
    """umi_group_kernel / select_template_kernel with 8, 16 or 32 lanes per cluster."""
    batch, genome, opt = dict(CASES)[name]()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        eng.set_debug(3, lanes)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), f"{name} lanes {lanes}")


@pytest.mark.parametrize("name", ["cfg2_40k", "cfg3_40k", "cfg4_40k", "ragged_duplex_5_big", "edge_strict", "deep_1100"])
def test_cluster_stats_kernel_matches_the_host_replay(engine_cls, name):
    """cluster_stats_kernel (gcb_get_cluster_stats) against hoststats, the replay pinned to the reference's Stats in tests/parity.py."""
    import dataclasses
    from gencore_b200.hoststats import stats_from_result
    batch, genome, opt = dict(CASES)[name]()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        eng.set_chunk_bytes(1 << 20)
        res = eng.cluster_by_umi(batch)
        got = eng.cluster_stats()
    want = stats_from_result(batch, res)
    for f in dataclasses.fields(got):
        assert np.array_equal(getattr(got, f.name), getattr(want, f.name)), (name, f.name, getattr(got, f.name), getattr(want, f.name))


@pytest.mark.parametrize("qbytes", [1 << 14, 1 << 20])
@pytest.mark.parametrize("name", ["cfg2_40k", "ragged_duplex_5_big", "cfg4_40k", "edge_strict", "deep50_noisy_40k"])
def test_slow_queue_overflow_does_not_change_results(engine_cls, oracle, name, qbytes):
    """A slow-column queue far too small: the tiles whose columns do not fit are redone by the generic kernel."""
    batch, genome, opt = dict(CASES)[name]()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        eng.set_slow_queue_bytes(qbytes)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), f"{name} queue {qbytes}")


@pytest.mark.parametrize("name", ["cfg3_40k", "ragged_duplex_5_big", "deep_1100", "cfg4_40k", "cfg5_40k", "deep50_noisy_40k"])
@pytest.mark.parametrize("chunk", [1 << 14, 1 << 18, 1 << 21])
def test_pipeline_chunks_do_not_change_results(engine_cls, oracle, name, chunk):
    batch, genome, opt = dict(CASES)[name]()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        eng.set_chunk_bytes(chunk)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), f"{name} chunk {chunk}")


@pytest.mark.parametrize("write_combined", [False, True])
def test_batches_in_gcb_host_alloc_memory(engine_cls, oracle, write_combined):
    """The batch arrays in page-locked memory from gcb_host_alloc (plain and write-combined) give the same results."""
    from gencore_b200.device import pinned_copy
    batch, genome, opt = dict(CASES)["cfg2_40k"]()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        pb = pinned_copy(batch, lib=eng.lib, write_combined=write_combined)
        res = eng.cluster_by_umi(pb)
        del pb
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), "cfg2_40k in gcb_host_alloc memory")


def test_capacity_error_is_reported(engine_cls):
    from gencore_b200.abi import GCB_ERR_CAPACITY
    from gencore_b200.engine import EngineError
    batch, genome, opt = parity_cases.fixed_case("cfg2", 1500)
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        small = Result.allocate(batch, out_capacity=64)
        with pytest.raises(EngineError) as ei:
            eng.cluster_by_umi(batch, small)
        assert ei.value.code == GCB_ERR_CAPACITY


def _oracle_parallel(batch, genome, opt, n_parts):
    """The oracle over the whole batch, clusters split over host processes (results are per cluster, so the parts concatenate)."""
    import multiprocessing as mp
    from gencore_b200.shard import split_batch
    parts = split_batch(batch, n_parts)
    with mp.get_context("fork").Pool(min(n_parts, os.cpu_count() or 1)) as pool:
        return parts, pool.starmap(_oracle_part, [(p, genome, opt) for p in parts])


def _oracle_part(part, genome, opt):
    from oracle.pyoracle import Oracle
    return Oracle().consensus(part, genome, opt)


FULL = [("cfg1", 10_000), ("cfg2", 1_000_000), ("cfg3", 1_000_000), ("cfg4", 1_000_000), ("cfg5", 1_000_000)]


@pytest.mark.parametrize("name,n_pairs", FULL, ids=[f"{n}_{p}" for n, p in FULL])
def test_full_size_matches_oracle(engine_cls, name, n_pairs):
    """Every BASELINE.json shape at bench size (cfg1 at its own 10 k pairs, the others at the 1 M pairs per GPU that bench.py
    times): the WHOLE result against the oracle, record for record, plus properties that need no oracle."""
    from gencore_b200 import synth
    cfg = synth.CONFIGS[name]
    cfg = dataclasses.replace(cfg, n_contigs=min(cfg.n_contigs, 2), contig_len=min(cfg.contig_len, 20_000_000))
    batch, genome, _ = synth.make_batch(cfg, seed=20261019, n_pairs=n_pairs, with_qnames=False)
    opt = Options.default(cluster_size_req=cfg.supporting_reads)
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        res = eng.cluster_by_umi(batch)
        res2 = eng.cluster_by_umi(batch)
    n = int(res.out_bytes[0])
    assert np.array_equal(res.groups, res2.groups) and np.array_equal(res.out_payload[:n], res2.out_payload[:n]), "not deterministic"
    slots = group_slots(batch, res)
    g = res.groups[slots]
    assert (res.pair_group >= 0).all()
    assert int(g["merge_reads"].sum()) == batch.n_pairs, "every pair is in exactly one family"
    # consensus records tile the output exactly, in slot order
    l = batch.reads["l_qseq"][np.maximum(g["tmpl_read"], 0)].astype(np.int64)
    sz = np.where(g["tmpl_read"] >= 0, ((l + 3) & ~3) + (((l + 1) // 2 + 3) & ~3), 0)
    offs = g["out_off"].reshape(-1)[(g["tmpl_read"] >= 0).reshape(-1)]
    assert np.array_equal(offs, np.concatenate([[0], np.cumsum(sz.reshape(-1)[sz.reshape(-1) > 0])[:-1]]))
    assert n == int(sz.sum())
    if name == "cfg2":  # no UMI pair is a duplex there, every family survives -s 1
        assert ((g["status"] == GROUP_SSCS) | (g["status"] == GROUP_DCS)).all()
    # the whole batch against the oracle (host cores in parallel: clusters are independent)
    parts, refs = _oracle_parallel(batch, genome, opt, 16)
    from gencore_b200.verify import assert_window_equal
    c0 = p0 = out0 = 0
    for part, ref in zip(parts, refs):
        out0 += assert_window_equal(res, part, ref, c0, p0, name)
        c0, p0 = c0 + part.n_clusters, p0 + part.n_pairs
    assert out0 == n and p0 == batch.n_pairs
