"""Parity tests proper: the CUDA library (gencore_b200/csrc/libgencore_b200.so) through its C ABI on a
B200 against the oracle, bit-exact, on the shared case list; then size-independent properties at
BASELINE.json's bench size."""
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
        assert eng.launches > 0 or batch.n_clusters == 0
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)


@pytest.mark.parametrize("name", ["cfg2_40k", "ragged_duplex_5_big", "deep_1100"])
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
        for _ in range(2):  # the second run must give the same answer (no state leaks between batches)
            eng.cluster_by_umi_device(db.struct, dr.struct, STAGE_ALL, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert eng.batch_status() == 0
            res = dr.to_host()
            assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)


@pytest.mark.parametrize("name,thunk", CASES, ids=[c[0] for c in CASES])
def test_pipelined_vote_matches_oracle(engine_cls, oracle, name, thunk):
    """The persistent pipelined vote kernel (gcb_set_vote_mode 1) on every case."""
    batch, genome, opt = thunk()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        eng.set_vote_mode(1)
        res = eng.cluster_by_umi(batch)
        res2 = eng.cluster_by_umi(batch)
    expect = oracle.consensus(batch, genome, opt)
    assert_results_equal(batch, res, expect, name)
    assert_results_equal(batch, res2, expect, name + " (second run)")


@pytest.mark.parametrize("name,thunk", CASES, ids=[c[0] for c in CASES])
def test_tiled_vote_matches_oracle(engine_cls, oracle, name, thunk):
    """vote_tiled_kernel (gcb_set_vote_mode 0: every CTA builds its tile's family-side table itself) on every case."""
    batch, genome, opt = thunk()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        eng.set_vote_mode(0)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)


@pytest.mark.parametrize("name,thunk", CASES, ids=[c[0] for c in CASES])
def test_staged_vote_matches_oracle(engine_cls, oracle, name, thunk):
    """vote_staged_kernel (gcb_set_vote_mode 2: slow columns decided inside the tile's CTA) on every case."""
    batch, genome, opt = thunk()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        eng.set_vote_mode(2)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)


@pytest.mark.parametrize("name,thunk", CASES, ids=[c[0] for c in CASES])
def test_split_vote_matches_oracle(engine_cls, oracle, name, thunk):
    """vote_fast_kernel (gcb_set_vote_mode 3: one CTA per tile, slow columns queued) on every case."""
    batch, genome, opt = thunk()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        eng.set_vote_mode(3)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)


@pytest.mark.parametrize("lanes", [8, 16, 32])
@pytest.mark.parametrize("name", ["cfg2_40k", "cfg4_40k", "edge_default", "ragged_duplex_5_big", "ragged_single_4_big", "deep_1100", "low_complexity",
                                  "wide_umi_3", "cfg3_40k", "tiny_reads"])
def test_lanes_per_cluster_do_not_change_results(engine_cls, oracle, name, lanes):
    """umi_group_kernel / select_template_kernel with 8, 16 or 32 lanes per cluster."""
    batch, genome, opt = dict(CASES)[name]()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        eng.set_debug(3, lanes)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), f"{name} lanes {lanes}")


@pytest.mark.parametrize("qbytes", [1 << 14, 1 << 20])
@pytest.mark.parametrize("name", ["cfg2_40k", "ragged_duplex_5_big", "cfg4_40k", "edge_strict"])
def test_slow_queue_overflow_does_not_change_results(engine_cls, oracle, name, qbytes):
    """Default vote mode with slow-column queues far too small: what does not fit is decided inside the fast kernel."""
    batch, genome, opt = dict(CASES)[name]()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        eng.set_slow_queue_bytes(qbytes)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), f"{name} queue {qbytes}")


@pytest.mark.parametrize("mode,threads", [(2, 128), (3, 192), (4, 768)])
@pytest.mark.parametrize("name", ["cfg2_40k", "ragged_duplex_5_big", "cfg4_40k"])
def test_vote_thread_count_does_not_change_results(engine_cls, oracle, name, mode, threads):
    batch, genome, opt = dict(CASES)[name]()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        eng.set_vote_mode(mode)
        eng.set_vote_threads(threads)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), f"{name} threads {threads}")


@pytest.mark.parametrize("name", ["cfg3_40k", "ragged_duplex_5_big", "deep_1100", "cfg4_40k"])
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


def test_full_size_properties(engine_cls, oracle):
    """cfg2 at the bench size (1M pairs): properties that need no oracle run, plus an oracle check on a slice."""
    from gencore_b200 import synth
    batch, genome, _ = synth.make_fixed_batch(synth.CONFIGS["cfg2"], seed=20261019, n_pairs=1_000_000, with_qnames=False)
    opt = Options.default()
    with engine_cls(opt, 0) as eng:
        eng.set_reference(genome)
        res = eng.cluster_by_umi(batch)
        res2 = eng.cluster_by_umi(batch)
    # determinism
    assert np.array_equal(res.groups, res2.groups) and np.array_equal(res.out_payload[:res.out_bytes[0]], res2.out_payload[:res2.out_bytes[0]])
    slots = group_slots(batch, res)
    g = res.groups[slots]
    # every pair is in exactly one family and family sizes add up
    assert (res.pair_group >= 0).all()
    assert int(g["merge_reads"].sum()) == batch.n_pairs
    # no UMI pair is a duplex here, every family survives with -s 1
    assert ((g["status"] == GROUP_SSCS) | (g["status"] == GROUP_DCS)).all()
    # consensus records tile the output exactly, in order
    l = batch.reads["l_qseq"][np.maximum(g["tmpl_read"], 0)].astype(np.int64)
    sz = np.where(g["tmpl_read"] >= 0, ((l + 3) & ~3) + (((l + 1) // 2 + 3) & ~3), 0)
    offs = g["out_off"].reshape(-1)[(g["tmpl_read"] >= 0).reshape(-1)]
    assert np.array_equal(offs, np.concatenate([[0], np.cumsum(sz.reshape(-1)[sz.reshape(-1) > 0])[:-1]]))
    assert int(res.out_bytes[0]) == int(sz.sum())
    # consensus of a family whose reads all agree is the template itself (idempotence): vote again over the output
    # -> checked through the oracle on the first 2000 clusters
    c1 = 2000
    p1 = int(batch.cluster_pair_off[c1])
    from gencore_b200.abi import Batch
    sub = Batch(batch.cluster_pair_off[:c1 + 1].copy(), batch.cluster_ref[:c1].copy(), batch.cluster_flags[:c1].copy(), batch.umi[:p1].copy(),
                batch.reads[:2 * p1].copy(), batch.cigar, batch.payload[:int(batch.reads["data_off"][2 * p1])].copy(), None, None, batch.umi_prefix)
    ref = oracle.consensus(sub, genome, opt)
    nslots = group_slots(sub, ref)
    assert np.array_equal(res.cluster_n_groups[:c1], ref.cluster_n_groups)
    for name in ref.groups.dtype.names:
        assert np.array_equal(res.groups[nslots][name], ref.groups[nslots][name]), name
    assert np.array_equal(res.out_payload[:ref.out_bytes[0]], ref.out_payload[:ref.out_bytes[0]])
