"""The kernel source itself (gencore_b200/csrc/*.cu*), compiled by g++ against the SIMT interpreter in
tests/simt_check/, executed on the CPU and compared bit-for-bit with the oracle.  This is a check of the
kernels' LOGIC on a box without a GPU; the same comparisons run on the real library in test_gpu_parity.py."""
import ctypes
import os
import sys

import pytest

import parity_cases
from parity import assert_results_equal

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt_check"))

CASES = parity_cases.small_cases()
COVERAGE = {}


@pytest.fixture(scope="module")
def simt_lib():
    import build as simt_build
    return simt_build.build()


@pytest.mark.parametrize("name,thunk", CASES, ids=[c[0] for c in CASES])
def test_kernels_match_oracle_under_simt_check(simt_lib, oracle, name, thunk):
    from gencore_b200.engine import ConsensusEngine
    batch, genome, opt = thunk()
    cnt = (ctypes.c_int64 * 6)()
    with ConsensusEngine(opt, 0, lib_path=simt_lib) as eng:
        eng.set_reference(genome)
        eng.lib.gcb_simt_counters(cnt, 1)
        res = eng.cluster_by_umi(batch)
        eng.lib.gcb_simt_counters(cnt, 1)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)
    tiled, generic, cols, slow, uni, nonuni = list(cnt)
    COVERAGE[name] = (tiled, generic, cols, slow, uni, nonuni)
    if name in ("deep_1100", "low_complexity"):
        assert generic > 0, "the >1000-pair clusters must take the generic kernel"
    elif batch.n_pairs > 0:
        assert tiled > 0 and generic == 0, (tiled, generic)
    if name.startswith(("cfg1", "cfg2", "cfg3")):
        assert 0 < slow < 0.2 * cols, "the fixed-length shapes must mostly take the fast columns"
        assert uni > 0 and nonuni == 0, "fixed-length families are uniform"
    if name.startswith("ragged") and "_3" not in name:
        assert nonuni > 0


def test_both_column_paths_are_exercised():
    assert COVERAGE, "runs after the parametrised cases"
    assert sum(v[2] - v[3] for v in COVERAGE.values()) > 0 and sum(v[3] for v in COVERAGE.values()) > 0
    print({k: v for k, v in COVERAGE.items()})


@pytest.mark.parametrize("name,chunk", [(n, c) for n in ["cfg3_1500", "ragged_duplex_1", "golden_cfg4_600"] for c in [1 << 12, 1 << 15, 1 << 17]] +
                         [("deep_1100", 1 << 15)])
def test_pipeline_chunks_do_not_change_results(simt_lib, oracle, name, chunk):
    """gcb_consensus_batch splits a batch into chunks of clusters that overlap copies and kernels; force many small chunks."""
    from gencore_b200.engine import ConsensusEngine
    batch, genome, opt = dict(CASES)[name]()
    with ConsensusEngine(opt, 0, lib_path=simt_lib) as eng:
        eng.set_reference(genome)
        eng.set_chunk_bytes(chunk)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), f"{name} chunk {chunk}")


PIPE_CASES = [c for c in CASES if c[0] in ("golden_cfg1_600", "golden_cfg2_600", "golden_cfg3_600", "golden_cfg4_600", "golden_ragged_duplex",
                                          "edge_default", "edge_strict", "ragged_none_0", "ragged_single_1", "ragged_duplex_2", "ragged_duplex_3",
                                          "deep_1100", "low_complexity", "no_reference", "empty", "tiny_reads", "cfg2_1500", "cfg3_1500",
                                          "cfg4_1500", "wide_umi_3")]


LIGHT_CASES = [c for c in PIPE_CASES if c[0] != "deep_1100"]  # (the >1000-pair cluster takes the generic kernel in every mode: once is enough)


@pytest.mark.parametrize("name,thunk", LIGHT_CASES, ids=[c[0] for c in LIGHT_CASES])
def test_pipelined_vote_matches_oracle_under_simt_check(simt_lib, oracle, name, thunk):
    """vote_pipe_kernel (persistent CTA, ring of staged tiles, producer thread + consumer warps) gives the same bytes."""
    from gencore_b200.engine import ConsensusEngine
    batch, genome, opt = thunk()
    cnt = (ctypes.c_int64 * 6)()
    with ConsensusEngine(opt, 0, lib_path=simt_lib) as eng:
        eng.set_reference(genome)
        eng.set_vote_mode(1)
        eng.lib.gcb_simt_counters(cnt, 1)
        res = eng.cluster_by_umi(batch)
        eng.lib.gcb_simt_counters(cnt, 1)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)
    if name.startswith(("cfg2", "cfg3", "golden_cfg")):
        assert cnt[0] > 0, "tiles must take the pipeline"


@pytest.mark.parametrize("name,thunk", PIPE_CASES, ids=[c[0] for c in PIPE_CASES])
def test_tiled_vote_matches_oracle_under_simt_check(simt_lib, oracle, name, thunk):
    """vote_tiled_kernel (vote mode 0: every CTA computes its tile's family-side table itself) gives the same bytes."""
    from gencore_b200.engine import ConsensusEngine
    batch, genome, opt = thunk()
    with ConsensusEngine(opt, 0, lib_path=simt_lib) as eng:
        eng.set_reference(genome)
        eng.set_vote_mode(0)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)


@pytest.mark.parametrize("mode,threads", [(2, 64), (3, 192), (4, 768)])
@pytest.mark.parametrize("name", ["cfg2_1500", "ragged_duplex_2", "golden_cfg4_600"])
def test_vote_thread_count_does_not_change_results(simt_lib, oracle, name, mode, threads):
    from gencore_b200.engine import ConsensusEngine
    batch, genome, opt = dict(CASES)[name]()
    with ConsensusEngine(opt, 0, lib_path=simt_lib) as eng:
        eng.set_reference(genome)
        eng.set_vote_mode(mode)
        eng.set_vote_threads(threads)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), f"{name} threads {threads}")


@pytest.mark.parametrize("name,thunk", LIGHT_CASES, ids=[c[0] for c in LIGHT_CASES])
def test_staged_vote_matches_oracle_under_simt_check(simt_lib, oracle, name, thunk):
    """vote_staged_kernel (vote mode 2: slow columns decided inside the tile's CTA, one thread per column) gives the same bytes."""
    from gencore_b200.engine import ConsensusEngine
    batch, genome, opt = thunk()
    with ConsensusEngine(opt, 0, lib_path=simt_lib) as eng:
        eng.set_reference(genome)
        eng.set_vote_mode(2)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)


@pytest.mark.parametrize("qbytes", [4096, 65536])
@pytest.mark.parametrize("name", ["cfg2_1500", "ragged_duplex_2", "golden_cfg4_600", "edge_strict"])
def test_slow_queue_overflow_does_not_change_results(simt_lib, oracle, name, qbytes):
    """Vote mode 3 with slow-column queues far too small: what does not fit is decided inside the fast kernel."""
    from gencore_b200.engine import ConsensusEngine
    batch, genome, opt = dict(CASES)[name]()
    with ConsensusEngine(opt, 0, lib_path=simt_lib) as eng:
        eng.set_reference(genome)
        eng.set_slow_queue_bytes(qbytes)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), f"{name} queue {qbytes}")


@pytest.mark.parametrize("name,thunk", LIGHT_CASES, ids=[c[0] for c in LIGHT_CASES])
def test_split_vote_matches_oracle_under_simt_check(simt_lib, oracle, name, thunk):
    """vote_fast_kernel (vote mode 3: one CTA per tile, slow columns queued for slow_columns_kernel) gives the same bytes."""
    from gencore_b200.engine import ConsensusEngine
    batch, genome, opt = thunk()
    with ConsensusEngine(opt, 0, lib_path=simt_lib) as eng:
        eng.set_reference(genome)
        eng.set_vote_mode(3)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)


@pytest.mark.parametrize("name,lanes", [(n, l) for n in ["golden_cfg2_600", "golden_cfg4_600", "edge_default", "ragged_duplex_2", "ragged_single_1",
                                                         "low_complexity", "wide_umi_3", "cfg3_1500", "tiny_reads"] for l in [8, 16, 32]] +
                         [("deep_1100", 8)])
def test_lanes_per_cluster_do_not_change_results(simt_lib, oracle, name, lanes):
    """umi_group_kernel / select_template_kernel with 8, 16 or 32 lanes per cluster (groups of a warp work on different clusters)."""
    from gencore_b200.engine import ConsensusEngine
    batch, genome, opt = dict(CASES)[name]()
    with ConsensusEngine(opt, 0, lib_path=simt_lib) as eng:
        eng.set_reference(genome)
        eng.set_debug(3, lanes)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), f"{name} lanes {lanes}")


UNIT_CASES = [c for c in CASES if c[0] in ("golden_cfg1_600", "golden_cfg2_600", "golden_cfg3_600", "golden_cfg4_600", "golden_ragged_duplex",
                                          "edge_default", "edge_strict", "edge_loose", "ragged_none_0", "ragged_single_1", "ragged_duplex_2",
                                          "ragged_duplex_3", "no_reference", "empty", "tiny_reads", "cfg2_1500", "cfg3_1500", "cfg4_1500",
                                          "wide_umi_3", "ragged_single_3", "ragged_none_2")]


@pytest.mark.parametrize("name,thunk", UNIT_CASES, ids=[c[0] for c in UNIT_CASES])
def test_ring_with_two_units_per_lane_matches_oracle(simt_lib, oracle, name, thunk):
    """vote_ring_kernel with thirty-two columns per lane (two units of sixteen) gives the same bytes."""
    from gencore_b200.engine import ConsensusEngine
    batch, genome, opt = thunk()
    with ConsensusEngine(opt, 0, lib_path=simt_lib) as eng:
        eng.set_reference(genome)
        eng.set_debug(4, 2)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)
