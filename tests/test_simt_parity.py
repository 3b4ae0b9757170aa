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
N_COUNTERS = 8  # ring tiles, generic tiles, voted columns, slow columns, uniform / non-uniform family sides, clusters selected in registers, arena wraps


@pytest.fixture(scope="module")
def simt_lib():
    import build as simt_build
    return simt_build.build()


def run(simt_lib, batch, genome, opt, setup=None):
    from gencore_b200.engine import ConsensusEngine
    cnt = (ctypes.c_int64 * N_COUNTERS)()
    with ConsensusEngine(opt, 0, lib_path=simt_lib) as eng:
        eng.set_reference(genome)
        if setup:
            setup(eng)
        eng.lib.gcb_simt_counters(cnt, 1)
        res = eng.cluster_by_umi(batch)
        eng.lib.gcb_simt_counters(cnt, 1)
    return res, list(cnt)


@pytest.mark.parametrize("name,thunk", CASES, ids=[c[0] for c in CASES])
def test_kernels_match_oracle_under_simt_check(simt_lib, oracle, name, thunk):
    batch, genome, opt = thunk()
    res, cnt = run(simt_lib, batch, genome, opt)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)
    ring, generic, cols, slow, uni, nonuni, in_regs, reuse = cnt
    COVERAGE[name] = cnt
    if name in ("deep_1100", "low_complexity"):
        assert generic > 0, "the >1000-pair clusters must take the generic kernel"
    elif batch.n_pairs > 0:
        assert ring > 0 and generic == 0, (ring, generic)
    if name.startswith(("cfg1", "cfg2", "cfg3")):
        assert 0 < slow < 0.2 * cols, "the fixed-length shapes must mostly take the fast columns"
        assert uni > 0 and nonuni == 0, "fixed-length families are uniform"
    if name.startswith(("ragged", "cfg5")) and "_3" not in name:
        assert nonuni > 0
    if name == "cfg2_6000":
        assert reuse > 0, "every CTA must go around its ring"
    if name.startswith(("cfg1", "cfg2", "cfg3")):
        assert in_regs > 0.6 * batch.n_clusters, "the usual clusters must be selected in registers"


def test_both_column_paths_are_exercised():
    assert COVERAGE, "runs after the parametrised cases"
    assert sum(v[2] - v[3] for v in COVERAGE.values()) > 0 and sum(v[3] for v in COVERAGE.values()) > 0
    print({k: v for k, v in COVERAGE.items()})


@pytest.mark.parametrize("name,chunk", [(n, c) for n in ["cfg3_1500", "ragged_duplex_1", "golden_cfg4_600", "cfg5_1500"] for c in [1 << 12, 1 << 15, 1 << 17]] +
                         [("deep_1100", 1 << 15)])
def test_pipeline_chunks_do_not_change_results(simt_lib, oracle, name, chunk):
    """gcb_consensus_batch splits a batch into chunks of clusters that overlap copies and kernels; force many small chunks."""
    batch, genome, opt = dict(CASES)[name]()
    res, _ = run(simt_lib, batch, genome, opt, lambda eng: eng.set_chunk_bytes(chunk))
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), f"{name} chunk {chunk}")


LIGHT_CASES = [c for c in CASES if c[0] in ("golden_cfg1_600", "golden_cfg2_600", "golden_cfg3_600", "golden_cfg4_600", "golden_ragged_duplex",
                                           "edge_default", "edge_strict", "ragged_none_0", "ragged_single_1", "ragged_duplex_2", "ragged_duplex_3",
                                           "low_complexity", "no_reference", "empty", "tiny_reads", "cfg2_1500", "cfg3_1500",
                                           "cfg4_1500", "cfg5_1500", "wide_umi_3", "deep30_noisy")]


@pytest.mark.parametrize("name,thunk", LIGHT_CASES, ids=[c[0] for c in LIGHT_CASES])
def test_generic_vote_matches_oracle_under_simt_check(simt_lib, oracle, name, thunk):
    """score_vote_kernel (the kernel without size limits that takes the tiles the ring cannot stage) on every tile."""
    batch, genome, opt = thunk()
    res, cnt = run(simt_lib, batch, genome, opt, lambda eng: eng.set_debug(5, 1))
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)
    assert cnt[0] == 0 and (cnt[1] > 0 or batch.n_pairs == 0), cnt


@pytest.mark.parametrize("name,thunk", LIGHT_CASES, ids=[c[0] for c in LIGHT_CASES])
def test_small_tile_window_matches_oracle_under_simt_check(simt_lib, oracle, name, thunk):
    """The vote over 16 KB payload windows instead of 32 KB ones (twice the tiles, the voter warps in three groups): same bytes."""
    batch, genome, opt = thunk()
    res, _ = run(simt_lib, batch, genome, opt, lambda eng: eng.set_debug(2, 14))
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)


@pytest.mark.parametrize("qbytes", [4096, 65536])
@pytest.mark.parametrize("name", ["cfg2_1500", "ragged_duplex_2", "golden_cfg4_600", "edge_strict", "deep30_noisy"])
def test_slow_queue_overflow_does_not_change_results(simt_lib, oracle, name, qbytes):
    """A slow-column queue far too small: the tiles whose columns do not fit are redone by the generic kernel."""
    batch, genome, opt = dict(CASES)[name]()
    res, cnt = run(simt_lib, batch, genome, opt, lambda eng: eng.set_slow_queue_bytes(qbytes))
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), f"{name} queue {qbytes}")
    if qbytes == 4096 and name != "golden_cfg4_600":
        assert cnt[1] > 0, "some tile must have been handed to the generic kernel"


@pytest.mark.parametrize("name,lanes", [(n, l) for n in ["golden_cfg2_600", "golden_cfg4_600", "edge_default", "ragged_duplex_2", "ragged_single_1",
                                                         "low_complexity", "wide_umi_3", "cfg3_1500", "tiny_reads", "cfg5_1500", "cfg3_crowded"] for l in [8, 16, 32]] +
                         [("deep_1100", 8)])
def test_lanes_per_cluster_do_not_change_results(simt_lib, oracle, name, lanes):
    """umi_group_kernel / select_template_kernel with 8, 16 or 32 lanes per cluster (groups of a warp work on different clusters)."""
    batch, genome, opt = dict(CASES)[name]()
    res, _ = run(simt_lib, batch, genome, opt, lambda eng: eng.set_debug(3, lanes))
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), f"{name} lanes {lanes}")


def assert_stats_equal(a, b, what):
    import dataclasses
    import numpy as np
    for f in dataclasses.fields(a):
        x, y = getattr(a, f.name), getattr(b, f.name)
        assert np.array_equal(x, y), f"{what}: {f.name} {x} vs {y}"


@pytest.mark.parametrize("name", ["golden_cfg2_600", "golden_cfg3_600", "golden_cfg4_600", "edge_default", "edge_strict", "ragged_duplex_2",
                                  "deep_1100", "empty", "cfg3_1500"])
def test_cluster_stats_kernel_matches_the_host_replay(simt_lib, name):
    """cluster_stats_kernel: the Stats side effects of clusterByUMI (cluster.cpp:102-186) summed on the device; hoststats is the
    numpy replay that tests/parity.py::assert_matches_reference pins against the reference's own Stats objects."""
    from gencore_b200.engine import ConsensusEngine
    from gencore_b200.hoststats import stats_from_result
    batch, genome, opt = dict(CASES)[name]()
    with ConsensusEngine(opt, 0, lib_path=simt_lib) as eng:
        eng.set_reference(genome)
        eng.set_chunk_bytes(1 << 15)  # several chunks: the counters are sums
        res = eng.cluster_by_umi(batch)
        got = eng.cluster_stats()
        assert_stats_equal(got, stats_from_result(batch, res), name)
        # the reset worked, and two batches add up
        eng.cluster_by_umi(batch)
        eng.cluster_by_umi(batch)
        twice = eng.cluster_stats()
    assert twice.pre_molecule == 2 * got.pre_molecule and twice.post_sscs == 2 * got.post_sscs and (twice.pre_hist == 2 * got.pre_hist).all()


@pytest.mark.parametrize("what", ["slab_misaligned", "pair_off_not_monotone", "data_off_unaligned", "cigar_out_of_range", "record_outside_slab"])
def test_malformed_batches_are_refused(simt_lib, what):
    """A batch that breaks a documented precondition comes back as GCB_ERR_MALFORMED (no fault, no hang)."""
    import numpy as np
    from gencore_b200.abi import GCB_ERR_MALFORMED
    from gencore_b200.engine import ConsensusEngine, EngineError
    batch, genome, opt = dict(CASES)["cfg2_1500"]()
    reads = batch.reads.copy()
    pair_off = batch.cluster_pair_off.copy()
    c = batch.n_clusters // 2
    p = int(pair_off[c])
    if what == "slab_misaligned":
        reads["data_off"][2 * p:] += 4
    elif what == "pair_off_not_monotone":
        pair_off[c] = pair_off[c + 1] + 3
    elif what == "data_off_unaligned":
        reads["data_off"][2 * p + 1] += 1
    elif what == "cigar_out_of_range":
        reads["cigar_off"][2 * p + 1] = len(batch.cigar) + 5
    else:
        reads["data_off"][2 * p + 3] = len(batch.payload) - 16
    import dataclasses
    bad = dataclasses.replace(batch, reads=reads, cluster_pair_off=pair_off)
    with ConsensusEngine(opt, 0, lib_path=simt_lib) as eng:
        eng.set_reference(genome)
        with pytest.raises(EngineError) as ei:
            eng.cluster_by_umi(bad)
        assert ei.value.code == GCB_ERR_MALFORMED
        # the context stays usable
        res = eng.cluster_by_umi(batch)
    assert int(res.out_bytes[0]) > 0


def test_duplex_strand_forms_match_is_duplex(simt_lib, oracle):
    """duplex_kernel compares precomputed strand forms (device_common.cuh umi_strand_forms) instead of splitting both UMIs per
    candidate: the same verdict as the literal split and as the oracle's Cluster::isDuplex on the reference's vectors, on
    split() corner cases and on random strings."""
    import random
    import numpy as np
    from gencore_b200.abi import encode_umi
    from test_oracle_kat import ISDUPLEX_EXTRA, ISDUPLEX_KAT
    lib = ctypes.CDLL(simt_lib)
    lib.gcb_simt_is_duplex.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    rng = random.Random(5)
    pairs = [(a.replace('B', 'C'), b.replace('B', 'C')) for a, b in [(a, b) for a, b, _ in ISDUPLEX_KAT] + list(ISDUPLEX_EXTRA)]  # ('B' has no code)
    for _ in range(3000):
        a = "".join(rng.choice("ACGT__") for _ in range(rng.randint(0, 9)))
        if rng.random() < 0.5 and "_" in a:  # a true partner, or a near one
            parts = a.split("_")
            b = "_".join(reversed(parts)) if rng.random() < 0.7 else "_".join(parts[1:] + parts[:1])
        else:
            b = "".join(rng.choice("ACGT__") for _ in range(rng.randint(0, 9)))
        pairs.append((a, b))
    n_true = 0
    for nw in (1, 2, 4):
        for a, b in pairs:
            ca, cb = np.ascontiguousarray(encode_umi(a, nw)), np.ascontiguousarray(encode_umi(b, nw))
            for x, y, sx, sy in ((ca, cb, a, b), (cb, ca, b, a)):
                got = lib.gcb_simt_is_duplex(x.ctypes.data, y.ctypes.data, nw)
                want = bool(oracle.is_duplex(sx, sy))
                assert got == (3 if want else 0), (sx, sy, got, want)
                n_true += want
    assert n_true > 100
