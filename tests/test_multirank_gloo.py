"""The N>1 path on CPU: world_size 2 over gloo.  Each rank takes its coordinate window of the same batch
(gencore_b200.shard), runs the kernels (the SIMT-check build of the kernel source stands in for the GPU),
the genome travels by one broadcast and the Stats by one all-reduce; the ordered concatenation of the
shards must equal the single-rank answer bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, simt_lib, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import parity_cases
    from gencore_b200 import shard
    from gencore_b200.abi import Genome
    from gencore_b200.engine import ConsensusEngine
    from gencore_b200.hoststats import stats_from_result
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch, genome, opt = parity_cases.fixed_case("cfg3", 3000)
        sub, (c0, c1) = shard.shard_batch(batch, world, rank)
        sub.validate()
        if rank != 0:  # only rank 0 holds the reference before the broadcast
            genome = Genome(np.zeros_like(genome.packed4), genome.contig_off, genome.contig_len, genome.names)
        g = shard.broadcast_genome(genome, torch.device("cpu"))
        genome = Genome(g.numpy(), genome.contig_off, genome.contig_len, genome.names)
        with ConsensusEngine(opt, 0, lib_path=simt_lib) as eng:
            eng.set_reference(genome)
            res = eng.cluster_by_umi(sub)
        st = shard.reduce_stats(stats_from_result(sub, res), torch.device("cpu"))
        q.put((rank, c0, c1, res.cluster_n_groups, res.pair_group, res.out_payload[:int(res.out_bytes[0])].copy(), shard.stats_to_vector(st)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_equal_one_rank(oracle):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests", "simt_check"))
    import build as simt_build
    import parity_cases
    from gencore_b200 import shard
    from gencore_b200.hoststats import stats_from_result
    simt_lib = simt_build.build()
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, simt_lib, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=500) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    batch, genome, opt = parity_cases.fixed_case("cfg3", 3000)
    whole = oracle.consensus(batch, genome, opt)
    assert got[0][1] == 0 and got[0][2] == got[1][1] and got[1][2] == batch.n_clusters and 0 < got[0][2] < batch.n_clusters
    parts = [{"cluster_n_groups": g[3], "pair_group": g[4], "out_payload": g[5]} for g in got]
    cat = shard.concat_results(parts)
    np.testing.assert_array_equal(cat["cluster_n_groups"], whole.cluster_n_groups)
    np.testing.assert_array_equal(cat["pair_group"], whole.pair_group)
    np.testing.assert_array_equal(cat["out_payload"], whole.out_payload[:int(whole.out_bytes[0])])
    want = shard.stats_to_vector(stats_from_result(batch, whole))
    for g in got:  # every rank holds the reduced Stats
        np.testing.assert_array_equal(g[6], want)


def test_window_bounds_are_balanced_and_cover():
    import parity_cases
    from gencore_b200 import shard
    batch, _, _ = parity_cases.fixed_case("cfg2", 4000)
    for world in (1, 2, 4, 8):
        b = shard.window_bounds(batch, world)
        assert b[0] == 0 and b[-1] == batch.n_clusters and (np.diff(b) >= 0).all()
        slab = batch.cluster_slab_bounds()
        sizes = slab[b[1:]] - slab[b[:-1]]
        assert sizes.sum() == len(batch.payload)
        assert sizes.max() - sizes.min() <= 2 * batch.max_cluster_bytes() + 16
