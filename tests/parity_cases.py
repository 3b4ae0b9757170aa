"""The batches every implementation of the hot path is compared on (SIMT-check build on CPU, CUDA build on
the B200): golden fixtures from the reference, the hand-built quirk batch under every option set, random
ragged batches, >1000-pair families, the fixed-length BASELINE.json shapes, and degenerate inputs."""
from __future__ import annotations

import dataclasses
import glob
import os

import numpy as np

import cases
from gencore_b200 import synth
from gencore_b200.abi import READ_DESC, Batch, Options
from parity import load_golden

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def golden_case(path):
    batch, genome, opt, *_ = load_golden(path)
    return batch, genome, opt


def edge_case(optname):
    batch, genome, _ = cases.edge_batch()
    return batch, genome, cases.OPTION_SETS[optname]


def ragged_case(seed, umi, n_clusters=60):
    batch, genome, _ = synth.make_ragged_batch(100 + seed, n_clusters=n_clusters, umi=umi, err=0.02 if seed % 2 else 0.005)
    return batch, genome, list(cases.OPTION_SETS.values())[seed % len(cases.OPTION_SETS)]


def fixed_case(name, n_pairs, contig_len=200_000, **over):
    cfg = synth.CONFIGS[name]
    small = dataclasses.replace(cfg, contig_len=contig_len, n_contigs=min(cfg.n_contigs, 2), **over)
    batch, genome, _ = synth.make_batch(small, seed=20261017, n_pairs=n_pairs, with_qnames=False)
    return batch, genome, Options.default(cluster_size_req=cfg.supporting_reads)


def noisy_deep_case(n_pairs, depth, err, contig_len=200_000):
    """Deep duplex families with many slow columns: the tile's slow-column list is long and other warps help to decide it."""
    return fixed_case("cfg3", n_pairs, contig_len, depth=float(depth), err=err)


def deep_case():
    batch, genome, _ = cases.deep_batch()
    return batch, genome, Options.default()


def low_complexity_case():
    batch, genome, _ = cases.low_complexity_batch()
    return batch, genome, Options.default()


def no_reference_case():
    batch, _genome, _ = synth.make_ragged_batch(5, n_clusters=40, umi="single")
    return batch, None, Options.default()


def wide_umi_case(words):
    """A duplex batch whose UMI array is widened to `words` u64 per pair (zero fields = past the end of the string):
    the 3- and 4-word instantiations of the grouping kernel."""
    batch, genome, opt = ragged_case(2, "duplex", 30)
    umi = np.zeros((batch.n_pairs, words), np.uint64)
    umi[:, :batch.umi.shape[1]] = batch.umi
    return dataclasses.replace(batch, umi=umi), genome, opt


def empty_case():
    batch = Batch(np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.uint8), np.zeros((0, 1), np.uint64),
                  np.zeros(0, READ_DESC), np.zeros(1, np.uint32), np.zeros(16, np.uint8), [], np.zeros(0, np.uint8), "")
    return batch, None, Options.default()


def tiny_reads_case():
    """Reads of 1..9 bases, odd lengths, mates missing: record padding and the nibble tail."""
    from gencore_b200.synth import BatchBuilder, SynthCluster, SynthPair, SynthRead
    rng = np.random.Generator(np.random.PCG64(3))
    contigs, genome = synth.random_genome(rng, [400])
    bb = BatchBuilder(1, "UMI")
    for c in range(12):
        prs = []
        for i in range(1 + c % 4):
            l = 1 + (c + i) % 9
            pos = 10 + 20 * c
            left = SynthRead(pos, f"{l}M", bytes(contigs[0][pos:pos + l]), rng.integers(2, 41, l).astype(np.uint8), 12)
            right = None
            if (c + i) % 3:
                right = SynthRead(pos + 3, f"{l}M", bytes(contigs[0][pos + 3:pos + 3 + l]), rng.integers(2, 41, l).astype(np.uint8), -12)
            prs.append(SynthPair(b"t%d_%d:UMI_ACGT" % (c, i), "ACGT", left, right))
        bb.add(SynthCluster(0, prs))
    return bb.build(), genome, Options.default()


def small_cases():
    """(id, thunk) for the cases cheap enough for the CPU SIMT interpreter."""
    out = [("golden_" + os.path.basename(p)[:-4], (lambda p=p: golden_case(p))) for p in GOLDEN]
    out += [("edge_" + n, (lambda n=n: edge_case(n))) for n in cases.OPTION_SETS]
    out += [(f"ragged_{umi}_{seed}", (lambda s=seed, u=umi: ragged_case(s, u, 30))) for seed in range(4) for umi in ("none", "single", "duplex")]
    out += [("deep_1100", deep_case), ("low_complexity", low_complexity_case), ("no_reference", no_reference_case),
            ("empty", empty_case), ("tiny_reads", tiny_reads_case),
            ("wide_umi_3", lambda: wide_umi_case(3)), ("wide_umi_4", lambda: wide_umi_case(4))]
    out += [(f"{n}_1500", (lambda n=n: fixed_case(n, 1500))) for n in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5")]
    out += [("cfg2_6000", lambda: fixed_case("cfg2", 6000)),           # many tiles per CTA: the ring's stages are used again and again
            ("deep30_noisy", lambda: noisy_deep_case(1500, 30, 0.01)),
            ("deep300_noisy", lambda: noisy_deep_case(1500, 300, 0.003)),  # clusters of ~130 KB: one tile fills the ring kernel's arena
            # many molecules per coordinate pair: more strand families per cluster than duplex_kernel gives lanes to one
            ("cfg3_crowded", lambda: fixed_case("cfg3", 1500, contig_len=245, depth=2.0, err=0.01, insert_sigma=0.5))]
    return out


def gpu_cases():
    out = small_cases()
    out += [(f"ragged_{umi}_{seed}_big", (lambda s=seed, u=umi: ragged_case(s, u, 400))) for seed in range(4, 8) for umi in ("none", "single", "duplex")]
    out += [(f"{n}_40k", (lambda n=n: fixed_case(n, 40_000, 2_000_000))) for n in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5")]
    out += [("deep50_noisy_40k", lambda: noisy_deep_case(40_000, 50, 0.01, 2_000_000)),
            ("deep400_40k", lambda: noisy_deep_case(40_000, 400, 0.002, 2_000_000))]
    return out
