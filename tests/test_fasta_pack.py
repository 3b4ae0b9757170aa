"""gcb_pack_fasta (k_fasta_pack.cuh) against the reference's own FastaReader: golden vectors generated from the unmodified
reference (tests/golden_fasta/fasta.npz, tests/make_golden_fasta.py), a literal restatement of readNext (fasta_cases.restate_fasta),
the kernels under the SIMT interpreter on the CPU box and the CUDA library on the B200."""
import os
import sys

import numpy as np
import pytest

from fasta_cases import MALFORMED, fasta_cases, restate_fasta

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "simt_check"))
CASES = fasta_cases()
GOLDEN = np.load(os.path.join(HERE, "golden_fasta", "fasta.npz"), allow_pickle=False)


def golden(name):
    ids = [bytes(x) for x in GOLDEN[name + "/ids"]]
    sizes, offs, packed = GOLDEN[name + "/sizes"], GOLDEN[name + "/offs"], GOLDEN[name + "/packed"]
    return ids, [int(s) for s in sizes], [packed[int(o):int(o) + (int(s) + 1) // 2] for o, s in zip(offs, sizes)]


def assert_genome(genome, ids, sizes, packed, what):
    assert [n.encode("latin-1") for n in genome.names] == ids, what
    assert [int(x) for x in genome.contig_len] == sizes, what
    for i, p in enumerate(packed):
        o = int(genome.contig_off[i])
        assert o % 16 == 0
        assert np.array_equal(genome.packed4[o:o + len(p)], p), f"{what}: contig {i}"


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_text_is_current(name):
    assert bytes(GOLDEN[name + "/text"]) == CASES[name], "tests/golden_fasta/fasta.npz is stale: run tests/make_golden_fasta.py"


@pytest.mark.parametrize("name", sorted(CASES))
def test_restatement_matches_reference_golden(name):
    ids, sizes, packed = restate_fasta(CASES[name])
    gids, gsizes, gpacked = golden(name)
    assert ids == gids and sizes == gsizes
    for a, b in zip(packed, gpacked):
        assert np.array_equal(a, b)


@pytest.fixture(scope="module")
def simt_engine():
    import build as simt_build
    from gencore_b200.engine import ConsensusEngine
    with ConsensusEngine(None, 0, lib_path=simt_build.build()) as eng:
        yield eng


@pytest.mark.parametrize("name", sorted(CASES))
def test_kernels_match_reference_under_simt_check(simt_engine, name):
    genome = simt_engine.pack_fasta(CASES[name], max_contigs=128)
    assert_genome(genome, *golden(name), name)


@pytest.mark.parametrize("name", sorted(MALFORMED))
def test_headers_the_reference_misreads_are_refused(simt_engine, name):
    from gencore_b200.abi import GCB_ERR_MALFORMED
    from gencore_b200.engine import EngineError
    with pytest.raises(EngineError) as ei:
        simt_engine.pack_fasta(MALFORMED[name])
    assert ei.value.code == GCB_ERR_MALFORMED


def test_too_many_contigs_is_a_capacity_error(simt_engine):
    from gencore_b200.abi import GCB_ERR_CAPACITY
    from gencore_b200.engine import EngineError
    with pytest.raises(EngineError) as ei:
        simt_engine.pack_fasta(CASES["many_contigs"], max_contigs=8)
    assert ei.value.code == GCB_ERR_CAPACITY


def test_packed_genome_equals_the_generator_s(simt_engine):
    """The synthetic genomes of the parity tests (Genome.from_bases) and the FASTA the reference arm reads are the same bytes."""
    from gencore_b200 import synth
    from oracle.pyoracle import write_fasta
    import tempfile
    rng = np.random.Generator(np.random.PCG64(5))
    contigs, genome = synth.random_genome(rng, [10_001, 777, 4096])
    with tempfile.NamedTemporaryFile(suffix=".fa") as f:
        write_fasta(f.name, contigs, genome.names)
        text = open(f.name, "rb").read()
    got = simt_engine.pack_fasta(text)
    assert got.names == genome.names and np.array_equal(got.contig_len, genome.contig_len) and np.array_equal(got.contig_off, genome.contig_off)
    assert np.array_equal(got.packed4[:len(genome.packed4)], genome.packed4)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_reference(name):
    from gencore_b200.engine import ConsensusEngine
    with ConsensusEngine(None, 0) as eng:
        genome = eng.pack_fasta(CASES[name], max_contigs=128)
    assert_genome(genome, *golden(name), name)


@pytest.mark.gpu
def test_cuda_large_fasta_matches_generator():
    """A 60 Mb two-contig FASTA (the size of the bench's reference): packed on the GPU = packed by the generator."""
    from gencore_b200 import synth
    from gencore_b200.engine import ConsensusEngine
    rng = np.random.Generator(np.random.PCG64(9))
    contigs, genome = synth.random_genome(rng, [50_000_000, 10_000_001])
    parts = []
    for name, c in zip(genome.names, contigs):
        parts.append(b">" + name.encode() + b" synthetic\n")
        body = np.empty(len(c) + (len(c) + 59) // 60, np.uint8)
        idx = np.arange(len(c)) + np.arange(len(c)) // 60
        body[:] = 10
        body[idx] = c
        parts.append(body.tobytes())
    text = b"".join(parts)
    with ConsensusEngine(None, 0) as eng:
        got = eng.pack_fasta(text)
    assert got.names == genome.names and np.array_equal(got.contig_len, genome.contig_len)
    for i in range(2):
        o, g, nb = int(got.contig_off[i]), int(genome.contig_off[i]), (int(genome.contig_len[i]) + 1) // 2
        assert np.array_equal(got.packed4[o:o + nb], genome.packed4[g:g + nb])
