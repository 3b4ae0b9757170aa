"""Sorted BAM in, consensus BAM out: gencore_b200/bin/gencore_b200 (the host pipeline around the C ABI) against the
UNMODIFIED reference binary (oracle/_ref/gencore) on the same synthetic BAM + FASTA — the comparison BASELINE.json's
metric names ("output BAM bit-exact vs reference").  On the CPU box the engine behind the ABI is the SIMT-check build of
the kernel source (test infrastructure); the `gpu` test loads the CUDA library."""
import dataclasses
import os
import subprocess
import sys

import pytest

import bamfile
from gencore_b200 import build as gbuild
from gencore_b200 import synth
from oracle import pyoracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt_check"))

CONFIGS = [
    ("cfg1", 3000, []),                 # no UMI, -s 1
    ("cfg2", 6000, []),                 # 8-nt UMI, -d 1, crosses the 10 000-read tick (Q18: -d for flushed clusters, 0 at the end)
    ("cfg3", 7000, []),                 # duplex UMIs
    ("cfg3", 2500, ["-s", "2", "-D", "1"]),
    ("cfg2", 2500, ["--no_duplex", "-d", "0"]),
    ("cfg4", 3000, ["-s", "2", "-x"]),
]


def _make_inputs(tmp_path, name, n_pairs):
    cfg = synth.CONFIGS[name]
    small = dataclasses.replace(cfg, contig_len=120_000, n_contigs=min(cfg.n_contigs, 2))
    batch, genome, contigs = synth.make_fixed_batch(small, seed=20261017 + n_pairs, n_pairs=n_pairs, with_qnames=True)
    fa, bam = str(tmp_path / "ref.fa"), str(tmp_path / "in.bam")
    bamfile.genome_to_fasta(fa, contigs, genome.names)
    n = bamfile.batch_to_bam(bam, batch, genome)
    return fa, bam, n


def _make_ragged_inputs(tmp_path, seed, umi):
    """Reads of 100-250 bases with soft clips and indels, pairs without a mate (written mate-unmapped: the pass-through
    of gencore.cpp:307-309), plus unmapped and secondary records the reference drops (Q19)."""
    batch, genome, contigs = synth.make_ragged_batch(seed, n_clusters=500, umi=umi)
    fa, bam = str(tmp_path / "ref.fa"), str(tmp_path / "in.bam")
    bamfile.genome_to_fasta(fa, contigs, genome.names)
    n = bamfile.batch_to_bam(bam, batch, genome, extras=True)
    return fa, bam, n


def _run_both(tmp_path, fa, bam, flags, engine_lib):
    if not pyoracle.reference_available():
        pytest.skip("oracle/_ref/gencore is not built")
    ref_out, my_out = str(tmp_path / "ref.bam"), str(tmp_path / "b200.bam")
    r = subprocess.run([pyoracle.REF_BIN, "-i", bam, "-o", ref_out, "-r", fa, "-j", str(tmp_path / "r.json"), "-h", str(tmp_path / "r.html")] + flags,
                       capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    cli = gbuild.build_cli()
    m = subprocess.run([cli, "-i", bam, "-o", my_out, "-r", fa, "--engine", engine_lib] + flags, capture_output=True, text=True, cwd=str(tmp_path))
    assert m.returncode == 0, m.stderr[-2000:]
    return bamfile.assert_same_bam(ref_out, my_out)


@pytest.mark.parametrize("name,n_pairs,flags", CONFIGS, ids=[f"{c[0]}_{c[1]}_{'_'.join(c[2]) or 'default'}" for c in CONFIGS])
def test_bam_pipeline_matches_reference_binary_simt(tmp_path, name, n_pairs, flags):
    import build as simt_build
    fa, bam, n_in = _make_inputs(tmp_path, name, n_pairs)
    n_out = _run_both(tmp_path, fa, bam, flags, simt_build.build())
    assert 0 < n_out < n_in


@pytest.mark.parametrize("seed,umi,flags", [(11, "none", []), (12, "single", []), (13, "duplex", []), (14, "duplex", ["-s", "2", "-c", "8"]),
                                            (15, "single", ["-u", "UMI", "--high_qual", "35", "--low_qual", "10"])])
def test_bam_pipeline_ragged_matches_reference_binary_simt(tmp_path, seed, umi, flags):
    import build as simt_build
    fa, bam, n_in = _make_ragged_inputs(tmp_path, seed, umi)
    n_out = _run_both(tmp_path, fa, bam, flags, simt_build.build())
    assert 0 < n_out < n_in


@pytest.mark.gpu
@pytest.mark.parametrize("name,n_pairs,flags", [("cfg2", 60_000, []), ("cfg3", 40_000, []), ("cfg1", 20_000, ["-s", "2"])])
def test_bam_pipeline_matches_reference_binary_cuda(tmp_path, name, n_pairs, flags):
    fa, bam, n_in = _make_inputs(tmp_path, name, n_pairs)
    n_out = _run_both(tmp_path, fa, bam, flags, gbuild.build())
    assert 0 < n_out < n_in


def _run_sharded(tmp_path, fa, bam, flags, engine_lib, n_shards):
    """One input over n_shards processes (`--shard i/N`: a window of the concatenated contigs each, every process counts all
    clustered reads so that the 10 000-read ticks and the threshold rule Q18 fall where they fall in one process), the shards'
    outputs joined by `--merge`: the result must be the reference binary's output."""
    if not pyoracle.reference_available():
        pytest.skip("oracle/_ref/gencore is not built")
    ref_out, merged = str(tmp_path / "ref.bam"), str(tmp_path / "merged.bam")
    r = subprocess.run([pyoracle.REF_BIN, "-i", bam, "-o", ref_out, "-r", fa, "-j", str(tmp_path / "r.json"), "-h", str(tmp_path / "r.html")] + flags,
                       capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    cli = gbuild.build_cli()
    parts, counts = [], []
    for i in range(n_shards):
        part = str(tmp_path / f"shard{i}.bam")
        m = subprocess.run([cli, "-i", bam, "-o", part, "-r", fa, "--engine", engine_lib, "--shard", f"{i}/{n_shards}"] + flags,
                           capture_output=True, text=True, cwd=str(tmp_path))
        assert m.returncode == 0, m.stderr[-2000:]
        parts.append(part)
        counts.append(len(bamfile.bam_records(part)[1]))
    m = subprocess.run([cli, "-o", merged, "--merge"] + parts, capture_output=True, text=True, cwd=str(tmp_path))
    assert m.returncode == 0, m.stderr[-2000:]
    n = bamfile.assert_same_bam(ref_out, merged)
    assert sum(counts) == n
    return n, counts


@pytest.mark.parametrize("name,n_pairs,flags,n_shards", [("cfg2", 6000, [], 2),          # crosses the tick: flushed clusters use -d, the rest 0
                                                         ("cfg3", 7000, [], 3),          # duplex, two contigs over three windows
                                                         ("cfg4", 3000, ["-s", "2"], 2)])
def test_sharded_bam_pipeline_matches_reference_binary_simt(tmp_path, name, n_pairs, flags, n_shards):
    import build as simt_build
    fa, bam, n_in = _make_inputs(tmp_path, name, n_pairs)
    n_out, counts = _run_sharded(tmp_path, fa, bam, flags, simt_build.build(), n_shards)
    assert 0 < n_out < n_in
    assert sum(1 for c in counts if c > 0) >= 2  # (more than one window has reads)


def test_sharded_bam_pipeline_ragged_simt(tmp_path):
    """Pass-through reads (mate unmapped), unmapped and secondary records, clips and indels over four windows."""
    import build as simt_build
    fa, bam, n_in = _make_ragged_inputs(tmp_path, 13, "duplex")
    n_out, counts = _run_sharded(tmp_path, fa, bam, [], simt_build.build(), 4)
    assert 0 < n_out < n_in


def test_shard_flag_is_validated(tmp_path):
    cli = gbuild.build_cli()
    m = subprocess.run([cli, "-i", "x.bam", "-o", "y.bam", "-r", "z.fa", "--shard", "2/2"], capture_output=True, text=True)
    assert m.returncode != 0 and "--shard" in m.stderr


@pytest.mark.gpu
def test_sharded_bam_pipeline_matches_reference_binary_cuda(tmp_path):
    fa, bam, n_in = _make_inputs(tmp_path, "cfg2", 60_000)
    n_out, counts = _run_sharded(tmp_path, fa, bam, [], gbuild.build(), 2)
    assert 0 < n_out < n_in and min(counts) > 0


def test_sharded_bam_script_simt(tmp_path):
    """scripts/sharded_bam.py: the shards as concurrent processes, then the merge."""
    import json
    import build as simt_build
    if not pyoracle.reference_available():
        pytest.skip("oracle/_ref/gencore is not built")
    fa, bam, n_in = _make_inputs(tmp_path, "cfg3", 4000)
    ref_out, out = str(tmp_path / "ref.bam"), str(tmp_path / "sharded.bam")
    r = subprocess.run([pyoracle.REF_BIN, "-i", bam, "-o", ref_out, "-r", fa, "-j", str(tmp_path / "r.json"), "-h", str(tmp_path / "r.html")],
                       capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    script = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "sharded_bam.py")
    m = subprocess.run([sys.executable, script, "-n", "3", "-i", bam, "-o", out, "-r", fa, "--engine", simt_build.build()], capture_output=True, text=True)
    assert m.returncode == 0, m.stderr[-2000:]
    assert json.loads(m.stdout.strip().splitlines()[-1])["shards"] == 3
    assert bamfile.assert_same_bam(ref_out, out) > 0


@pytest.mark.parametrize("name,n_pairs,batch", [("cfg2", 12000, 1500), ("cfg3", 7000, 500)])
def test_bam_pipeline_many_batches_simt(tmp_path, monkeypatch, name, n_pairs, batch):
    """The tool hands batches of its event log from the reading thread to the engine thread to the writing thread: with a small
    batch (GCB_BATCH_PAIRS) an input goes through in many batches, tick flushes and watermarks in between, and must come out
    as from the reference binary."""
    import build as simt_build
    monkeypatch.setenv("GCB_BATCH_PAIRS", str(batch))
    fa, bam, n_in = _make_inputs(tmp_path, name, n_pairs)
    n_out = _run_both(tmp_path, fa, bam, [], simt_build.build())
    assert 0 < n_out < n_in


def test_sharded_bam_pipeline_many_windows_simt(tmp_path):
    """More windows than the input has busy regions: shards that keep nothing write a header-only BAM and the merge still yields
    the reference's output; one shard of one is the unsharded tool."""
    import build as simt_build
    fa, bam, n_in = _make_inputs(tmp_path, "cfg1", 800)
    n_out, counts = _run_sharded(tmp_path, fa, bam, [], simt_build.build(), 16)
    assert 0 < n_out < n_in
    n_one, counts_one = _run_sharded(tmp_path, fa, bam, [], simt_build.build(), 1)
    assert n_one == n_out and counts_one == [n_out]
