"""The drop-in proof: the reference itself, with its two Cluster::clusterByUMI call sites (gencore.cpp:355, gencore.cpp:409)
bound to libgencore_b200.so through integration/gcbbridge.h (oracle/_ref/gencore_bridged, built from the reference's own sources
by oracle/Makefile `bridge`), against the UNMODIFIED reference binary on the same BAM + FASTA: identical output BAM and an
identical JSON report (every Stats counter of preStats / postStats, the supporting-reads histogram, the coverage arrays).
On the CPU box the engine behind the ABI is the SIMT-check build of the kernel source; the `gpu` tests load the CUDA library."""
import os
import subprocess
import sys

import pytest

import bamfile
from gencore_b200 import build as gbuild
from oracle import pyoracle
from test_bam_pipeline import CONFIGS, _make_inputs, _make_ragged_inputs

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt_check"))

BRIDGED = os.path.join(os.path.dirname(pyoracle.REF_BIN), "gencore_bridged")


def _bridged_available():
    if not os.path.exists(BRIDGED) and os.path.exists("/root/reference/src/cluster.cpp"):
        subprocess.run(["make", "-C", os.path.dirname(os.path.abspath(pyoracle.__file__)), "bridge"], capture_output=True)
    return os.path.exists(BRIDGED) and pyoracle.reference_available()


def _report_lines(path):
    """The JSON report as text, without the line that names the run itself (the reference's JSON is not always valid JSON — an
    empty array prints as `[,]`-like text in some sections — so it is compared as written)."""
    with open(path) as f:
        return [ln for ln in f.read().splitlines() if '"command"' not in ln]


def _run_stock_and_bridged(tmp_path, fa, bam, flags, engine_lib):
    if not _bridged_available():
        pytest.skip("oracle/_ref/gencore_bridged is not built")
    outs = {}
    for tag, binary in (("stock", pyoracle.REF_BIN), ("bridged", BRIDGED)):
        d = tmp_path / tag
        d.mkdir()
        env = dict(os.environ, GENCORE_B200_ENGINE=engine_lib)
        r = subprocess.run([binary, "-i", bam, "-o", str(d / "out.bam"), "-r", fa, "-j", str(d / "out.json"), "-h", str(d / "out.html")] + flags,
                           capture_output=True, text=True, cwd=str(d), env=env)
        assert r.returncode == 0, (tag, r.stderr[-2000:])
        outs[tag] = d
    n = bamfile.assert_same_bam(str(outs["stock"] / "out.bam"), str(outs["bridged"] / "out.bam"))
    a, b = _report_lines(outs["stock"] / "out.json"), _report_lines(outs["bridged"] / "out.json")
    assert a == b, "the JSON reports differ: " + "; ".join(f"{x} | {y}" for x, y in zip(a, b) if x != y)[:500]
    assert any("total_fragments" in ln for ln in a) and any("supporting_reads" in ln or "duplication" in ln for ln in a), "the report holds the Stats of the run"
    return n


@pytest.mark.parametrize("name,n_pairs,flags", CONFIGS, ids=[f"{c[0]}_{c[1]}_{'_'.join(c[2]) or 'default'}" for c in CONFIGS])
def test_bridged_reference_equals_stock_reference_simt(tmp_path, name, n_pairs, flags):
    import build as simt_build
    fa, bam, n_in = _make_inputs(tmp_path, name, n_pairs)
    n_out = _run_stock_and_bridged(tmp_path, fa, bam, flags, simt_build.build())
    assert 0 < n_out < n_in


@pytest.mark.parametrize("seed,umi,flags", [(11, "none", []), (13, "duplex", []), (14, "duplex", ["-s", "2", "-c", "8"]),
                                            (15, "single", ["-u", "UMI", "--high_qual", "35", "--low_qual", "10"])])
def test_bridged_reference_equals_stock_reference_ragged_simt(tmp_path, seed, umi, flags):
    import build as simt_build
    fa, bam, n_in = _make_ragged_inputs(tmp_path, seed, umi)
    n_out = _run_stock_and_bridged(tmp_path, fa, bam, flags, simt_build.build())
    assert 0 < n_out < n_in


@pytest.mark.gpu
@pytest.mark.parametrize("name,n_pairs,flags", [("cfg2", 60_000, []), ("cfg3", 40_000, []), ("cfg4", 30_000, ["-s", "2"])])
def test_bridged_reference_equals_stock_reference_cuda(tmp_path, name, n_pairs, flags):
    fa, bam, n_in = _make_inputs(tmp_path, name, n_pairs)
    n_out = _run_stock_and_bridged(tmp_path, fa, bam, flags, gbuild.build())
    assert 0 < n_out < n_in
