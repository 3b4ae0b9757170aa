"""gcb_stat_depth (stat_depth_kernel, Stats::statDepth stats.cpp:56-83) pinned against the UNMODIFIED reference binary: the
"coverage" arrays of the before_processing section of its JSON report (Stats::addRead of every input record, gencore.cpp:222)
must equal the kernel's bins over the same records (the report prints round(bin / coverage_sampling), stats.cpp:176-178; with
--coverage_sampling 1 that is the bin itself).  (The after_processing section cannot serve: the reference writes its report
before the destructor flushes the last records, gencore.cpp:21-22, so it covers an unspecified prefix of the output.)"""
import os
import re
import struct
import subprocess
import sys

import numpy as np
import pytest

import bamfile
from oracle import pyoracle
from test_bam_pipeline import _make_inputs, _make_ragged_inputs

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt_check"))


def _records(path):
    """(tid, pos, l_qseq) of every record of a BAM, and the target lengths of its header."""
    raw = bamfile.read_bgzf(path)
    l_text = struct.unpack_from("<i", raw, 4)[0]
    i = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, i)[0]
    i += 4
    tlen = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", raw, i)[0]
        tlen.append(struct.unpack_from("<i", raw, i + 4 + l_name)[0])
        i += 4 + l_name + 4
    tid, pos, lq = [], [], []
    while i < len(raw):
        block = struct.unpack_from("<i", raw, i)[0]
        t, p = struct.unpack_from("<ii", raw, i + 4)
        l = struct.unpack_from("<i", raw, i + 4 + 16)[0]
        tid.append(t); pos.append(p); lq.append(l)
        i += 4 + block
    return np.array(tid, np.int32), np.array(pos, np.int32), np.array(lq, np.int32), np.array(tlen, np.int64)


def _report_coverage(json_path, section):
    """The coverage arrays of one section of the reference's JSON report, concatenated in contig order."""
    text = open(json_path).read()
    part = text[text.index(f'"{section}"'):]
    cov = part[part.index('"coverage":{'):]
    cov = cov[:cov.index("}")]
    out = []
    for m in re.finditer(r'"[^"]+":\[([^\]]*)\]', cov):
        out += [int(x) for x in m.group(1).split(",") if x.strip()]
    return np.array(out, np.int64)


def _c_round_div(a, step):
    """round((double)a / step) as stats.cpp:178 computes it (half away from zero)."""
    return np.floor(a.astype(np.float64) / step + 0.5).astype(np.int64)


def _check(tmp_path, fa, bam, step, engine_factory):
    if not pyoracle.reference_available():
        pytest.skip("oracle/_ref/gencore is not built")
    out, js = str(tmp_path / "ref.bam"), str(tmp_path / "r.json")
    r = subprocess.run([pyoracle.REF_BIN, "-i", bam, "-o", out, "-r", fa, "-j", js, "-h", str(tmp_path / "r.html"), "--coverage_sampling", str(step)],
                       capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-1500:]
    with engine_factory() as eng:
        for path, section in ((bam, "before_processing"),):
            tid, pos, lq, tlen = _records(path)
            half = len(tid) // 2  # two calls: the bins are additive
            d = eng.stat_depth(tid[:half], pos[:half], lq[:half], tlen, step)
            d = eng.stat_depth(tid[half:], pos[half:], lq[half:], tlen, step, depth=d)
            want = _report_coverage(js, section)
            assert len(want) == len(d) and d.sum() > 0
            np.testing.assert_array_equal(_c_round_div(d, step), want, err_msg=f"{section} step {step}")


def _simt_engine():
    import build as simt_build
    from gencore_b200.abi import Options
    from gencore_b200.engine import ConsensusEngine
    return ConsensusEngine(Options.default(), 0, lib_path=simt_build.build())


@pytest.mark.parametrize("step", [1, 7, 150, 10000])
def test_stat_depth_kernel_matches_the_reference_report_simt(tmp_path, step):
    fa, bam, _ = _make_inputs(tmp_path, "cfg2", 2500)
    _check(tmp_path, fa, bam, step, _simt_engine)


@pytest.mark.parametrize("step", [1, 64])
def test_stat_depth_kernel_ragged_with_unmapped_records_simt(tmp_path, step):
    fa, bam, _ = _make_ragged_inputs(tmp_path, 13, "duplex")  # mixed lengths, unmapped and secondary records, reads near contig ends
    _check(tmp_path, fa, bam, step, _simt_engine)


@pytest.mark.gpu
@pytest.mark.parametrize("step", [1, 1000])
def test_stat_depth_kernel_matches_the_reference_report_cuda(tmp_path, step):
    from gencore_b200.abi import Options
    from gencore_b200.engine import ConsensusEngine
    fa, bam, _ = _make_inputs(tmp_path, "cfg3", 30000)
    _check(tmp_path, fa, bam, step, lambda: ConsensusEngine(Options.default(), 0))
