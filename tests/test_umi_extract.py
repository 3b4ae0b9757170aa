"""gcb_extract_umi (BamUtil::getUMI, bamutil.cpp:40-112, as a kernel) against the reference's own known-answer
vectors (bamutil.cpp:385-423), the extra probes answered by the compiled reference, and the oracle on random names.
The CPU run drives the kernel source through the SIMT-check build; the `gpu` run drives the CUDA library."""
import os
import sys

import numpy as np
import pytest

from test_oracle_kat import GETUMI_EXTRA, GETUMI_KAT

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt_check"))


def _names(rng, n):
    alphabet = list("ACGT_:NUMI/ x1")
    out = []
    for _ in range(n):
        l = int(rng.integers(0, 40))
        s = "".join(rng.choice(alphabet, l))
        if rng.random() < 0.5:
            s += ":" + rng.choice(["UMI_", "umi_", "", "_", "UMI"]) + "".join(rng.choice(list("ACGT"), int(rng.integers(0, 20))))
            if rng.random() < 0.4:
                s += "_" + "".join(rng.choice(list("ACGT"), int(rng.integers(0, 12))))
        out.append(s)
    return out


def _check(eng, oracle):
    from gencore_b200.abi import encode_umi
    # the reference's vectors
    for prefix in ("", "UMI"):
        kat = [(q, u) for q, p, u in GETUMI_KAT if p == prefix]
        out, st = eng.extract_umi([q.encode() for q, _ in kat], prefix, 2)
        assert (st == 0).all()
        for row, (q, u) in zip(out, kat):
            assert np.array_equal(row, encode_umi(u, 2)), (q, prefix, u)
    # probes + random names, every prefix, against the oracle
    rng = np.random.Generator(np.random.PCG64(7))
    names = [q for q, _ in GETUMI_EXTRA] + _names(rng, 3000) + [""]
    for prefix in ("", "UMI", "umi", "U", ":"):
        for words in (1, 2, 4):
            out, st = eng.extract_umi([q.encode() for q in names], prefix, words)
            for row, s, q in zip(out, st, names):
                u = oracle.get_umi(q, prefix)
                if len(u) > 16 * words:
                    assert s == 1 and np.array_equal(row, encode_umi(u[:16 * words], words)), (q, prefix, u)
                else:
                    assert s == 0 and np.array_equal(row, encode_umi(u, words)), (q, prefix, u, row)
    out, st = eng.extract_umi([], "UMI", 1)
    assert out.shape == (0, 1)


def test_extract_umi_simt(oracle):
    import build as simt_build
    from gencore_b200.engine import ConsensusEngine
    with ConsensusEngine(None, 0, lib_path=simt_build.build()) as eng:
        _check(eng, oracle)


@pytest.mark.gpu
def test_extract_umi_cuda(oracle):
    from gencore_b200.engine import ConsensusEngine
    with ConsensusEngine(None, 0) as eng:
        n0 = eng.launches
        _check(eng, oracle)
        assert eng.launches > n0
