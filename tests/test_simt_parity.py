"""The kernel source itself (gencore_b200/csrc/*.cu*), compiled by g++ against the SIMT interpreter in
tests/simt_check/, executed on the CPU and compared bit-for-bit with the oracle.  This is a check of the
kernels' LOGIC on a box without a GPU; the same comparisons run on the real library in test_gpu_parity.py."""
import os
import sys

import pytest

import parity_cases
from parity import assert_results_equal

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt_check"))

CASES = parity_cases.small_cases()


@pytest.fixture(scope="module")
def simt_lib():
    import build as simt_build
    return simt_build.build()


@pytest.mark.parametrize("name,thunk", CASES, ids=[c[0] for c in CASES])
def test_kernels_match_oracle_under_simt_check(simt_lib, oracle, name, thunk):
    from gencore_b200.engine import ConsensusEngine
    batch, genome, opt = thunk()
    with ConsensusEngine(opt, 0, lib_path=simt_lib) as eng:
        eng.set_reference(genome)
        res = eng.cluster_by_umi(batch)
    assert_results_equal(batch, res, oracle.consensus(batch, genome, opt), name)
