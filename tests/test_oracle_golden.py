"""The oracle restatement against committed outputs of the reference itself (tests/golden/*.npz,
made by tests/make_golden.py).  Needs neither /root/reference nor oracle/_ref."""
import glob
import os

import pytest

from parity import assert_matches_reference, load_golden

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def test_golden_present():
    assert len(GOLDEN) >= 8


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_golden(oracle, path):
    batch, genome, opt, ref_pairs, ref_out, ref_stats, n_ref = load_golden(path)
    res = oracle.consensus(batch, genome, opt)
    assert_matches_reference(batch, res, ref_pairs, ref_out, ref_stats, n_ref)
