"""The reference's own known-answer vectors for the hot path against the oracle
restatement, and — when oracle/_ref is built — against the reference's compiled functions."""
import pytest

# BamUtil::test, bamutil.cpp:385-423
GETUMI_KAT = [
    ("NB551106:8:H5Y57BGX2:1:13304:3538:1404", "", ""),
    ("NB551106:8:H5Y57BGX2:1:13304:3538:1404:UMI_GAGCATAC", "UMI", "GAGCATAC"),
    ("NB551106:8:H5Y57BGX2:1:13304:3538:1404:UMI_GAGC_ATAC", "UMI", "GAGC_ATAC"),
    ("NB551106:8:H5Y57BGX2:1:13304:3538:1404:GAGC_ATAC", "", "GAGC_ATAC"),
    ("NB551106:8:H5Y57BGX2:1:13304:3538:1404:UMI_X", "UMI", ""),
    ("@V300034954L1C001R0040000002/1:UMI_ATG_AAT", "UMI", "ATG_AAT"),
    ("@V300034954L1C001R0040000002:UMI_ATG_AAT /1", "UMI", "ATG_AAT"),
]
# Cluster::test, cluster.cpp:275-288
UMIDIFF_KAT = [("ATCGATCG", "ATCGATCG", 0), ("ATCGATCG", "ATCGTTC", 2), ("ATCGATCG", "ATCGTTCG", 1), ("AAAA_ATCG", "AAAA_ATCG", 0)]
ISDUPLEX_KAT = [("ATCG_CTAG", "CTAG_ATCG", True), ("AGC_TGA", "TGA_AGC", True), ("AAAA_AAAA", "AAAA_AAAA", True),
                ("CTAG", "CTAG_ATCG", False), ("CTAG", "CCCAGG", False), ("", "", False)]
# extra probes of util.h:59-88 split() corner cases, answered by the compiled reference when available
ISDUPLEX_EXTRA = [("A_", "_A"), ("_A_B", "B_A"), ("A__B", "B__A"), ("A_B_", "B_A"), ("__", "__"), ("A_B", "B_A_"),
                  ("_A", "A_"), ("A_B", "A_B"), ("AC_", "_AC"), ("AC_", "AC_")]
GETUMI_EXTRA = [("x:UMI_ACGT", ""), ("x:ACGT", ""), ("x:_ACGT", ""), ("x:AC_GT_A", ""), ("x:ACNT", ""), ("x:", ""),
                ("nocolon", ""), ("r:UMI_AC_GT", "UMI"), ("r:umi_ACGT", "umi"), ("rUMACGT", "UMI"), ("a:b:ACGT_TTTT", ""),
                ("read/1:UMI_ACGTTGCA:extra", "UMI"), ("IUM_ACGT", "UMI")]


@pytest.mark.parametrize("qname,prefix,umi", GETUMI_KAT)
def test_get_umi_kat(oracle, qname, prefix, umi):
    assert oracle.get_umi(qname, prefix) == umi


@pytest.mark.parametrize("a,b,d", UMIDIFF_KAT)
def test_umi_diff_kat(oracle, a, b, d):
    assert oracle.umi_diff(a, b) == d


@pytest.mark.parametrize("a,b,r", ISDUPLEX_KAT)
def test_is_duplex_kat(oracle, a, b, r):
    assert oracle.is_duplex(a, b) == r


def test_umi_code_is_string_order_and_hamming(oracle):
    """The 4-bit field code (gencore_b200.h) must preserve std::string order and umiDiff."""
    import itertools
    import numpy as np
    from gencore_b200.abi import encode_umi
    umis = ["", "A", "AC", "ACGT", "ACGA", "T", "_", "A_C", "AC_", "_A", "ACGTACGT", "ACGTACGA", "ACGTACG",
            "AAAAAAAA_CCCCCCCC", "CCCCCCCC_AAAAAAAA", "AAAAAAAA_CCCCCCCG", "GGGGGGGGGGGGGGGGG"]
    for a, b in itertools.product(umis, umis):
        ca, cb = encode_umi(a, 2), encode_umi(b, 2)
        assert np.array_equal(ca, oracle.encode_umi(a, 2))
        fa = [(int(ca[k >> 4]) >> (60 - 4 * (k & 15))) & 15 for k in range(32)]
        fb = [(int(cb[k >> 4]) >> (60 - 4 * (k & 15))) & 15 for k in range(32)]
        assert sum(x != y for x, y in zip(fa, fb)) == oracle.umi_diff(a, b), (a, b)
        assert (fa < fb) == (a < b) and (fa == fb) == (a == b), (a, b)
        assert ((int(ca[0]), int(ca[1])) < (int(cb[0]), int(cb[1]))) == (a < b), (a, b)


def test_kat_against_compiled_reference(oracle, have_reference):
    if not have_reference:
        pytest.skip("oracle/_ref not built")
    from oracle.pyoracle import Reference
    ref = Reference(None, "", None, None)
    assert ref.lib.gcr_self_test() == 1  # `gencore test` (unittest.cpp:10-16)
    for q, p, u in GETUMI_KAT:
        assert ref.get_umi(q, p) == u
    for q, p in GETUMI_EXTRA:
        assert oracle.get_umi(q, p) == ref.get_umi(q, p), (q, p)
    for a, b, d in UMIDIFF_KAT:
        assert ref.umi_diff(a, b) == d
    for a, b in [(x, y) for x, y, _ in ISDUPLEX_KAT] + ISDUPLEX_EXTRA:
        assert oracle.is_duplex(a, b) == ref.is_duplex(a, b), (a, b)
        assert oracle.is_duplex(b, a) == ref.is_duplex(b, a), (b, a)
    ref.close()
