import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device and the compiled library: without them they are skipped, not failed (the product itself
    refuses to run there: gcb_create returns GCB_ERR_NO_DEVICE)."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    have = have and os.path.exists(os.path.join(ROOT, "gencore_b200", "csrc", "libgencore_b200.so"))
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and gencore_b200/csrc/libgencore_b200.so")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def have_reference():
    from oracle import pyoracle
    if not pyoracle.reference_available() and os.path.exists("/root/reference/src/cluster.cpp"):
        pyoracle.build(port=False, ref=True)
    return pyoracle.reference_available()
