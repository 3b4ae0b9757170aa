"""The product library loads and exports every symbol include/gencore_b200.h declares (no compute: this runs
without a GPU), and refuses to start without an sm_100 device instead of falling back to the CPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from gencore_b200 import build
    return build.build()


def test_header_symbols_are_exported(lib_path):
    from gencore_b200.engine import ABI_SYMBOLS
    hdr = open(os.path.join(ROOT, "include", "gencore_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(gcb_[a-z_]+)\s*\(", hdr)))
    assert declared == sorted(ABI_SYMBOLS)
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), name


def test_default_options_match_reference_defaults(lib_path):
    from gencore_b200.abi import Options
    from gencore_b200.engine import load_library
    lib = load_library(lib_path)
    o = Options()
    lib.gcb_default_options(ctypes.byref(o))
    d = Options.default()
    for name, _ in Options._fields_:
        assert getattr(o, name) == getattr(d, name), name


def test_no_device_means_no_engine(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from gencore_b200.abi import GCB_ERR_NO_DEVICE
    from gencore_b200.engine import ConsensusEngine, EngineError
    with pytest.raises(EngineError) as ei:
        ConsensusEngine()
    assert ei.value.code == GCB_ERR_NO_DEVICE


def test_missing_library_fails_loudly(tmp_path):
    from gencore_b200.engine import ConsensusEngine
    with pytest.raises(FileNotFoundError):
        ConsensusEngine(lib_path=str(tmp_path / "nope.so"))


def test_header_is_self_contained_c(tmp_path):
    """include/gencore_b200.h compiles on its own as C and as C++ (a missing <stddef.h> once broke the oracle's build on a fresh tree)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "only_header.c"
    src.write_text('#include "gencore_b200.h"\nint main(void) { return 0; }\n')
    for cc, lang in (("gcc", "c"), ("g++", "c++")):
        p = subprocess.run([cc, "-x", lang, "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(root, "include"), str(src)], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
