"""The C-ABI library builds, loads, and exports every symbol include/adccb.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "adccb.h")).read()
    return sorted(set(re.findall(r"\b(adccb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported():
    import __graft_entry__ as g
    g.build()
    from adcc_b200 import _cabi
    assert os.path.exists(_cabi.LIB_PATH)
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in adccb.h but not exported"
    assert sorted(_cabi.EXPORTS) == declared
    lib.adccb_version.restype = ctypes.c_int
    assert lib.adccb_version() >= 100


def test_no_cpu_fallback_in_product():
    # the product package must not import the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "adcc_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
