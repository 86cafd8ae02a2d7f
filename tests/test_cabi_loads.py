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


def test_header_is_plain_c_and_links(tmp_path):
    """include/adccb.h compiles as C99 (no C++ / torch types in the signatures) and a C program that
    takes the address of every declared entry point links against libadccb.so: the boundary a cgo /
    JNI / ctypes binding would use."""
    import shutil
    import subprocess
    import __graft_entry__ as g
    g.build()
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    names = _declared()
    src = tmp_path / "probe.c"
    body = "\n".join(f"    table[{k}] = (fn_t){name};" for k, name in enumerate(names))
    src.write_text('#include "adccb.h"\n#include <stdio.h>\ntypedef void (*fn_t)(void);\n'
                   f"int main(void) {{\n    fn_t table[{len(names)}];\n{body}\n"
                   f'    printf("%d %d\\n", {len(names)}, table[0] != 0);\n    return 0;\n}}\n')
    libdir = os.path.join(ROOT, "adcc_b200")
    exe = tmp_path / "probe"
    cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
           "-L", libdir, "-ladccb", "-Wl,-rpath," + libdir, "-Wl,--unresolved-symbols=ignore-in-shared-libs",
           "-o", str(exe)]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert os.path.exists(exe)
