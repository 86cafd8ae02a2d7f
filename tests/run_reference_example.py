"""Execute one of the reference's offline example scripts (examples/water/sto3g_*.py), unmodified, on
this repository's backend: reference `adcc` package on the `libadcc/` surface, CPU stand-in kernels
(build container only).    python tests/run_reference_example.py sto3g_adc3.py"""
import os
import runpy
import sys

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
EX = os.path.join(REF, "examples", "water")
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools", "refshim"), REF, EX]
import cpu_shim  # noqa: E402


class _Patch:
    def setattr(self, obj, name, val, raising=True):
        setattr(obj, name, val)


cpu_shim.install(_Patch())
os.chdir(EX)                     # the scripts read data.py / write nothing
runpy.run_path(os.path.join(EX, sys.argv[1]), run_name="__main__")
