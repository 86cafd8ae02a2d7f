"""The built library really contains the sm_100a instructions the design relies on (checked on the
SASS with cuobjdump, no GPU needed): FP64 tensor-core DMMA, TMA tensor loads, mbarrier transaction
arrives, bulk async copies -- and it is built for sm_100a only."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sass():
    import __graft_entry__ as g
    g.build()
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([tool, "-sass", os.path.join(ROOT, "adcc_b200", "libadccb.so")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    return out.stdout


def test_architecture_is_sm_100a_only(sass):
    archs = {line.split("=")[1].strip() for line in sass.splitlines() if line.startswith("arch =")}
    assert archs == {"sm_100a"}


def _body(sass, needle):
    """SASS text of all functions whose (mangled) name contains ``needle``."""
    parts = sass.split("Function : ")
    return "\n".join(p for p in parts[1:] if needle in p.split("\n", 1)[0])


def test_gemm_uses_dmma_fed_by_tma(sass):
    gemm = _body(sass, "dgemm_tn_tma_kernel")
    assert gemm.count("DMMA.8x8x4") > 500                 # the only native FP64 tensor shape on sm_100a
    assert "UTMALDG.2D" in gemm                           # cp.async.bulk.tensor.2d (TMA) operand loads
    assert "SYNCS.ARRIVE.TRANS64" in gemm and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in gemm    # mbarrier pipeline
    assert "LDS.128" in gemm                              # conflict-free 16-byte fragment loads
    assert "HMMA" not in gemm and "IMMA" not in gemm
    # the peer-destination variant exists beside the plain one
    names = [p.split("\n", 1)[0] for p in sass.split("Function : ")[1:] if "dgemm_tn_tma_kernel" in p]
    assert any("Lb1E" in n for n in names) and any("Lb0E" in n for n in names)


def test_streamed_dot_uses_bulk_copies(sass):
    dot = _body(sass, "dot_bulk_kernel")
    assert "UBLKCP" in dot and "SYNCS.ARRIVE.TRANS64" in dot
    assert "DFMA" in dot


def test_headline_gemm_register_budget():
    """The 112x128 tile (vvvv GEMM of the headline workload) must stay within the 168 registers a
    288-thread CTA can have (3 warps on one SM sub-partition) WITHOUT spilling: accumulators are
    112 of them, a spill in the main loop costs DMMA issue slots."""
    import re
    import __graft_entry__ as g
    g.build()
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([tool, "-res-usage", os.path.join(ROOT, "adcc_b200", "libadccb.so")],
                         capture_output=True, text=True, timeout=600).stdout
    m = re.search(r"dgemm_tn_tma_kernelILi1ELi8ELi14ELi2ELi6ELb0E[^\n]*\n\s*REG:(\d+) STACK:(\d+)", out)
    assert m, "headline GEMM instantiation not found"
    regs, stack = int(m.group(1)), int(m.group(2))
    assert regs <= 168 and stack == 0
