"""CPU oracle: NumPy restatement of adcc's ADC(n) matvec + Jacobi-Davidson hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``adcc_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` use it, and there only as the checker / the CPU arm.

Parity status: PINNED.  The restatement reproduces the reference's in-source known
answers for H2O/STO-3G (adcc/tests/smoke_test.py:32-58: adc0..adc3 and cvs-adc0..3,
3 singlets + 3 triplets, atol 1e-6) from the reference's own fixture
(examples/water/data.py -> tests/golden/h2o_sto3g.npz), see tests/test_oracle_golden.py, and
again from its second, PySCF-generated fixture (examples/0_basics/h2o_sto3g_hfdata.hdf5 ->
tests/golden/h2o_sto3g_pyscf_hfdata.npz, full 14^4 eri_ffff), see tests/test_hdf5_cpu.py; the
(anti)symmetrisers are checked against the reference's vendored tensor goldens.
Sigma vectors are not dumped anywhere in the reference tree, so per-entry sigma parity
between the reference and this oracle is pinned only through those eigenvalues and
through hermiticity / symmetry invariants.

Everything works on dense spin-orbital arrays in adcc's index convention
(libadcc_src/MoSpaces.cc:379-405: alpha block then beta block per axis, each sorted by
subspace then orbital energy).
"""
from .hfdata import OracleHf, load_h2o_sto3g, restricted_dict_from_spatial  # noqa: F401
from .refstate import OracleRefState  # noqa: F401
from .mp import OracleMp  # noqa: F401
from .adc import OracleAdcMatrix  # noqa: F401
from .davidson import jacobi_davidson, davidson_eigsh  # noqa: F401
from .run import run_adc  # noqa: F401
