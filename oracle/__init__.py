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
Round 2 adds per-entry pins: the reference's own, unmodified Python layer (adcc/workflow.py,
AdcMatrix.py, adc_pp/matrix.py, LazyMp.py, guess/*, solver/davidson.py, adc_pp/transition_dm.py)
was imported from /root/reference and run on the libadcc/ module surface of this repository; the
sigma vectors, diagonals, per-block results, MP energies, states, Davidson iteration counts and
oscillator strengths it produced for ten methods are committed as
tests/golden/reference_adcc_h2o_sto3g.npz (generator beside it) and this oracle agrees with them
to 1e-12 (tests/test_reference_goldens_cpu.py).  libadcc itself (libtensor arithmetic) cannot be
built offline, so there is no oracle/_ref; its floating-point results are pinned only through the
known answers above.

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
