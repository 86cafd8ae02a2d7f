"""CUDA product path (through the C ABI) against goldens produced by the reference's OWN Python
layer running its unmodified equations and solver (tests/golden/make_reference_adcc_goldens.py):
sigma, diagonal, every block, MP energies, singlet / triplet energies and Davidson iteration
counts for all ten methods on H2O / STO-3G."""
import pytest

from reference_golden_common import METHODS, check_product

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("method", METHODS)
def test_cuda_path_vs_reference_produced_goldens(method):
    check_product(method)


@pytest.mark.parametrize("method", ["adc2", "adc2x", "adc3"])
def test_cuda_packed_engine_vs_reference_produced_goldens(method):
    check_product(method, doubles_storage="packed")


@pytest.mark.parametrize("method,storage", [("adc2", "dense"), ("adc2x", "packed"), ("cvs-adc2", "dense")])
def test_cuda_lanczos_vs_reference_produced_goldens(method, storage):
    """run_adc(..., eigensolver="lanczos") on the device: the reference's energies and iteration counts."""
    from reference_golden_common import check_lanczos
    check_lanczos(method, storage, iter_slack=2)


@pytest.mark.parametrize("method", ["adc2", "adc3"])
def test_cuda_open_shell_vs_reference_produced_goldens(method):
    """Unrestricted open-shell input, kinds "any" and "spin_flip", on the device."""
    from reference_golden_common import check_open_shell
    check_open_shell(method, iter_slack=2)      # 30 - 46 iterations: device rounding may move the last one


@pytest.mark.parametrize("key", ["c1", "c1_adc3", "c2", "c4", "c4_ne", "c3"])
def test_cuda_baseline_shapes_vs_reference_run(key):
    """BASELINE configs[0..3] shapes: sigma, converged states, oscillator strengths and the iteration
    count of the reference's own run_adc on the same synthetic SCF input."""
    from reference_golden_common import check_baseline_shape
    check_baseline_shape(key)


@pytest.mark.parametrize("method", ["adc2", "adc2x", "adc3"])
@pytest.mark.parametrize("no,nv,seed", [(6, 10, 5), (8, 20, 11)])
def test_cuda_synthetic_workload_vs_reference_run(no, nv, seed, method):
    """Device generator + pair-packed store + fused engine of the headline benchmark, at small sizes,
    against the reference's own run on the same integrals."""
    from reference_golden_common import check_synthetic_workload
    check_synthetic_workload(no, nv, seed, method)
