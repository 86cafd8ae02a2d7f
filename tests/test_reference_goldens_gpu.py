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
