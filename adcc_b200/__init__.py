"""adcc_b200: B200-native (sm_100a) engine for adcc's ADC(n) matrix-vector product and
Jacobi-Davidson loop, behind adcc's own Tensor / AmplitudeVector / AdcMatrix / ReferenceState
API.  Usage mirrors adcc:

    import adcc_b200 as adcc
    state = adcc.adc2(scf_dict, n_singlets=3)
"""
from functools import wraps as _wraps

from .AdcMatrix import (AdcBlock, AdcExtraTerm, AdcMatrix, AdcMatrixlike,  # noqa: F401
                        AdcMatrixProjected, AdcMatrixShifted)
from .AdcMethod import AdcMethod  # noqa: F401
from .AmplitudeVector import AmplitudeVector  # noqa: F401
from .AoEriProvider import AoDataHfProvider  # noqa: F401
from .DataHfProvider import DataHfProvider, HartreeFockProvider  # noqa: F401
from .exceptions import InputError  # noqa: F401
from .ExcitedStates import ExcitedStates, State2States  # noqa: F401
from .functions import (contract, copy, direct_sum, dot, einsum, empty_like, evaluate,  # noqa: F401
                        lincomb, linear_combination, nosym_like, ones_like, tensordot, trace,
                        transpose, zeros_like)
from .guess import (guess_symmetries, guess_zero, guesses_any, guesses_singlet,  # noqa: F401
                    guesses_spin_flip, guesses_triplet)
from .Intermediates import Intermediates, register_as_intermediate  # noqa: F401
from .LazyMp import LazyMp  # noqa: F401
from .memory_pool import memory_pool  # noqa: F401
from .MoSpaces import MoSpaces  # noqa: F401
from .packed import shared_host_vector  # noqa: F401
from .projection import Projector, SubspacePartitioning  # noqa: F401
from .ReferenceState import ReferenceState  # noqa: F401
from .solver.davidson import jacobi_davidson  # noqa: F401
from .Symmetry import Symmetry  # noqa: F401
from .Tensor import Tensor  # noqa: F401
from .workflow import run_adc  # noqa: F401

__version__ = "0.1.0"
__backend__ = {"name": "adccb (sm_100a CUDA, dense spin-orbital FP64 store)", "version": __version__,
               "features": ["dmma", "tma"], "blas": "none (hand-written DMMA GEMM)",
               "authors": "adcc_b200"}


def with_runadc_doc(func):
    func.__doc__ = run_adc.__doc__
    return func


def _method_shortcut(name):
    @with_runadc_doc
    def inner(*args, **kwargs):
        return run_adc(*args, **kwargs, method=name)
    inner.__name__ = name.replace("-", "_")
    return inner


@with_runadc_doc
def cis(*args, **kwargs):
    """ADC(1) states with zeroth-order (CIS) properties (adcc/__init__.py:93-96)."""
    from .ExcitedStates import ExcitedStates
    return ExcitedStates(run_adc(*args, **kwargs, method="adc1"), property_method="isr0")


adc0 = _method_shortcut("adc0")
adc1 = _method_shortcut("adc1")
adc2 = _method_shortcut("adc2")
adc2x = _method_shortcut("adc2x")
adc3 = _method_shortcut("adc3")
cvs_adc0 = _method_shortcut("cvs-adc0")
cvs_adc1 = _method_shortcut("cvs-adc1")
cvs_adc2 = _method_shortcut("cvs-adc2")
cvs_adc2x = _method_shortcut("cvs-adc2x")
cvs_adc3 = _method_shortcut("cvs-adc3")


def banner(colour=False):
    """Version / backend summary (the reference prints its own banner, adcc/__init__.py:149-200)."""
    line = "+" + 70 * "-" + "+\n"
    rows = ["adcc_b200: adcc's ADC(n) matvec + Jacobi-Davidson path on NVIDIA B200",
            f"version {__version__}",
            f"backend: {__backend__['name']}"]
    return line + "".join("|{0:^70s}|\n".format(r) for r in rows) + line


def set_n_threads(n):
    """Kept for API compatibility (libadcc_src/pyiface/export_threading.cc:40-58): the device
    path has no host thread pool."""


def get_n_threads():
    return 1
