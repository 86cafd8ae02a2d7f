"""Oracle driver mirroring adcc.run_adc (adcc/workflow.py:45-218,351-422) for the hot path.

Test infrastructure, see oracle/__init__.py.
"""
import numpy as np
from .hfdata import OracleHf
from .refstate import OracleRefState
from .adc import OracleAdcMatrix
from .davidson import jacobi_davidson, Symmetrisation
from .guess import obtain_guesses, estimate_n_guesses


def run_adc(data, method, n_states=None, kind="any", n_singlets=None, n_triplets=None,
            conv_tol=None, core_orbitals=None, frozen_core=None, frozen_virtual=None,
            n_guesses=None, n_guesses_doubles=None, max_iter=70, max_subspace=None,
            callback=None):
    if isinstance(data, OracleAdcMatrix):
        matrix = data
    else:
        hf = data if isinstance(data, OracleHf) else OracleHf(data)
        ref = OracleRefState(hf, core_orbitals=core_orbitals, frozen_core=frozen_core,
                             frozen_virtual=frozen_virtual)
        matrix = OracleAdcMatrix(method, ref)
    if n_singlets is not None:
        kind, n_states = "singlet", n_singlets
    if n_triplets is not None:
        kind, n_states = "triplet", n_triplets
    if conv_tol is None:
        conv_tol = max(10 * matrix.hf.conv_tol, 1e-6)     # workflow.py:362-363
    if n_guesses is None:
        n_guesses = estimate_n_guesses(matrix, n_states)
    guesses = obtain_guesses(matrix, n_guesses, kind, n_guesses_doubles)
    symm = Symmetrisation(matrix, kind)
    state = jacobi_davidson(matrix, guesses, n_ep=n_states, conv_tol=conv_tol, which="SA",
                            explicit_symmetrisation=symm, max_iter=max_iter,
                            max_subspace=max_subspace, callback=callback)
    state.kind = kind
    state.excitation_energy = np.array(state.eigenvalues)
    return state
