## ---------------------------------------------------------------------
##
## This file is derived from adcc v0.18.0 (adcc/workflow.py).
## Copyright (C) 2018 by the adcc authors
##
## adcc is free software: you can redistribute it and/or modify
## it under the terms of the GNU General Public License as published
## by the Free Software Foundation, either version 3 of the License, or
## (at your option) any later version.
##
## adcc is distributed in the hope that it will be useful,
## but WITHOUT ANY WARRANTY; without even the implied warranty of
## MERCHANTABILITY or FITNESS FOR A PARTICULAR PURPOSE.  See the
## GNU General Public License for more details.
##
## You should have received a copy of the GNU General Public License
## along with adcc. If not, see <http://www.gnu.org/licenses/>.
##
## Changes with respect to adcc: condensed; the vector algebra is mapped onto
## batched device kernels of libadccb.so (see the module docstring).
##
## ---------------------------------------------------------------------
"""User-facing driver: run_adc and the adcN / cvs_adcN short-hands.

Same call signature, validation, defaults and guess heuristics as adcc/workflow.py:45-520
(``construct_adcmatrix``, ``validate_state_parameters``, ``diagonalise_adcmatrix``,
``estimate_n_guesses``, ``obtain_guesses_by_inspection``); the Jacobi-Davidson eigensolver of the
hot path (eigensolver="davidson", default) and the block Lanczos solver (eigensolver="lanczos").
"""
import contextlib
import sys
import warnings

from .AdcMatrix import AdcMatrix, AdcMatrixlike
from .AdcMethod import AdcMethod
from .exceptions import InputError
from .ExcitedStates import ExcitedStates
from .guess import guesses_any, guesses_singlet, guesses_spin_flip, guesses_triplet
from .LazyMp import LazyMp
from .ReferenceState import ReferenceState
from .solver.davidson import default_print as davidson_default_print
from .solver.davidson import jacobi_davidson
from .solver.lanczos import default_print as lanczos_default_print
from .solver.lanczos import lanczos
from .solver.explicit_symmetrisation import IndexSpinSymmetrisation, IndexSymmetrisation

__all__ = ["run_adc"]


def run_adc(data_or_matrix, n_states=None, kind="any", conv_tol=None, eigensolver=None,
            guesses=None, n_guesses=None, n_guesses_doubles=None, output=sys.stdout,
            core_orbitals=None, frozen_core=None, frozen_virtual=None, method=None,
            n_singlets=None, n_triplets=None, n_spin_flip=None, environment=None, **solverargs):
    """Run an ADC calculation (see adcc.run_adc for the meaning of every argument)."""
    matrix = construct_adcmatrix(data_or_matrix, core_orbitals=core_orbitals,
                                 frozen_core=frozen_core, frozen_virtual=frozen_virtual,
                                 method=method)
    n_states, kind = validate_state_parameters(
        matrix.reference_state, n_states=n_states, n_singlets=n_singlets,
        n_triplets=n_triplets, n_spin_flip=n_spin_flip, kind=kind)
    spin_change = -1 if (kind == "spin_flip" and guesses is None) else None
    if eigensolver is None:
        eigensolver = "davidson"
    if environment not in (None, False, {}):
        raise NotImplementedError("Environment coupling (PE/PCM) is outside the B200 hot path.")
    diagres = diagonalise_adcmatrix(
        matrix, n_states, kind, guesses=guesses, n_guesses=n_guesses,
        n_guesses_doubles=n_guesses_doubles, conv_tol=conv_tol, output=output,
        eigensolver=eigensolver, **solverargs)
    exstates = ExcitedStates(diagres)
    exstates.kind = kind
    exstates.spin_change = spin_change
    return exstates


def construct_adcmatrix(data_or_matrix, core_orbitals=None, frozen_core=None,
                        frozen_virtual=None, method=None):
    if not isinstance(data_or_matrix, AdcMatrixlike) and method is None:
        raise InputError("method needs to be explicitly provided unless data_or_matrix is an "
                         "AdcMatrixlike.")
    if method is not None and not isinstance(method, AdcMethod):
        try:
            method = AdcMethod(method)
        except ValueError as e:
            raise InputError(str(e))
    if not isinstance(data_or_matrix, (ReferenceState, AdcMatrixlike, LazyMp)):
        if method.is_core_valence_separated and core_orbitals is None:
            raise InputError("If core-valence separation approximation is applied then the number "
                             "of core orbitals needs to be specified via the parameter core_orbitals.")
        try:
            data_or_matrix = ReferenceState(data_or_matrix, core_orbitals=core_orbitals,
                                            frozen_core=frozen_core, frozen_virtual=frozen_virtual)
        except ValueError as e:
            raise InputError(str(e))
    else:
        mospaces = data_or_matrix.mospaces
        for name, val, sp in (("core_orbitals", core_orbitals, "o2"), ("frozen_core", frozen_core, "o3"),
                              ("frozen_virtual", frozen_virtual, "v2")):
            if val is not None:
                n = mospaces.n_orbs_alpha(sp) if mospaces.has_subspace(sp) else 0
                warnings.warn(f"Ignored {name} parameter because data_or_matrix is a ReferenceState, "
                              f"a LazyMp or an AdcMatrixlike object  (which has a value of {name}={n}).")
    if isinstance(data_or_matrix, (ReferenceState, LazyMp)):
        try:
            return AdcMatrix(method, data_or_matrix)
        except ValueError as e:
            raise InputError(str(e))
    if method is not None and method != data_or_matrix.method:
        warnings.warn("Ignored method parameter because data_or_matrix is an AdcMatrixlike, which "
                      "implicitly sets the method")
    return data_or_matrix


def validate_state_parameters(reference_state, n_states=None, n_singlets=None, n_triplets=None,
                              n_spin_flip=None, kind="any"):
    if sum(nst is not None for nst in [n_states, n_singlets, n_triplets, n_spin_flip]) > 1:
        raise InputError("One may only specify one out of n_states, n_singlets, n_triplets and n_spin_flip")
    if n_singlets is not None:
        if not reference_state.restricted:
            raise InputError("The n_singlets parameter may only be employed for restricted references")
        if kind not in ["singlet", "any"]:
            raise InputError(f"Kind parameter {kind} not compatible with n_singlets > 0")
        kind, n_states = "singlet", n_singlets
    if n_triplets is not None:
        if not reference_state.restricted:
            raise InputError("The n_triplets parameter may only be employed for restricted references")
        if kind not in ["triplet", "any"]:
            raise InputError(f"Kind parameter {kind} not compatible with n_triplets > 0")
        kind, n_states = "triplet", n_triplets
    if n_spin_flip is not None:
        if reference_state.restricted:
            raise InputError("The n_spin_flip parameter may only be employed for unrestricted references")
        if kind not in ["spin_flip", "any"]:
            raise InputError(f"Kind parameter {kind} not compatible with n_spin_flip > 0")
        kind, n_states = "spin_flip", n_spin_flip
    if n_states is None or n_states == 0:
        raise InputError("No excited states to be computed. Specify at least one of n_states, "
                         "n_singlets, n_triplets, or n_spin_flip")
    if n_states < 0:
        raise InputError("n_states needs to be positive")
    if kind not in ["any", "spin_flip", "singlet", "triplet"]:
        raise InputError("The kind parameter may only take the values 'any', 'singlet', 'triplet' "
                         "or 'spin_flip'")
    if kind in ["singlet", "triplet"] and not reference_state.restricted:
        raise InputError("kind==singlet and kind==triplet are only valid for ADC calculations in "
                         "combination with a restricted ground state.")
    if kind in ["spin_flip"] and reference_state.restricted:
        raise InputError("kind==spin_flip is only valid for ADC calculations in combination with "
                         "an unrestricted ground state.")
    return n_states, kind


def setup_solver_printing(solmethod_name, matrix, kind, default_print, output=None):
    kstr = " " if kind == "any" else " " + kind
    method_name = f"{matrix}"
    if hasattr(matrix, "method"):
        method_name = matrix.method.name
    if output is not None:
        print(f"Starting {method_name}{kstr} {solmethod_name} ...", file=output)

        def inner_callback(state, identifier):
            default_print(state, identifier, output)
        return inner_callback
    return None


def diagonalise_adcmatrix(matrix, n_states, kind, eigensolver="davidson", guesses=None,
                          n_guesses=None, n_guesses_doubles=None, conv_tol=None, output=sys.stdout,
                          **solverargs):
    reference_state = matrix.reference_state
    if conv_tol is None:
        conv_tol = max(10 * reference_state.conv_tol, 1e-6)
    if reference_state.conv_tol > conv_tol:
        raise InputError(f"Convergence tolerance of SCF results (== {reference_state.conv_tol}) needs "
                         f"to be lower than ADC convergence tolerance parameter conv_tol (== {conv_tol}).")
    explicit_symmetrisation = IndexSymmetrisation
    if kind in ["singlet", "triplet"]:
        explicit_symmetrisation = IndexSpinSymmetrisation(matrix, enforce_spin_kind=kind)
    if eigensolver == "davidson":
        n_guesses_per_state = 2
        callback = setup_solver_printing("Jacobi-Davidson", matrix, kind,
                                         davidson_default_print, output=output)
        run_eigensolver = jacobi_davidson
    elif eigensolver == "lanczos":
        n_guesses_per_state = 1
        callback = setup_solver_printing("Lanczos", matrix, kind, lanczos_default_print, output=output)
        run_eigensolver = lanczos
    else:
        raise InputError(f"Solver {eigensolver} unknown, try 'davidson'.")

    structure = contextlib.nullcontext()
    if guesses is None:
        if kind in ("singlet", "triplet", "any") and hasattr(matrix, "spin_conserving_vectors"):
            structure = matrix.spin_conserving_vectors()      # the guesses made below conserve M_s
        if n_guesses is None:
            n_guesses = estimate_n_guesses(matrix=matrix, n_states=n_states,
                                           singles_only=("pphh" not in matrix.axis_blocks),
                                           n_guesses_per_state=n_guesses_per_state)
        guesses = obtain_guesses_by_inspection(matrix, n_guesses, kind, n_guesses_doubles)
    else:
        if len(guesses) < n_states:
            raise InputError(f"Less guesses provided via guesses (== {len(guesses)}) than states to "
                             f"be computed (== {n_states})")
        if n_guesses is not None:
            warnings.warn("Ignoring n_guesses parameter, since guesses are explicitly provided.")
        if n_guesses_doubles is not None:
            warnings.warn("Ignoring n_guesses_doubles parameter, since guesses are explicitly provided.")
    solverargs.setdefault("which", "SA")
    with structure:
        return run_eigensolver(matrix, guesses, n_ep=n_states, conv_tol=conv_tol, callback=callback,
                               explicit_symmetrisation=explicit_symmetrisation, **solverargs)


def estimate_n_guesses(matrix, n_states, singles_only=True, n_guesses_per_state=2):
    n_guesses = n_guesses_per_state * max(2, n_states)
    if singles_only:
        mospaces = matrix.mospaces
        sp_occ = "o2" if matrix.is_core_valence_separated else "o1"
        n_guesses = min(n_guesses, mospaces.n_orbs_alpha(sp_occ) * mospaces.n_orbs_alpha("v1"))
    return max(n_states, n_guesses)


def obtain_guesses_by_inspection(matrix, n_guesses, kind, n_guesses_doubles=None):
    if n_guesses_doubles is not None and n_guesses_doubles > 0 and "pphh" not in matrix.axis_blocks:
        raise InputError("n_guesses_doubles > 0 is only sensible if the ADC method has a doubles "
                         "block (i.e. it is *not* ADC(0), ADC(1) or a variant thereof.")
    guess_function = {"any": guesses_any, "singlet": guesses_singlet, "triplet": guesses_triplet,
                      "spin_flip": guesses_spin_flip}[kind]
    n_guess_singles = n_guesses
    if n_guesses_doubles is not None:
        n_guess_singles = n_guesses - n_guesses_doubles
    singles_guesses = guess_function(matrix, n_guess_singles, block="ph")
    doubles_guesses = []
    if "pphh" in matrix.axis_blocks:
        if n_guesses_doubles is None:
            n_guesses_doubles = n_guesses - len(singles_guesses)
        if n_guesses_doubles > 0:
            doubles_guesses = guess_function(matrix, n_guesses_doubles, block="pphh")
    total_guesses = singles_guesses + doubles_guesses
    if len(total_guesses) < n_guesses:
        raise InputError(f"Less guesses found than requested: {len(total_guesses)} found, "
                         f"{n_guesses} requested")
    return total_guesses
