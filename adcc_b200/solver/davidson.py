## ---------------------------------------------------------------------
##
## This file is derived from adcc v0.18.0 (adcc/solver/davidson.py).
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
"""Block Jacobi-Davidson eigensolver for the ADC matrix.

Follows the algorithm and interface of adcc/solver/davidson.py:78-477: Rayleigh-Ritz in the
subspace, residuals as ONE linear combination over Ax + SS, Jacobi preconditioning with shifts
ritz - 1e-6, explicit symmetrisation, conventional Gram-Schmidt with conditional
re-orthogonalisation (trigger max|<p,SS>|/|p| > N*eps), collapse when the subspace would exceed
max_subspace, same defaults (max_subspace = max(6 n_ep, 20, 5 n_guesses), max_iter = 70,
which = "SA"), same warnings, same state fields and timer tasks.  Device mapping: every
``p @ SS`` is one batched-dot launch, every linear combination one fused multi-AXPY launch,
the only host round trips per iteration are the small subspace matrix and the norms.
The dense subspace problem is solved with ``scipy.linalg.eigh`` and the ``which`` n_block
pairs are kept (the reference uses ARPACK when n_ss > n_block, davidson.py:185-193; the
selected pairs are identical, the result is deterministic).
"""
import sys
import warnings

import numpy as np
import scipy.linalg as la

from ..AdcMatrix import AdcMatrixlike
from ..AmplitudeVector import AmplitudeVector, dot_rows
from ..functions import evaluate, lincomb, lincomb_many
from ..timings import strtime, strtime_short
from .common import select_eigenpairs
from .explicit_symmetrisation import IndexSymmetrisation
from .preconditioner import JacobiPreconditioner
from .SolverStateBase import EigenSolverStateBase


class DavidsonState(EigenSolverStateBase):
    def __init__(self, matrix, guesses):
        super().__init__(matrix)
        self.residuals = None
        self.subspace_vectors = guesses.copy()
        self.algorithm = "davidson"


def default_print(state, identifier, file=sys.stdout):
    if identifier == "start" and state.n_iter == 0:
        print("Niter n_ss  max_residual  time  Ritz values", file=file)
    elif identifier == "next_iter":
        time_iter = state.timer.current("iteration")
        fmt = "{n_iter:3d}  {ss_size:4d}  {residual:12.5g}  {tstr:5s}"
        print(fmt.format(n_iter=state.n_iter, tstr=strtime_short(time_iter),
                         ss_size=len(state.subspace_vectors),
                         residual=np.max(state.residual_norms)),
              "", state.eigenvalues[:7], file=file)
    elif identifier == "is_converged":
        print("=== Converged ===", file=file)
        print("    Number of matrix applies:   ", state.n_applies, file=file)
        print("    Total solver time:          ", strtime(state.timer.total("iteration")), file=file)
    elif identifier == "restart":
        print("=== Restart ===", file=file)


def _project(Ass, SS, Ax, rows, upper):
    """Fill rows (and, by symmetry, columns) of the subspace matrix: one batched dot per row."""
    rows = list(rows)
    allvals = dot_rows([(SS[i], Ax[i:] if upper else Ax[:i + 1]) for i in rows])
    for i, vals in zip(rows, allvals):
        if upper:
            Ass[i, i:] = vals
            Ass[i:, i] = vals
        else:
            Ass[i, :i + 1] = vals
            Ass[:i + 1, i] = vals


def davidson_iterations(matrix, state, max_subspace, max_iter, n_ep, n_block, is_converged, which,
                        callback=None, preconditioner=None, preconditioning_method="Davidson",
                        debug_checks=False, residual_min_norm=None, explicit_symmetrisation=None,
                        max_subspace_iter=None):
    if preconditioning_method not in ["Davidson", "Sleijpen-van-der-Vorst"]:
        raise ValueError("Only 'Davidson' and 'Sleijpen-van-der-Vorst' are valid preconditioner methods")
    if preconditioning_method == "Sleijpen-van-der-Vorst":
        raise NotImplementedError("Sleijpen-van-der-Vorst preconditioning not yet implemented.")
    if callback is None:
        def callback(state, identifier):
            pass

    n_problem = matrix.shape[1]
    n_ss_vec = len(state.subspace_vectors)
    assert n_block >= n_ep and n_block <= n_ss_vec
    SS = state.subspace_vectors
    Ass_cont = np.empty((max_subspace, max_subspace))
    eps = np.finfo(float).eps
    if residual_min_norm is None:
        residual_min_norm = 2 * n_problem * eps

    callback(state, "start")
    state.timer.restart("iteration")
    with state.timer.record("projection"):
        Ax = evaluate(matrix @ SS)
        state.n_applies += n_ss_vec
    Ass = Ass_cont[:n_ss_vec, :n_ss_vec]
    with state.timer.record("projection"):
        _project(Ass, SS, Ax, range(n_ss_vec), upper=True)

    def ritz_vectors(rvecs, mask):
        rows = [v for i, v in enumerate(np.transpose(rvecs)) if i in mask]
        return lincomb_many(rows, SS) if rows else []

    while state.n_iter < max_iter:
        state.n_iter += 1
        assert n_block <= len(SS) <= max_subspace

        with state.timer.record("rayleigh_ritz"):
            evals, evecs = la.eigh(Ass)
            if Ass.shape == (n_block, n_block):
                rvals, rvecs = evals, evecs
            else:
                sel = select_eigenpairs(evals, n_block, which)
                rvals, rvecs = evals[sel], evecs[:, sel]

        with state.timer.record("residuals"):
            # device mapping: all n_block combinations from ONE pass over Ax + SS
            residuals = lincomb_many([np.hstack((v, -rvals[i] * v))
                                      for i, v in enumerate(np.transpose(rvecs))], Ax + SS)
            assert len(residuals) == n_block
            epair_mask = select_eigenpairs(rvals, n_ep, which)
            state.eigenvalues = rvals[epair_mask]
            state.residuals = [residuals[i] for i in epair_mask]
            state.residual_norms = np.sqrt(np.array(
                [float(d[0]) for d in dot_rows([(r, [r]) for r in state.residuals])]))

        callback(state, "next_iter")
        state.timer.restart("iteration")
        if is_converged(state):
            state.eigenvectors = ritz_vectors(rvecs, epair_mask)
            state.converged = True
            callback(state, "is_converged")
            state.timer.stop("iteration")
            return state

        if state.n_iter == max_iter:
            warnings.warn(la.LinAlgWarning(
                f"Maximum number of iterations (== {max_iter}) reached in davidson procedure."))
            state.eigenvectors = ritz_vectors(rvecs, epair_mask)
            state.timer.stop("iteration")
            state.converged = False
            return state

        if n_ss_vec + n_block > max_subspace:
            callback(state, "restart")
            with state.timer.record("projection"):
                SS = lincomb_many(np.transpose(rvecs), SS)
                state.subspace_vectors = SS
                Ax = lincomb_many(np.transpose(rvecs), Ax)
                n_ss_vec = len(SS)
                Ass = Ass_cont[:n_ss_vec, :n_ss_vec]
                _project(Ass, SS, Ax, range(n_ss_vec), upper=True)

        with state.timer.record("preconditioner"):
            preconds = None
            if preconditioner and hasattr(preconditioner, "update_shifts"):
                rvals_eps = 1e-6
                preconditioner.update_shifts(rvals - rvals_eps)
            fused = getattr(explicit_symmetrisation, "precondition_and_symmetrise", None)
            if preconditioner and fused is not None:
                # device mapping: preconditioner + explicit symmetrisation in one kernel per block
                preconds = fused(preconditioner, residuals)
            if preconds is None:
                preconds = evaluate(preconditioner @ residuals) if preconditioner else residuals
                if explicit_symmetrisation:
                    explicit_symmetrisation.symmetrise(preconds)

        with state.timer.record("orthogonalisation"):
            n_ss_added = 0
            for i in range(n_block):
                pvec = preconds[i]
                coefficients = np.hstack(([1], -np.asarray(pvec @ SS)))
                pvec = lincomb(coefficients, [pvec] + SS, evaluate=True)
                overlaps = np.array(pvec @ ([pvec] + SS))      # norm and re-check in one pass
                pnorm = np.sqrt(overlaps[0])
                if pnorm < residual_min_norm:
                    continue
                with state.timer.record("reorthogonalisation"):
                    ss_overlap = overlaps[1:]
                    max_ortho_loss = np.max(np.abs(ss_overlap)) / pnorm
                    if max_ortho_loss > n_problem * eps:
                        coefficients = np.hstack(([1], -ss_overlap))
                        pvec = lincomb(coefficients, [pvec] + SS, evaluate=True)
                        pnorm = np.sqrt(pvec @ pvec)
                        state.reortho_triggers.append(max_ortho_loss)
                if pnorm >= residual_min_norm:
                    SS.append(evaluate(pvec / pnorm))
                    n_ss_added += 1
                    n_ss_vec = len(SS)

            if debug_checks:
                orth = np.array([np.asarray(SS[j] @ SS) for j in range(n_ss_vec)])
                orth -= np.eye(n_ss_vec)
                state.subspace_orthogonality = np.max(np.abs(orth))
                if state.subspace_orthogonality > n_problem * eps:
                    warnings.warn(la.LinAlgWarning(
                        "Subspace in Davidson has lost orthogonality. Max. deviation from "
                        "orthogonality is {:.4E}. Expect inaccurate results.".format(
                            state.subspace_orthogonality)))

        if n_ss_added == 0:
            state.timer.stop("iteration")
            state.converged = False
            state.eigenvectors = ritz_vectors(rvecs, epair_mask)
            warnings.warn(la.LinAlgWarning(
                "Davidson procedure could not generate any further vectors for the subspace. "
                "Iteration cannot be continued like this and will be aborted without "
                "convergence. Try a different guess."))
            return state

        with state.timer.record("projection"):
            Ax.extend(matrix @ SS[-n_ss_added:])
            state.n_applies += n_ss_added
        Ass = Ass_cont[:n_ss_vec, :n_ss_vec]
        with state.timer.record("projection"):
            _project(Ass, SS, Ax, range(n_ss_vec - n_ss_added, n_ss_vec), upper=False)
    return state


def eigsh(matrix, guesses, n_ep=None, n_block=None, max_subspace=None, conv_tol=1e-9, which="SA",
          max_iter=70, callback=None, preconditioner=None, preconditioning_method="Davidson",
          debug_checks=False, residual_min_norm=None, explicit_symmetrisation=IndexSymmetrisation,
          max_subspace_iter=None):
    """Davidson eigensolver for ADC problems (adcc/solver/davidson.py:353-468)."""
    if not isinstance(matrix, AdcMatrixlike):
        raise TypeError("matrix is not of type AdcMatrixlike")
    for guess in guesses:
        if not isinstance(guess, AmplitudeVector):
            raise TypeError("One of the guesses is not of type AmplitudeVector")
    if preconditioner is not None and isinstance(preconditioner, type):
        preconditioner = preconditioner(matrix)
    if explicit_symmetrisation is not None and isinstance(explicit_symmetrisation, type):
        explicit_symmetrisation = explicit_symmetrisation(matrix)

    if n_ep is None:
        n_ep = len(guesses)
    elif n_ep > len(guesses):
        raise ValueError(f"n_ep (= {n_ep}) cannot exceed the number of guess vectors (= {len(guesses)}).")
    if n_block is None:
        n_block = n_ep
    elif n_block < n_ep:
        raise ValueError(f"n_block (= {n_block}) cannot be smaller than the number of states "
                         f"requested (= {n_ep}).")
    elif n_block > len(guesses):
        raise ValueError(f"n_block (= {n_block}) cannot exceed the number of guess vectors "
                         f"(= {len(guesses)}).")
    if not max_subspace:
        max_subspace = max(6 * n_ep, 20, 5 * len(guesses))
    elif max_subspace < 2 * n_block:
        raise ValueError(f"max_subspace (= {max_subspace}) needs to be at least twice as large as "
                         f"n_block (n_block = {n_block}).")
    elif max_subspace < len(guesses):
        raise ValueError(f"max_subspace (= {max_subspace}) cannot be smaller than the number of "
                         f"guess vectors (= {len(guesses)}).")

    def convergence_test(state):
        state.residuals_converged = state.residual_norms < conv_tol
        state.converged = bool(np.all(state.residuals_converged))
        return state.converged

    if conv_tol < matrix.shape[1] * np.finfo(float).eps:
        warnings.warn(la.LinAlgWarning(
            "Convergence tolerance (== {:5.2g}) lower than estimated maximal numerical accuracy "
            "(== {:5.2g}). Convergence might be hard to achieve.".format(
                conv_tol, matrix.shape[1] * np.finfo(float).eps)))

    state = DavidsonState(matrix, guesses)
    davidson_iterations(matrix, state, max_subspace, max_iter, n_ep=n_ep, n_block=n_block,
                        is_converged=convergence_test, callback=callback, which=which,
                        preconditioner=preconditioner,
                        preconditioning_method=preconditioning_method, debug_checks=debug_checks,
                        residual_min_norm=residual_min_norm,
                        explicit_symmetrisation=explicit_symmetrisation,
                        max_subspace_iter=max_subspace_iter)
    return state


def jacobi_davidson(*args, **kwargs):
    return eigsh(*args, preconditioner=JacobiPreconditioner, preconditioning_method="Davidson", **kwargs)


def davidson(*args, **kwargs):
    return eigsh(*args, preconditioner=None, **kwargs)
