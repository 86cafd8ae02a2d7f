## ---------------------------------------------------------------------
##
## This file is derived from adcc v0.18.0 (adcc/solver/lanczos.py,
## adcc/solver/LanczosIterator.py).
## Copyright (C) 2020 by the adcc authors
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
## Changes with respect to adcc: iterator and driver condensed into one
## module; the vector algebra is mapped onto batched device kernels of
## libadccb.so (see the module docstring).
##
## ---------------------------------------------------------------------
"""Block Lanczos eigensolver with full re-orthogonalisation and thick restarts: the second caller
of the matvec that ``run_adc(..., eigensolver="lanczos")`` offers (adcc/workflow.py:385-390).

Follows the algorithm and interface of adcc/solver/lanczos.py:36-311 and
adcc/solver/LanczosIterator.py:33-266: block three-term recurrence ``r = A v - q beta^T - v alpha``,
Gram-Schmidt QR of the residual block, full re-orthogonalisation against the Krylov vectors and the
Ritz vectors kept at a restart, ARPACK's convergence criterion on ``|r| |b^T y|``, thick restart
onto ``min_subspace`` Ritz vectors when ``max_subspace`` would be exceeded, true residuals at the
end, same defaults, warnings, state fields and callback identifiers.  Device mapping: the block of
matvecs is one ``AdcMatrix @ [vectors]`` (batched on the engines that batch), every row of alpha /
R / overlaps is one batched-dot launch, every update one fused multi-AXPY, the Ritz vectors, the
reconstructed ``A V`` and the true residuals come from multi-output linear combinations."""
import sys
import warnings

import numpy as np
import scipy.linalg as la

from ..AmplitudeVector import AmplitudeVector
from ..functions import lincomb
from ..timings import Timer
from .common import select_eigenpairs
from .explicit_symmetrisation import IndexSymmetrisation
from .orthogonaliser import GramSchmidtOrthogonaliser
from .SolverStateBase import EigenSolverStateBase


class LanczosIterator:
    """Iterator over growing block-Krylov subspaces (LanczosIterator.py:33-169); a thick restart
    passes the kept Ritz vectors ``Y``, their values ``Theta`` and ``Sigma = Y^T A guesses``."""

    def __init__(self, matrix, guesses, ritz_vectors=None, ritz_values=None, ritz_overlaps=None,
                 explicit_symmetrisation=None):
        if not isinstance(guesses, list):
            guesses = [guesses]
        for guess in guesses:
            if not isinstance(guess, AmplitudeVector):
                raise TypeError("One of the guesses is not an AmplitudeVector")
        n_block = len(guesses)
        if ritz_values is None:
            ritz_vectors, ritz_values, ritz_overlaps = [], np.empty((0,)), np.empty((0, n_block))
        elif len(ritz_vectors) != len(ritz_values) or ritz_overlaps.shape != (len(ritz_values), n_block):
            raise ValueError("Restart vector shape does not agree with problem.")
        self.matrix = matrix
        self.ritz_vectors, self.ritz_values, self.ritz_overlaps = ritz_vectors, ritz_values, ritz_overlaps
        self.n_problem = matrix.shape[1]
        self.n_block = n_block
        self.n_restart = len(ritz_values)
        self.ortho = GramSchmidtOrthogonaliser(explicit_symmetrisation)
        self.explicit_symmetrisation = explicit_symmetrisation
        self.timer = Timer()
        self.lanczos_subspace = []
        self.alphas, self.betas = [], []       # diagonal / side-diagonal blocks of the subspace matrix
        self.residual = guesses
        self.n_iter = 0
        self.n_applies = 0

    def __iter__(self):
        return self

    def _alpha(self, v, r):
        return np.array([v[p] @ r for p in range(self.n_block)]).reshape(self.n_block, self.n_block)

    def __next__(self):
        nb = self.n_block
        if self.n_iter == 0:
            v = self.ortho.orthogonalise(self.residual)
            self.lanczos_subspace = v
            r = self.matrix @ v
            alpha = self._alpha(v, r)
            # r = r - v alpha - Y Sigma, then full re-orthogonalisation against Y
            Sigma, Y = self.ritz_overlaps, self.ritz_vectors
            r = [lincomb(np.hstack(([1], -alpha[:, p], -Sigma[:, p])), [r[p]] + v + Y, evaluate=True)
                 for p in range(nb)]
            r = [self.ortho.orthogonalise_against(r[p], Y) for p in range(nb)]
            self.residual = r
            self.n_iter, self.n_applies = 1, nb
            self.alphas, self.betas = [alpha], []
            return LanczosSubspace(self)

        q = self.lanczos_subspace[-nb:]
        v, beta = self.ortho.qr(self.residual)
        if np.linalg.norm(beta) < np.finfo(float).eps * self.n_problem:
            raise StopIteration()                 # the new vectors would be decoupled from the old ones
        self.n_applies += nb
        r = self.matrix @ v
        r = [lincomb(np.hstack(([1], -(beta.T)[:, p])), [r[p]] + q, evaluate=True) for p in range(nb)]
        alpha = self._alpha(v, r)
        r = [lincomb(np.hstack(([1], -alpha[:, p])), [r[p]] + v, evaluate=True) for p in range(nb)]
        everything = self.lanczos_subspace + self.ritz_vectors
        r = [self.ortho.orthogonalise_against(r[p], everything) for p in range(nb)]
        self.n_iter += 1
        self.lanczos_subspace.extend(v)
        self.residual = r
        self.alphas.append(alpha)
        self.betas.append(beta)
        return LanczosSubspace(self)


class LanczosSubspace:
    """Snapshot of the Krylov subspace after one step (LanczosIterator.py:172-266)."""

    def __init__(self, iterator):
        self._it = iterator
        self.n_iter = iterator.n_iter
        self.n_restart = iterator.n_restart
        self.residual = iterator.residual
        self.n_block = iterator.n_block
        self.n_problem = iterator.n_problem
        self.matrix = iterator.matrix
        self.ortho = iterator.ortho
        self.alphas, self.betas = iterator.alphas, iterator.betas
        self.n_applies = iterator.n_applies
        self.subspace = iterator.ritz_vectors + iterator.lanczos_subspace

    @property
    def subspace_matrix(self):
        """Projection of the matrix into the subspace: diag(Theta) bordered by Sigma for the kept
        Ritz vectors, then the block-tridiagonal (alpha, beta) part."""
        nb, nr = self.n_block, self.n_restart
        n_ss = len(self._it.lanczos_subspace) + nr
        T = np.zeros((n_ss, n_ss))
        if nr > 0:
            T[:nr, :nr] = np.diag(self._it.ritz_values)
            T[:nr, nr:nr + nb] = self._it.ritz_overlaps
            T[nr:nr + nb, :nr] = self._it.ritz_overlaps.T
        for i, alpha in enumerate(self.alphas):
            rng = slice(nr + i * nb, nr + (i + 1) * nb)
            T[rng, rng] = alpha
        for i, beta in enumerate(self.betas):
            rng = slice(nr + i * nb, nr + (i + 1) * nb)
            nxt = slice(rng.start + nb, rng.stop + nb)
            T[rng, nxt] = beta.T
            T[nxt, rng] = beta
        return T

    @property
    def rayleigh_extension(self):
        n_ss = len(self.subspace)
        b = np.zeros((n_ss, self.n_block))
        for i in range(self.n_block):
            b[n_ss - self.n_block + i, i] = 1
        return b

    @property
    def matrix_product(self):
        """A V reconstructed from the Lanczos relation A V = V T + r b^T."""
        r, T = self.residual, self.subspace_matrix
        n_ss = len(self.subspace)
        AV = []
        for i in range(n_ss):
            coefficients = [T[j, i] for j in range(n_ss) if T[j, i] != 0]
            vectors = [v for j, v in enumerate(self.subspace) if T[j, i] != 0]
            if i >= n_ss - self.n_block:
                coefficients.append(1)
                vectors.append(r[i - (n_ss - self.n_block)])
            AV.append(lincomb(np.array(coefficients), vectors, evaluate=True))
        return AV

    def check_orthogonality(self, tolerance=None):
        if tolerance is None:
            tolerance = self.n_problem * np.finfo(float).eps
        orth = np.array([v @ self.subspace for v in self.subspace]).reshape(len(self.subspace), -1)
        orth = np.max(np.abs(orth - np.eye(len(self.subspace))))
        if orth > tolerance:
            warnings.warn(la.LinAlgWarning("LanczosSubspace has lost orthogonality. Expect inaccurate results."))
        return orth


class LanczosState(EigenSolverStateBase):
    def __init__(self, iterator):
        super().__init__(iterator.matrix)
        self.n_restart = 0
        self.subspace_residual = None
        self.subspace_vectors = None
        self.algorithm = "lanczos"
        self.timer = iterator.timer


def default_print(state, identifier, file=sys.stdout):
    """Progress lines of the callback (lanczos.py:48-73)."""
    from ..timings import strtime, strtime_short
    if identifier == "start" and state.n_iter == 0:
        print("Niter n_ss  max_residual  time  Ritz values", file=file)
    elif identifier == "next_iter":
        time_iter = state.timer.current("iteration")
        print(f"{state.n_iter:3d}  {len(state.subspace_vectors):4d}  {np.max(state.residual_norms):12.5g}  "
              f"{strtime_short(time_iter):5s}", "", state.eigenvalues[:7], file=file)
        if hasattr(state, "subspace_orthogonality"):
            print(33 * " " + f"nonorth: {state.subspace_orthogonality:5.3g}", file=file)
    elif identifier == "is_converged":
        print("=== Converged ===", file=file)
        print("    Number of matrix applies:   ", state.n_applies, file=file)
        print("    Total solver time:          ", strtime(state.timer.total("iteration")), file=file)
    elif identifier == "restart":
        print("=== Restart ===", file=file)


def compute_true_residuals(subspace, rvals, rvecs, epair_mask):
    """Eigenvectors, residuals A x - theta x and their norms for the selected Ritz pairs, from the
    reconstructed A V (lanczos.py:76-96)."""
    V, AV = subspace.subspace, subspace.matrix_product
    cols = [rvecs[:, i] for i in range(rvecs.shape[1]) if i in epair_mask]
    vals = [rvals[i] for i in range(rvecs.shape[1]) if i in epair_mask]
    residuals = [lincomb(np.hstack((c, -t * c)), AV + V, evaluate=True) for c, t in zip(cols, vals)]
    eigenvectors = [lincomb(c, V, evaluate=True) for c in cols]
    rnorms = np.array([np.sqrt(r @ r) for r in residuals])
    return eigenvectors, residuals, rnorms


def amend_true_residuals(state, subspace, rvals, rvecs, epair_mask):
    state.eigenvectors, state.residuals, state.residual_norms = \
        compute_true_residuals(subspace, rvals, rvecs, epair_mask)
    return state


def check_convergence(subspace, rvals, rvecs, tol):
    """ARPACK's criterion |r| |b^T y_i| <= max(eps max|theta|, tol |theta_i|) (lanczos.py:110-131)."""
    b = subspace.rayleigh_extension
    norm_residual = np.sqrt(sum(subspace.residual[p] @ subspace.residual[p] for p in range(subspace.n_block)))
    mintol = np.finfo(float).eps * np.max(np.abs(rvals))
    error = []
    converged = np.ones_like(rvals, dtype=bool)
    for i, rval in enumerate(rvals):
        lhs = norm_residual * np.linalg.norm(b.T @ rvecs[:, i])
        if lhs > max(mintol, tol * abs(rval)):
            converged[i] = False
        error.append(lhs / abs(rval) if mintol < tol * abs(rval) else lhs * tol / mintol)
    return converged, np.array(error)


def lanczos_iterations(iterator, n_ep, min_subspace, max_subspace, conv_tol=1e-9, which="LA", max_iter=100,
                       callback=None, debug_checks=False, state=None):
    """Drive the Lanczos iterations (lanczos.py:134-252)."""
    if callback is None:
        def callback(state, identifier):
            pass
    if state is None:
        state = LanczosState(iterator)
        callback(state, "start")
        state.timer.restart("iteration")
        n_applies_offset = 0
    else:
        n_applies_offset = state.n_applies

    subspace = rvals = rvecs = epair_mask = None
    for subspace in iterator:
        b = subspace.rayleigh_extension
        with state.timer.record("rayleigh_ritz"):
            rvals, rvecs = np.linalg.eigh(subspace.subspace_matrix)
        if debug_checks:
            orthotol = max(conv_tol / 1000, subspace.n_problem * np.finfo(float).eps)
            state.subspace_orthogonality = subspace.check_orthogonality(orthotol)
        is_rval_converged, eigenpair_error = check_convergence(subspace, rvals, rvecs, conv_tol)

        state.n_iter += 1
        state.n_applies = subspace.n_applies + n_applies_offset
        state.converged = False
        state.eigenvectors = None                       # only formed at the end
        state.subspace_vectors = subspace.subspace
        state.subspace_residual = subspace.residual
        epair_mask = select_eigenpairs(rvals, n_ep, which)
        state.eigenvalues = rvals[epair_mask]
        state.residual_norms = eigenpair_error[epair_mask]
        converged = np.all(is_rval_converged[epair_mask])
        callback(state, "next_iter")
        state.timer.restart("iteration")

        if converged:
            state = amend_true_residuals(state, subspace, rvals, rvecs, epair_mask)
            state.converged = True
            callback(state, "is_converged")
            state.timer.stop("iteration")
            return state
        if state.n_iter >= max_iter:
            warnings.warn(la.LinAlgWarning(
                f"Maximum number of iterations (== {max_iter}) reached in lanczos procedure."))
            state = amend_true_residuals(state, subspace, rvals, rvecs, epair_mask)
            state.timer.stop("iteration")
            state.converged = False
            return state
        if len(rvecs) + subspace.n_block > max_subspace:
            # thick restart: keep min_subspace Ritz vectors, continue from the orthonormalised residual
            callback(state, "restart")
            keep = select_eigenpairs(rvals, min_subspace, which)
            V = subspace.subspace
            vn, betan = subspace.ortho.qr(subspace.residual)
            Y = [lincomb(rvecs[:, i], V, evaluate=True) for i in range(rvecs.shape[1]) if i in keep]
            Theta = rvals[keep]
            Sigma = rvecs[:, keep].T @ b @ betan.T
            iterator = LanczosIterator(iterator.matrix, vn, ritz_vectors=Y, ritz_values=Theta,
                                       ritz_overlaps=Sigma,
                                       explicit_symmetrisation=iterator.explicit_symmetrisation)
            state.n_restart += 1
            return lanczos_iterations(iterator, n_ep, min_subspace, max_subspace, conv_tol, which, max_iter,
                                      callback, debug_checks, state)

    state = amend_true_residuals(state, subspace, rvals, rvecs, epair_mask)
    state.timer.stop("iteration")
    state.converged = False
    warnings.warn(la.LinAlgWarning(
        "Lanczos procedure found maximal subspace possible. Iteration cannot be "
        "continued like this and will be aborted without convergence. Try a different guess."))
    return state


def lanczos(matrix, guesses, n_ep, max_subspace=None, conv_tol=1e-9, which="LM", max_iter=100, callback=None,
            debug_checks=False, explicit_symmetrisation=IndexSymmetrisation, min_subspace=None):
    """Lanczos eigensolver for ADC problems (lanczos.py:255-311); ``guesses`` fix the block size."""
    if explicit_symmetrisation is not None and isinstance(explicit_symmetrisation, type):
        explicit_symmetrisation = explicit_symmetrisation(matrix)
    iterator = LanczosIterator(matrix, guesses, explicit_symmetrisation=explicit_symmetrisation)
    if not isinstance(guesses, list):
        guesses = [guesses]
    if not max_subspace:
        max_subspace = max(2 * n_ep + len(guesses), 20, 8 * len(guesses))
    if not min_subspace:
        min_subspace = n_ep + 2 * len(guesses)
    if conv_tol < matrix.shape[1] * np.finfo(float).eps:
        warnings.warn(la.LinAlgWarning(
            "Convergence tolerance (== {:5.2g}) lower than estimated maximal numerical accuracy (== {:5.2g}). "
            "Convergence might be hard to achieve.".format(conv_tol, matrix.shape[1] * np.finfo(float).eps)))
    return lanczos_iterations(iterator, n_ep, min_subspace, max_subspace, conv_tol, which, max_iter, callback,
                              debug_checks)
