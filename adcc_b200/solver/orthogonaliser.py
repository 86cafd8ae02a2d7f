## ---------------------------------------------------------------------
##
## This file is derived from adcc v0.18.0 (adcc/solver/orthogonaliser.py).
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
## Changes with respect to adcc: condensed; every projection is one batched
## dot launch plus one fused multi-AXPY launch of libadccb.so.
##
## ---------------------------------------------------------------------
"""Gram-Schmidt orthogonaliser of the Lanczos solver (adcc/solver/orthogonaliser.py:28-96)."""
import numpy as np

from ..functions import lincomb


class GramSchmidtOrthogonaliser:
    def __init__(self, explicit_symmetrisation=None, n_rounds=1):
        self.explicit_symmetrisation = explicit_symmetrisation
        self.n_rounds = n_rounds

    def qr(self, vectors):
        """(Q, R) with orthonormal Q spanning ``vectors`` and upper-triangular R = Q^T vectors."""
        if len(vectors) == 0:
            return []
        if len(vectors) == 1:
            norm_v = np.sqrt(vectors[0] @ vectors[0])
            return [vectors[0] / norm_v], np.array([[norm_v]])
        n_vec = len(vectors)
        Q = self.orthogonalise(vectors)
        R = np.zeros((n_vec, n_vec))
        for i in range(n_vec):
            R[i, i:] = Q[i] @ vectors[i:]              # one batched dot launch per row
        return Q, R

    def orthogonalise(self, vectors):
        """Orthonormal vectors spanning the same space, built one by one."""
        if len(vectors) == 0:
            return []
        subspace = [vectors[0] / np.sqrt(vectors[0] @ vectors[0])]
        for v in vectors[1:]:
            w = self.orthogonalise_against(v, subspace)
            subspace.append(w / np.sqrt(w @ w))
        return subspace

    def orthogonalise_against(self, vector, subspace):
        """(1 - SS SS^T) vector for an orthonormal ``subspace``, followed by the explicit
        symmetrisation (the projection may lose index / spin symmetry to rounding)."""
        if len(subspace) == 0 and self.explicit_symmetrisation is None:
            return vector
        for _ in range(self.n_rounds):
            if len(subspace) > 0:
                coefficients = np.hstack(([1], -(vector @ subspace)))
                vector = lincomb(coefficients, [vector] + subspace, evaluate=True)
            else:                       # nothing to project out: a new container, the caller's stays as it is
                vector = type(vector)(**{b: vector[b] for b in vector.blocks})
            if self.explicit_symmetrisation is not None:
                vector = self.explicit_symmetrisation.symmetrise(vector)
        return vector
