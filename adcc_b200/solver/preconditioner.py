## ---------------------------------------------------------------------
##
## This file is derived from adcc v0.18.0 (adcc/solver/preconditioner.py).
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
"""Jacobi preconditioner (D - sigma)^-1 (adcc/solver/preconditioner.py:29-87).

The reference evaluates ``v / (D - sigma)`` through a temporary filled tensor
(libadcc_src/pyiface/export_Tensor.cc:380-391); here it is one fused kernel per block."""
import numpy as np

from ..AdcMatrix import AdcMatrixlike
from ..AmplitudeVector import AmplitudeVector


class PreconditionerIdentity:
    def apply(self, invecs):
        return invecs

    def __matmul__(self, x):
        return x


class JacobiPreconditioner:
    def __init__(self, adcmatrix, shifts=0.0):
        if not isinstance(adcmatrix, AdcMatrixlike):
            raise TypeError("Only an AdcMatrixlike may be used with this preconditioner for now.")
        self.diagonal = adcmatrix.diagonal()
        self.shifts = shifts

    def update_shifts(self, shifts):
        self.shifts = shifts

    def _apply_one(self, vec, shift):
        return AmplitudeVector(**{b: t.div_shifted(self.diagonal[b], shift) for b, t in vec.items()})

    def apply(self, invecs):
        if isinstance(invecs, AmplitudeVector):
            if not isinstance(self.shifts, (float, np.number)):
                raise TypeError("Can only apply JacobiPreconditioner to a single vector if shifts "
                                "is only a single number.")
            return self._apply_one(invecs, float(self.shifts))
        if isinstance(invecs, list):
            if len(self.shifts) != len(invecs):
                raise ValueError("Number of vectors passed does not agree with number of shifts "
                                 "stored inside precoditioner. Update using the 'update_shifts' method.")
            return [self._apply_one(v, float(self.shifts[i])) for i, v in enumerate(invecs)]
        raise TypeError("Input type not understood: " + str(type(invecs)))

    def __matmul__(self, invecs):
        return self.apply(invecs)
