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
