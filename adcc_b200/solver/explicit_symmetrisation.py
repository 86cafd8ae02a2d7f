"""Explicit (index and spin) symmetrisation of new subspace vectors.

adcc/solver/explicit_symmetrisation.py:38-98 plus the native singlet purification
libadcc_src/amplitude_vector_enforce_spin_kind.cc:33-170.  libtensor keeps the spin-block maps
of restricted singlet/triplet vectors structurally (mapped blocks are never stored); on the dense
device store the same subspace is enforced by the projector kernel ``spin_project`` before the
index symmetrisation (see oracle/davidson.py: enforce_spin_maps for the CPU statement)."""
from .. import backend as be
from ..AmplitudeVector import AmplitudeVector
from ..Tensor import Tensor


def amplitude_vector_enforce_spin_kind(tensor, block, spin_kind):
    """In-place, libadcc.amplitude_vector_enforce_spin_kind(tensor, "d", kind)."""
    if block == "s":
        return
    if block != "d":
        raise NotImplementedError("Not implemented for block != 'd'")
    if spin_kind == "triplet":
        return
    if spin_kind != "singlet":
        raise NotImplementedError("Only implemented for spin_kind == 'singlet' and spin_kind == 'triplet'.")
    n_alpha = [tensor.mospaces.n_orbs_alpha(s) for s in tensor.subspaces]
    be.spin_singlet_(tensor._contig(), n_alpha)


class IndexSymmetrisation:
    #: spin projection of the fused correction step: 0 none, +1 singlet, -1 triplet
    _spin_fac = 0.0

    def __init__(self, matrix):
        self.symmetrisation_functions = matrix.construct_symmetrisation_for_blocks()
        # the fused kernel implements AdcMatrix's own recipes (AdcMatrix.py:431-472) only
        from ..AdcMatrix import AdcMatrix
        standard = getattr(type(matrix), "construct_symmetrisation_for_blocks", None) \
            is AdcMatrix.construct_symmetrisation_for_blocks
        self._sym_mode = None
        if standard and type(self) in (IndexSymmetrisation, IndexSpinSymmetrisation):
            self._sym_mode = {"pphh": 2 if matrix.is_core_valence_separated else 1}

    def precondition_and_symmetrise(self, preconditioner, vectors):
        """[Purify(Sym(SpinProject(v_k / (D - shift_k))))]: the division as one coalesced launch, then
        ONE gather kernel (adccb_precond_apply without a diagonal) for the projector, the three index
        (anti)symmetrisations and the purification together - two launches per block instead of
        six.  (Measured: dividing inside the gather kernel, 16 FP64 divisions per doubles element,
        made the single-launch form slower than the chain it replaced.)  Blocks that need neither
        projection nor symmetrisation take the division only.  Returns None when the fused kernel
        does not apply (packed / distributed stores, custom symmetrisation functions, non-Jacobi
        preconditioners); the caller then takes the separate steps."""
        from .preconditioner import JacobiPreconditioner
        if self._sym_mode is None or type(preconditioner) is not JacobiPreconditioner:
            return None
        shifts = preconditioner.shifts
        if not hasattr(shifts, "__len__") or len(shifts) != len(vectors):
            return None
        for vec in vectors:
            for b, t in vec.items():
                if type(t) is not Tensor or t.ndim not in (2, 4) or b not in preconditioner.diagonal.blocks:
                    return None
        out = []
        for vec, shift in zip(vectors, shifts):
            blocks = {}
            for b, t in vec.items():
                n_alpha = [t.mospaces.n_orbs_alpha(s) for s in t.subspaces]
                data = be.div_shift(t._contig(), preconditioner.diagonal[b]._contig(), float(shift))
                sym_mode = self._sym_mode.get(b, 0)
                if sym_mode != 0 or self._spin_fac != 0.0:
                    data = be.precond_apply(data, None, 0.0, n_alpha, sym_mode=sym_mode, spin_fac=self._spin_fac,
                                            purify=(self._spin_fac > 0 and b == "pphh"))
                blocks[b] = Tensor._wrap(t.mospaces, t.subspaces, data, t.symmetry)
            out.append(AmplitudeVector(**blocks))
        return out

    def symmetrise(self, new_vectors):
        if isinstance(new_vectors, AmplitudeVector):
            return self.symmetrise([new_vectors])[0]
        for vec in new_vectors:
            if not isinstance(vec, AmplitudeVector):
                raise TypeError("new_vectors has to be an iterable of AmplitudeVector")
            for b in vec.blocks:
                if b not in self.symmetrisation_functions:
                    continue
                sym = vec[b].symmetry
                vec[b] = self.symmetrisation_functions[b](vec[b]).evaluate()
                vec[b].symmetry = sym
        return new_vectors


class IndexSpinSymmetrisation(IndexSymmetrisation):
    def __init__(self, matrix, enforce_spin_kind="singlet"):
        super().__init__(matrix)
        if enforce_spin_kind not in ("singlet", "triplet"):
            raise ValueError("enforce_spin_kind needs to be 'singlet' or 'triplet'")
        self.enforce_spin_kind = enforce_spin_kind
        self._spin_fac = 1.0 if enforce_spin_kind == "singlet" else -1.0

    def symmetrise(self, new_vectors):
        if isinstance(new_vectors, AmplitudeVector):
            return self.symmetrise([new_vectors])[0]
        from ..packed import PackedDoubles
        fac = 1.0 if self.enforce_spin_kind == "singlet" else -1.0
        packed_done = []
        for vec in new_vectors:
            for b in vec.blocks:
                t = vec[b]
                n_alpha = [t.mospaces.n_orbs_alpha(s) for s in t.subspaces]
                if isinstance(t, PackedDoubles):
                    # packed doubles: project and purify on the unpacked tensor, pack again (index
                    # antisymmetry is structural in this store, both steps preserve it)
                    dense = be.spin_project(t.to_dense()._contig(), n_alpha, fac)
                    if self.enforce_spin_kind == "singlet":
                        be.spin_singlet_(dense, n_alpha)
                    store = PackedDoubles if t._storage in ("packed", "slab") else type(t)   # CVS: virtual pairs
                    new = store.from_dense(Tensor._wrap(t.mospaces, t.subspaces, dense))
                    if getattr(t, "slab", None) is not None:      # distributed vector: keep this rank's slab
                        from ..packed import SlabDoubles
                        new = SlabDoubles.from_full(new, t.slab)
                    new.symmetry = t.symmetry
                    vec[b] = new
                    packed_done.append(id(vec))
                    continue
                vec[b] = Tensor._wrap(t.mospaces, t.subspaces,
                                      be.spin_project(t._contig(), n_alpha, fac), t.symmetry)
        new_vectors = super().symmetrise(new_vectors)
        for vec in new_vectors:
            if "pphh" in vec.blocks and id(vec) not in packed_done:
                amplitude_vector_enforce_spin_kind(vec.pphh, "d", self.enforce_spin_kind)
        return new_vectors
