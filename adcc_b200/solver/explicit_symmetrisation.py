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
    def __init__(self, matrix):
        self.symmetrisation_functions = matrix.construct_symmetrisation_for_blocks()

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
                    new = PackedDoubles.from_dense(Tensor._wrap(t.mospaces, t.subspaces, dense))
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
