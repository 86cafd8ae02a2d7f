"""Subspace partitioning and projectors.

Same interface as adcc/projection.py:33-236: ``SubspacePartitioning(mospaces, core_orbitals,
outer_virtuals)`` splits o1 into valence ``o`` / core ``c`` and v1 into inner ``v`` / outer ``w``;
``Projector(subspaces, partitioning, blocks_to_keep)`` zeroes every partition block of a tensor that
is not listed.  The projector is a 0/1 mask tensor in HBM (built once from per-axis membership
vectors) applied by the elementwise-multiply kernel.
"""
import itertools
import warnings

import numpy as np

from .MoSpaces import expand_spaceargs
from .Tensor import Tensor


class SubspacePartitioning:
    def __init__(self, mospaces, core_orbitals=None, outer_virtuals=None, partitions=None, aliases=None):
        if mospaces.has_core_occupied_space or "v2" in mospaces.subspaces_virtual:
            raise ValueError("Cannot use MoSpaces with core-valence separation or frozen virtuals.")
        assert mospaces.subspaces_occupied == ["o1"]
        if not partitions or not aliases:
            if core_orbitals is None and outer_virtuals is None:
                raise ValueError("One of core_orbitals or outer_virtuals should be passed.")
            noa, nva = mospaces.n_orbs_alpha("o1"), mospaces.n_orbs_alpha("v1")
            core = expand_spaceargs(noa, core_orbitals=core_orbitals)["core_orbitals"]
            fv = expand_spaceargs(nva, frozen_virtual=outer_virtuals)["frozen_virtual"]
            occ = [i for i in range(mospaces.n_orbs("o1")) if i not in core]
            av = [i for i in range(mospaces.n_orbs("v1")) if i not in fv]
            assert occ and av
            partitions = {"o1.1": occ, "v1.1": av}
            aliases = {"o": "o1.1", "v": "v1.1"}
            if core:
                partitions["o1.2"] = core
                aliases["c"] = "o1.2"
            if fv:
                partitions["v1.2"] = fv
                aliases["w"] = "v1.2"
        else:
            assert core_orbitals is None and outer_virtuals is None
        self.partitions, self.aliases, self.mospaces = partitions, aliases, mospaces
        self.keeps_spin_symmetry = mospaces.restricted
        for label, indices in partitions.items():
            if not self.keeps_spin_symmetry:
                break
            nalpha = mospaces.n_orbs_alpha(label[:2])
            alpha_part = [i for i in indices if i < nalpha]
            beta_part = [i - nalpha for i in indices if i >= nalpha]
            self.keeps_spin_symmetry = alpha_part == beta_part
        if mospaces.restricted and not self.keeps_spin_symmetry:
            warnings.warn("For restricted references the case of a space partitioning, which breaks "
                          "spin has not been tested. You are on your own.")

    def list_space_partitions(self, space):
        parts = [s for s in self.partitions if s.startswith(space + ".")]
        return sorted(key for key, value in self.aliases.items() if value in parts)

    def get_partition(self, label):
        return self.partitions[self.aliases[label]]


class Projector:
    def __init__(self, subspaces, partitioning, blocks_to_keep):
        subspaces = list(subspaces)
        if len(subspaces) == 3 or len(subspaces) > 4 or (len(subspaces) > 2 and subspaces[1] == subspaces[2]):
            raise NotImplementedError(f"Projector not implemented for subspaces {subspaces}")

        def normalise_block(labels):
            labels = list(labels)
            if len(subspaces) > 1 and subspaces[0] == subspaces[1] and labels[0] > labels[1]:
                labels[0], labels[1] = labels[1], labels[0]
            if len(subspaces) == 4 and subspaces[2] == subspaces[3] and labels[2] > labels[3]:
                labels[2], labels[3] = labels[3], labels[2]
            return labels

        blocks_to_keep = ["".join(normalise_block(blk)) for blk in blocks_to_keep]
        parts = [partitioning.list_space_partitions(sp) for sp in subspaces]
        known = {"".join(normalise_block(blk)) for blk in itertools.product(*parts)}
        for blk in blocks_to_keep:
            if blk not in known:
                raise ValueError(f"Partition block {blk} not known.")
        ms = partitioning.mospaces
        member = {}
        for k, sp in enumerate(subspaces):
            for lab in parts[k]:
                vec = np.zeros(ms.n_orbs(sp))
                vec[partitioning.get_partition(lab)] = 1.0
                member[(k, lab)] = vec
        shape = tuple(ms.n_orbs(sp) for sp in subspaces)
        mask = np.zeros(shape)
        nd = len(subspaces)
        swaps = [tuple(range(nd))]
        if nd >= 2 and subspaces[0] == subspaces[1]:
            swaps += [tuple([1, 0] + list(range(2, nd)))]
        if nd == 4 and subspaces[2] == subspaces[3]:
            swaps += [tuple(list(p[:2]) + [3, 2]) for p in list(swaps)]
        for blk in blocks_to_keep:
            for perm in swaps:
                labs = [blk[p] for p in perm]
                if any((k, lab) not in member for k, lab in enumerate(labs)):
                    continue
                term = member[(0, labs[0])]
                for k in range(1, nd):
                    term = np.multiply.outer(term, member[(k, labs[k])])
                mask = np.maximum(mask, term)
        self.subspaces, self.partitioning, self.blocks_to_keep = subspaces, partitioning, blocks_to_keep
        self.kernel = Tensor.from_ndarray(ms, subspaces, mask)
        self.kernel.set_immutable()

    def matvec(self, v):
        return self.kernel * v

    apply = matvec

    def __matmul__(self, other):
        if isinstance(other, Tensor):
            return self.matvec(other)
        if isinstance(other, list) and all(isinstance(e, Tensor) for e in other):
            return [self.matvec(o) for o in other]
        return NotImplemented


def transfer_amplitude_cvs_to_full(mospaces_cvs, cvs, symmetry_full, tol=1e-12):
    """Copy a CVS amplitude (spaces o2 v1 / o1 o2 v1 v1) into a full-space amplitude with the
    symmetry ``symmetry_full`` (adcc/projection.py:238-314).  The core orbitals have to be the
    lowest occupied orbitals of each spin (``core_orbitals`` given as an integer).  The ocvv part
    enters twice in the full doubles amplitude (oc and co), hence the factor 1/sqrt(2)."""
    from .AmplitudeVector import AmplitudeVector
    from .Tensor import Tensor
    if cvs.keys() != symmetry_full.keys():
        raise ValueError("Blocks present in CVS and full vector should agree.")
    if not mospaces_cvs.has_core_occupied_space:
        raise ValueError("Should have CVS enabled in mospaces_cvs")
    if not isinstance(cvs, AmplitudeVector):
        raise TypeError("Expected cvs to be an AmplitudeVector.")
    mospaces_full = symmetry_full["ph"].mospaces
    nca, noa = mospaces_cvs.n_orbs_alpha("o2"), mospaces_cvs.n_orbs_alpha("o1")
    nfa = mospaces_full.n_orbs_alpha("o1")
    core, occ_full = list(mospaces_cvs.core_orbitals), list(mospaces_full.occupied_orbitals)
    if core[:nca] != occ_full[:nca] or core[nca:] != occ_full[nfa:nfa + nca]:
        raise NotImplementedError(
            "For transfer CVS to full only the case where the CVS space is contiguous in the full "
            "space has been implemented, i.e. when the `core_orbitals` are chosen as an integer in "
            "the run_adc routines.")
    full = AmplitudeVector(**{block: Tensor(sym) for block, sym in symmetry_full.items()})
    if cvs.ph.subspaces != ["o2", "v1"] or full.ph.subspaces != ["o1", "v1"]:
        raise ValueError("Unexpected subspaces of the singles blocks")
    # where the core (alpha, beta) and valence (alpha, beta) orbitals sit on a full occupied axis
    c_a, c_b = slice(0, nca), slice(nfa, nfa + nca)
    o_a, o_b = slice(nca, nfa), slice(nfa + nca, None)
    src_ph = cvs.ph.to_ndarray()
    dst_ph = np.zeros(full.ph.shape)
    dst_ph[c_a] = src_ph[:nca]
    dst_ph[c_b] = src_ph[nca:]
    full.ph.set_from_ndarray(dst_ph, tol)
    if "pphh" in cvs:
        if cvs.pphh.subspaces != ["o1", "o2", "v1", "v1"] or full.pphh.subspaces != ["o1", "o1", "v1", "v1"]:
            raise ValueError("Unexpected subspaces of the doubles blocks")
        src = cvs.pphh.to_ndarray() / np.sqrt(2)
        dst = np.zeros(full.pphh.shape)
        for o_full, o_cvs in ((o_a, slice(0, noa)), (o_b, slice(noa, None))):
            for c_full, c_cvs in ((c_a, slice(0, nca)), (c_b, slice(nca, None))):
                part = src[o_cvs, c_cvs]
                dst[o_full, c_full] = part
                dst[c_full, o_full] = -part.transpose(1, 0, 2, 3)
        full.pphh.set_from_ndarray(dst, tol)
    return full


def transfer_cvs_to_full(state_matrix_cvs, matrix_full, vector=None, kind=None, spin_change=0,
                         spin_block_symmetrisation="none"):
    """Transfer ``vector`` (or a list of vectors) from the CVS space to the full space of
    ``matrix_full`` (adcc/projection.py:317-359); ``state_matrix_cvs`` is the CVS matrix or the
    ExcitedStates holding the solved CVS excitations (then the spin symmetry is taken from it).
    Typical use: CVS eigenvectors as guesses for a projected full matrix."""
    from .AdcMatrix import AdcMatrixlike
    from .guess import guess_kwargs_kind, guess_symmetries
    if kind is None and hasattr(state_matrix_cvs, "kind"):
        kind = state_matrix_cvs.kind
    if spin_change is None and spin_block_symmetrisation is None:
        if kind is None:
            raise ValueError("kind needs to be given if first argument is not an ExcitedStates "
                             "object and spin symmetry setup is not explicitly given.")
        return transfer_cvs_to_full(state_matrix_cvs, matrix_full, vector, kind, **guess_kwargs_kind(kind))
    if vector is None:
        if hasattr(state_matrix_cvs, "excitation_vector"):
            vector = state_matrix_cvs.excitation_vector
        else:
            raise ValueError("vector needs to be given if first argument is not an ExcitedStates object.")
    if isinstance(vector, list):
        return [transfer_cvs_to_full(state_matrix_cvs, matrix_full, v, kind, **guess_kwargs_kind(kind))
                for v in vector]
    if isinstance(state_matrix_cvs, AdcMatrixlike):
        mospaces_cvs = state_matrix_cvs.mospaces
    else:
        mospaces_cvs = state_matrix_cvs.matrix.mospaces
    sym = guess_symmetries(matrix_full, spin_change=spin_change,
                           spin_block_symmetrisation=spin_block_symmetrisation)
    return transfer_amplitude_cvs_to_full(mospaces_cvs, vector, sym)
