"""``libadcc`` on the B200 backend: the module surface adcc's own Python layer imports.

The reference's seam is the pybind11 module ``libadcc`` (libadcc_src/pyiface/ExportAdcc.cc:41-71,
stub libadcc.pyi) whose single tensor implementation is libtensor (libadcc_src/TensorImpl.cc).
This package presents the same names on top of adcc_b200 (device tensors + the C ABI of
include/adccb.h), so that the UNMODIFIED ``adcc`` package of the reference - its working
equations adcc/adc_pp/matrix.py, adcc/LazyMp.py, its solver adcc/solver/davidson.py - runs on the
CUDA kernels:

    PYTHONPATH=<this repo>:<adcc source tree> python -c "import adcc; adcc.adc2(data, n_singlets=3)"

Scope: everything the ADC matvec + Jacobi-Davidson path touches (libadcc.pyi:34-66, 67-178,
180-279, 281-388, 390-539, 541-631, 633-793, 795-936), plus the tensors with an AO-basis axis
(``orbital_coefficients*``, ``make_symmetry_operator_basis``) the reference's operator-integral
import needs, so that its transition dipole moments / oscillator strengths run too.  Triples
symmetries raise NotImplementedError.

adcc additionally needs ``opt_einsum`` (pairwise contraction paths); where that third-party
package is not installed, tools/refshim/ holds a minimal stand-in.
"""
import numpy as _np

import adcc_b200 as _p
from adcc_b200 import functions as _fn
from adcc_b200.DataHfProvider import HartreeFockProvider as _ProductHfProvider
from adcc_b200.memory_pool import AdcMemory  # noqa: F401
from adcc_b200.MoSpaces import MoIndexTranslation as _ProductMoIndexTranslation
from adcc_b200.MoSpaces import MoSpaces as _ProductMoSpaces
from adcc_b200.MoSpaces import split_spaces as _split
from adcc_b200.ReferenceState import ReferenceState as _ProductReferenceState
from adcc_b200.Symmetry import Symmetry  # noqa: F401  (libadcc.pyi:541-631)
from adcc_b200.Tensor import Tensor  # noqa: F401      (libadcc.pyi:633-793)

# libtensor semantics of select_n_*: symmetry-unique elements only (see Tensor.select_spin_unique)
Tensor.select_spin_unique = True

__backend__ = dict(_p.__backend__)
__version__ = _p.__version__
__all__ = ["AdcMemory", "HartreeFockProvider", "HartreeFockSolution_i", "MoIndexTranslation", "MoSpaces",
           "ReferenceState", "Symmetry", "Tensor", "amplitude_vector_enforce_spin_kind", "direct_sum",
           "evaluate", "fill_pp_doubles_guesses", "get_n_threads", "get_n_threads_total",
           "linear_combination_strict", "make_symmetry_eri", "make_symmetry_operator",
           "make_symmetry_operator_basis", "make_symmetry_orbital_coefficients",
           "make_symmetry_orbital_energies", "make_symmetry_triples", "set_n_threads",
           "set_n_threads_total", "tensordot", "trace"]


# ------------------------------------------------------------------------------ SCF data interface
class HartreeFockSolution_i(_ProductHfProvider):
    """libadcc_src/HartreeFockSolution_i.hh:34-263 / libadcc.pyi:145-178."""

    @property
    def n_alpha(self):
        occ = self.occupation_f
        return int(round(float(_np.sum(occ[:self.n_orbs_alpha]))))

    @property
    def n_beta(self):
        occ = self.occupation_f
        return int(round(float(_np.sum(occ[self.n_orbs_alpha:]))))

    @property
    def n_orbs_beta(self):
        return self.n_orbs_alpha


class HartreeFockProvider(HartreeFockSolution_i):
    """Base class Python SCF providers derive from (libadcc.pyi:67-143,
    libadcc_src/pyiface/export_HartreeFockProvider.cc:31-200): subclasses override get_* / fill_*."""

    def __init__(self):
        pass

    def has_eri_phys_asym_ffff(self):
        return False

    def has_eri_phys_asym_ffff_inner(self):           # name used by the product's staging code
        return bool(self.has_eri_phys_asym_ffff())

    def get_nuclear_multipole(self, order, gauge_origin=(0.0, 0.0, 0.0)):
        raise NotImplementedError("get_nuclear_multipole")

    def transform_gauge_origin_to_xyz(self, gauge_origin):
        """libadcc.pyi:140: providers that know the molecule translate "origin", "mass_center", ... to
        coordinates; the base class only knows the coordinate origin."""
        if isinstance(gauge_origin, str):
            if gauge_origin == "origin":
                return (0.0, 0.0, 0.0)
            raise NotImplementedError(f"gauge origin {gauge_origin!r} is not known to this provider")
        return tuple(float(x) for x in gauge_origin)


# ------------------------------------------------------------------------------ MO spaces
class MoSpaces:
    """libadcc.pyi:281-388 / libadcc_src/MoSpaces.hh: constructed from explicit index lists."""

    def __init__(self, hfdata, adcmem, core_orbitals, frozen_core, frozen_virtual):
        self._impl = _ProductMoSpaces(hfdata, core_orbitals=list(core_orbitals), frozen_core=list(frozen_core),
                                      frozen_virtual=list(frozen_virtual), explicit_index_lists=True,
                                      max_block_size=getattr(adcmem, "max_block_size", 16) or 16)
        self._adcmem = adcmem

    def n_orbs(self, space="f"): return self._impl.n_orbs(space)
    def n_orbs_alpha(self, space="f"): return self._impl.n_orbs_alpha(space)
    def n_orbs_beta(self, space="f"): return self._impl.n_orbs_beta(space)
    def has_subspace(self, space): return self._impl.has_subspace(space)
    def subspace_name(self, space): return self._impl.subspace_name(space)
    def describe(self): return self._impl.describe()

    has_core_occupied_space = property(lambda self: self._impl.has_core_occupied_space)
    irrep_totsym = property(lambda self: self._impl.irrep_totsym)
    irreps = property(lambda self: list(self._impl.irreps))
    map_block_irrep = property(lambda self: self._impl.map_block_irrep)
    map_block_spin = property(lambda self: self._impl.map_block_spin)
    map_block_start = property(lambda self: self._impl.map_block_start)
    map_index_hf_provider = property(lambda self: self._impl.map_index_hf_provider)
    point_group = property(lambda self: self._impl.point_group)
    restricted = property(lambda self: self._impl.restricted)
    subspaces = property(lambda self: list(self._impl.subspaces))
    subspaces_occupied = property(lambda self: list(self._impl.subspaces_occupied))
    subspaces_virtual = property(lambda self: list(self._impl.subspaces_virtual))


class MoIndexTranslation(_ProductMoIndexTranslation):
    """libadcc.pyi:180-279."""


class _AxisSymmetry(Symmetry):
    """Symmetry of tensors that carry the AO-basis axis "b" (libadcc_src/make_symmetry.cc:275-342,
    428-463): the axis is not an MO subspace, so the extents are given explicitly; no permutational
    symmetry is recorded (the dense store does not need it)."""

    def __init__(self, mospaces, subspaces, shape):
        self.mospaces = mospaces
        self.subspaces = list(subspaces)
        self.space = "".join(self.subspaces)
        self.ndim = len(self.subspaces)
        self.shape = tuple(int(x) for x in shape)
        self.irreps_allowed = ["A"]
        self._perms, self._perm_strings = [], []
        self.spin_block_maps, self.spin_blocks_forbidden = [], []


# ------------------------------------------------------------------------------ reference state
class ReferenceState:
    """libadcc.pyi:390-539 / libadcc_src/ReferenceState.hh: lazy, cached Fock / ERI import."""

    def __init__(self, hfdata, mospaces, symmetry_check_on_import=False):
        self._impl = _ProductReferenceState(hfdata, mospaces=mospaces,
                                            symmetry_check_on_import=symmetry_check_on_import,
                                            import_all_below_n_orbs=None)

    def eri(self, space): return self._impl.eri(space)
    def fock(self, space): return self._impl.fock(space)
    def orbital_energies(self, space): return self._impl.orbital_energies(space)
    def import_all(self): return self._impl.import_all()
    def flush_hf_cache(self): return self._impl.flush_hf_cache()

    def _coeff(self, space, keep):
        """MO coefficients of a subspace as a tensor with an AO-basis axis "b"
        (libadcc_src/ReferenceState.cc:101-334): keep "ab" -> [n_space, 2 n_bas], block diagonal
        (alpha rows [C, 0], beta rows [0, C]); "a" / "b" -> [n_space, n_bas] with the rows of the
        other spin zeroed."""
        impl = self._impl
        if not (isinstance(space, str) and space.endswith("b") and len(space) == 3):
            raise ValueError(f"Invalid space string {space}: an object orbital_coefficients({space}) is not known.")
        sub = space[:2]
        c = _np.asarray(impl._coefficients(space, keep))              # rows of the other spin already zero
        n_bas = c.shape[1]
        if keep == "ab":
            is_alpha = _np.asarray(impl._host_index(sub)) < impl.hf_provider.n_orbs_alpha
            full = _np.zeros((c.shape[0], 2 * n_bas))
            full[is_alpha, :n_bas] = c[is_alpha]
            full[~is_alpha, n_bas:] = c[~is_alpha]
            c = full
        t = Tensor.from_ndarray(impl.mospaces, [sub, "b"], c, _AxisSymmetry(impl.mospaces, [sub, "b"], c.shape))
        t.set_immutable()
        return t

    def orbital_coefficients(self, space): return self._coeff(space, "ab")
    def orbital_coefficients_alpha(self, space): return self._coeff(space, "a")
    def orbital_coefficients_beta(self, space): return self._coeff(space, "b")

    def gauge_origin_to_xyz(self, gauge_origin):
        return self._impl.hf_provider.transform_gauge_origin_to_xyz(gauge_origin)

    def nuclear_quadrupole(self, gauge_origin):
        return _np.asarray(self._impl.hf_provider.get_nuclear_multipole(2, gauge_origin))

    @property
    def cached_eri_blocks(self): return self._impl.cached_eri_blocks

    @cached_eri_blocks.setter
    def cached_eri_blocks(self, keep): self._impl.cached_eri_blocks = keep

    @property
    def cached_fock_blocks(self): return self._impl.cached_fock_blocks

    @cached_fock_blocks.setter
    def cached_fock_blocks(self, keep): self._impl.cached_fock_blocks = keep

    backend = property(lambda self: self._impl.backend)
    conv_tol = property(lambda self: self._impl.conv_tol)
    energy_scf = property(lambda self: self._impl.energy_scf)
    has_core_occupied_space = property(lambda self: self._impl.has_core_occupied_space)
    irreducible_representation = property(lambda self: self._impl.irreducible_representation)
    mospaces = property(lambda self: self._impl.mospaces)
    n_alpha = property(lambda self: self._impl.n_alpha)
    n_beta = property(lambda self: self._impl.n_beta)
    n_orbs = property(lambda self: self._impl.n_orbs)
    n_orbs_alpha = property(lambda self: self._impl.n_orbs_alpha)
    n_orbs_beta = property(lambda self: self._impl.n_orbs_beta)
    nuclear_dipole = property(lambda self: self._impl.nuclear_dipole)
    nuclear_total_charge = property(lambda self: float(_np.asarray(
        self._impl.hf_provider.get_nuclear_multipole(0, "origin")).reshape(-1)[0]))
    restricted = property(lambda self: self._impl.restricted)
    spin_multiplicity = property(lambda self: self._impl.spin_multiplicity)
    timer = property(lambda self: self._impl.timer)


# ------------------------------------------------------------------------------ functions
def tensordot(a, b, axes=2):
    """libadcc_src/pyiface/export_Tensor.cc:198-218."""
    return _fn.tensordot(a, b, axes=axes)


def direct_sum(a, b):
    """out[p..., q...] = a[p...] + b[q...]  (export_Tensor.cc:236-240)."""
    return _fn.direct_sum("a,b", a, b) if False else _direct_sum(a, b)


def _direct_sum(a, b):
    from adcc_b200 import backend as be
    data = be.direct_sum(a._contig(), b._contig(), 1.0, 1.0)
    return Tensor._wrap(a.mospaces, list(a.subspaces) + list(b.subspaces), data)


def evaluate(t):
    return t.evaluate()


def _spin_flip_factor(sym):
    """+1 / -1 if every spin-block map of ``sym`` sends a block to its complete spin flip
    (aa <-> bb, abab <-> baba, ...) with one common factor - the maps of restricted singlet / triplet
    vectors (adcc/guess/guess_zero.py:121-161) -, else None."""
    maps = getattr(sym, "spin_block_maps", None) if sym is not None else None
    if not maps:
        return None
    flip = str.maketrans("ab", "ba")
    facs = set()
    for frm, to, fac in maps:
        if to != frm.translate(flip):
            return None
        facs.add(float(fac))
    return facs.pop() if len(facs) == 1 and abs(next(iter(facs))) == 1.0 else None


def linear_combination_strict(coefficients, tensors):
    """export_Tensor.cc:260-277.

    libtensor stores only the canonical spin block of a tensor with spin-block maps, so any linear
    combination that involves such a tensor HAS that symmetry exactly.  The dense store computes
    the mapped blocks independently; their rounding noise is harmless at O(1) but dominates a
    Davidson residual of 1e-9, and the reference's solver - which relies on the structural symmetry
    for the singles block (adcc/solver/explicit_symmetrisation.py:80-98 touches the doubles only) -
    then lets lower-lying states of the other spin kind grow in.  The result is therefore projected
    onto the maps of its operands (adccb_spin_project), which is what libtensor's storage does.
    Operands without symmetry metadata (results of contractions, e.g. A x) are taken to share the
    maps: on this dense store a contraction of symmetric tensors is symmetric up to rounding, and a
    calculation whose vectors are not (spin-flip, kind "any") carries no maps at all."""
    tensors = list(tensors)
    res = _fn.linear_combination_strict(coefficients, tensors)
    sym = next((t.symmetry for t in tensors if _spin_flip_factor(getattr(t, "symmetry", None)) is not None), None)
    if sym is not None and type(res) is Tensor and res.ndim in (2, 4):
        from adcc_b200 import backend as _be
        n_alpha = [res.mospaces.n_orbs_alpha(sp) for sp in res.subspaces]
        res = Tensor._wrap(res.mospaces, res.subspaces,
                               _be.spin_project(res._contig(), n_alpha, _spin_flip_factor(sym)),
                               res.symmetry if res.symmetry is not None else sym)
    return res


def trace(*args):
    """trace("ijij", t) or trace(matrix)  (export_Tensor.cc:242-258)."""
    if len(args) == 1:
        return _fn.trace("ii", args[0])
    return _fn.trace(args[0], args[1])


def amplitude_vector_enforce_spin_kind(tensor, block, spin_kind):
    """libadcc_src/amplitude_vector_enforce_spin_kind.cc:33-170."""
    from adcc_b200.solver.explicit_symmetrisation import amplitude_vector_enforce_spin_kind as impl
    return impl(tensor, block, spin_kind)


def fill_pp_doubles_guesses(guesses_d, mospaces, df02, df13, spin_change_twice, degeneracy_tolerance):
    """libadcc_src/fill_pp_doubles_guesses.cc:26-71."""
    from adcc_b200.guess.guesses_from_diagonal import _fill_pp_doubles_guesses
    return _fill_pp_doubles_guesses(guesses_d, mospaces, df02, df13, spin_change_twice, degeneracy_tolerance)


def _eri_like_symmetry(mospaces, space):
    """libadcc_src/make_symmetry.cc:277-356: <pq||rs> = -<qp||rs> = -<pq||sr> = <rs||pq> wherever the
    index pairs share a subspace."""
    ss = _split(space)
    perms = ["ijkl"]
    if ss[0] == ss[1]:
        perms.append("-jikl")
    if ss[2] == ss[3]:
        perms.append("-ijlk")
    if ss[0] == ss[2] and ss[1] == ss[3]:
        perms.append("klij")
    return Symmetry(mospaces, space, permutations=perms if len(perms) > 1 else None)


def make_symmetry_eri(mospaces, space):
    return _eri_like_symmetry(mospaces, space)


def make_symmetry_orbital_energies(mospaces, space):
    return Symmetry(mospaces, space)


def make_symmetry_operator(mospaces, space, symmetry="nosymmetry", cartesian_transformation="1"):
    """libadcc_src/make_symmetry.cc:358-425 (one- and two-particle operator blocks, C1 point group)."""
    ss = _split(space)
    perms = None
    if len(ss) == 2 and ss[0] == ss[1]:
        if symmetry == "hermitian":
            perms = ["ij", "ji"]
        elif symmetry == "antihermitian":
            perms = ["ij", "-ji"]
    elif len(ss) == 4 and symmetry == "hermitian":
        return _eri_like_symmetry(mospaces, space)
    return Symmetry(mospaces, space, permutations=perms)


def _not_on_path(name):
    def fail(*args, **kwargs):
        raise NotImplementedError(f"libadcc.{name} belongs to the property infrastructure (AO-basis tensors / "
                                  "triples), which is outside the ADC matvec + Davidson path of this backend")
    fail.__name__ = name
    return fail


def make_symmetry_operator_basis(mospaces, n_bas, symmetry="nosymmetry", n_particle_op=1, blocks="ab"):
    """Operator in the AO basis, block diagonal over (alpha | beta) copies of the n_bas functions
    (libadcc_src/make_symmetry.cc:428-463)."""
    n = 2 * n_bas if blocks == "ab" else n_bas
    rank = 2 * int(n_particle_op)
    return _AxisSymmetry(mospaces, ["b"] * rank, (n,) * rank)


def make_symmetry_orbital_coefficients(mospaces, space, n_bas, blocks="ab"):
    if not space.endswith("b") or len(space) != 3:
        raise ValueError(f"Expect exactly a two-dimensional space string like 'o1b', not {space}.")
    n = 2 * n_bas if blocks == "ab" else n_bas
    return _AxisSymmetry(mospaces, [space[:2], "b"], (mospaces.n_orbs(space[:2]), n))


make_symmetry_triples = _not_on_path("make_symmetry_triples")


def get_n_threads():
    return _p.get_n_threads()


def set_n_threads(n):
    return _p.set_n_threads(n)


def get_n_threads_total():
    return _p.get_n_threads()


def set_n_threads_total(n):
    return _p.set_n_threads(n)
