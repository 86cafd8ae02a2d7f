"""Zero trial vectors carrying the symmetry every Davidson vector has
(adcc/guess/guess_zero.py:28-171)."""
from ..AdcMatrix import AdcMatrixlike
from ..AmplitudeVector import AmplitudeVector
from ..Symmetry import Symmetry
from ..Tensor import Tensor

_FORBIDDEN_D = ["aabb", "bbaa", "aaab", "aaba", "abaa", "baaa", "abbb", "babb", "bbab", "bbba"]


def guess_zero(matrix, spin_change=0, spin_block_symmetrisation="none"):
    packed = getattr(matrix, "doubles_storage", "dense") == "packed"

    def make(block, sym):
        if packed and block == "pphh":
            from ..packed import PackedDoubles, SlabDoubles
            if matrix.is_core_valence_separated:
                from ..packed_cvs import CvsPackedDoubles
                return CvsPackedDoubles.zeros(sym.mospaces, sym.subspaces, sym)
            if getattr(matrix, "vector_distribution", None) == "slab":
                return SlabDoubles.zeros(sym.mospaces, sym.subspaces, matrix._engine.layout, sym)
            return PackedDoubles.zeros(sym.mospaces, sym.subspaces, sym)
        return Tensor(sym)
    return AmplitudeVector(**{
        block: make(block, sym) for block, sym in guess_symmetries(
            matrix, spin_change=spin_change,
            spin_block_symmetrisation=spin_block_symmetrisation).items()})


def guess_symmetries(matrix, spin_change=0, spin_block_symmetrisation="none"):
    if not isinstance(matrix, AdcMatrixlike):
        raise TypeError("matrix needs to be of type AdcMatrixlike")
    if spin_block_symmetrisation not in ["none", "symmetric", "antisymmetric"]:
        raise ValueError(f"Invalid value for spin_block_symmetrisation: {spin_block_symmetrisation}")
    if spin_block_symmetrisation != "none" and not matrix.reference_state.restricted:
        raise ValueError("spin_block_symmetrisation != none is only valid for ADC calculations on "
                         "top of restricted reference states.")
    if int(spin_change * 2) / 2 != spin_change:
        raise ValueError(f"Only integer or half-integer spin_change is allowed. You passed {spin_change}")
    max_spin_change = 2 if "pphh" in matrix.axis_blocks else 1
    if spin_change > max_spin_change:
        raise ValueError(f"spin_change may only be in the range [{-max_spin_change}, "
                         f"{max_spin_change}] and not {spin_change}.")
    if spin_change != 0 and spin_block_symmetrisation != "none":
        raise NotImplementedError("spin_symmetrisation != 'none' only implemented for spin_change == 0")
    fac = {"symmetric": 1, "antisymmetric": -1}.get(spin_block_symmetrisation)
    symmetries = {}
    if "ph" in matrix.axis_blocks:
        sym = Symmetry(matrix.mospaces, "".join(matrix.axis_spaces["ph"]))
        if fac is not None:
            sym.spin_block_maps = [("aa", "bb", fac)]
            sym.spin_blocks_forbidden = ["ab", "ba"]
        symmetries["ph"] = sym
    if "pphh" in matrix.axis_blocks:
        spaces_d = matrix.axis_spaces["pphh"]
        sym = Symmetry(matrix.mospaces, "".join(spaces_d))
        if fac is not None:
            sym.spin_block_maps = [("aaaa", "bbbb", fac), ("abab", "baba", fac), ("abba", "baab", fac)]
            sym.spin_blocks_forbidden = list(_FORBIDDEN_D)
        permutations = ["ijab"]
        if spaces_d[0] == spaces_d[1]:
            permutations.append("-jiab")
        if spaces_d[2] == spaces_d[3]:
            permutations.append("-ijba")
        if len(permutations) > 1:
            sym.permutations = permutations
        symmetries["pphh"] = sym
    return symmetries
