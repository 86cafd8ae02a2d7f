"""Guess vectors from the smallest diagonal elements.

Host-side, one-shot logic mirroring adcc/guess/guesses_from_diagonal.py:33-305 for singles
(``find_smallest_matching_elements``, spin-adapted +/- pairs) and, for doubles,
libadcc_src/fill_pp_doubles_guesses.cc:26-71 -> libadcc_src/guess/adc_guess_d.cc:205-543 and
guess/index_group_d.cc:38-94 (smallest df02 + df13 sums grouped by spatial index, spin-adapted
linear combinations).  Elements are written through ``Tensor.__setitem__`` which applies the
symmetry orbit of the trial vector (permutational signs and spin-block maps).
"""
import itertools
from itertools import groupby

import numpy as np

from ..AdcMatrix import AdcMatrixlike
from ..functions import evaluate
from ..MoSpaces import MoIndexTranslation
from .guess_zero import guess_zero


def guesses_from_diagonal(matrix, n_guesses, block="ph", spin_change=0,
                          spin_block_symmetrisation="none", degeneracy_tolerance=1e-14,
                          max_diagonal_value=1000):
    if not isinstance(matrix, AdcMatrixlike):
        raise TypeError("matrix needs to be of type AdcMatrixlike")
    if spin_block_symmetrisation not in ["none", "symmetric", "antisymmetric"]:
        raise ValueError(f"Invalid value for spin_block_symmetrisation: {spin_block_symmetrisation}")
    if spin_block_symmetrisation != "none" and not matrix.reference_state.restricted:
        raise ValueError("spin_block_symmetrisation != none is only valid for ADC calculations on "
                         "top of restricted reference states.")
    if int(spin_change * 2) / 2 != spin_change:
        raise ValueError(f"Only integer or half-integer spin_change is allowed. You passed {spin_change}")
    if block not in matrix.axis_blocks:
        raise ValueError(f"The passed ADC matrix does not have the block '{block}.'")
    if n_guesses == 0:
        return []
    if block == "ph":
        fn = guesses_from_diagonal_singles
    elif block == "pphh":
        fn = guesses_from_diagonal_doubles
    else:
        raise ValueError(f"Don't know how to generate guesses for block {block}")
    return fn(matrix, n_guesses, spin_change, spin_block_symmetrisation, degeneracy_tolerance,
              max_diagonal_value)


class TensorElement:
    def __init__(self, motrans, index, value):
        self.index = tuple(index)
        self.subspaces = motrans.subspaces
        self.value = value
        self.spin_block, self.block_index_spatial, self.inblock_index = motrans.split_spin(tuple(index))

    @property
    def spin_change(self):
        m = {("o", "a"): -0.5, ("o", "b"): +0.5, ("v", "a"): +0.5, ("v", "b"): -0.5}
        return int(sum(m[(space[0], spin)] for space, spin in zip(self.subspaces, self.spin_block)))

    def __repr__(self):
        return f"({self.index}  {self.value})"


def find_smallest_matching_elements(predicate, tensor, motrans, n_elements, degeneracy_tolerance=1e-12):
    n_searched_for = max(10, 2 * n_elements + 6)
    while True:
        res = []
        found = tensor.select_n_min(n_searched_for)
        for index, value in found:
            telem = TensorElement(motrans, index, value)
            if predicate(telem):
                res.append(telem)
        if len(res) >= n_elements or len(found) < n_searched_for:
            break
        n_searched_for *= 2
    if len(res) == 0:
        return []

    def telem_nospin(t):
        return (t.value, t.block_index_spatial, t.inblock_index)
    res = sorted(res, key=telem_nospin)
    istart = 0
    for i in range(len(res)):
        if abs(res[istart].value - res[i].value) > degeneracy_tolerance:
            avg = np.average([r.value for r in res[istart:i]])
            for j in range(istart, i):
                res[j].value = avg
            istart = i
    avg = np.average([r.value for r in res[istart:]])
    for j in range(istart, len(res)):
        res[j].value = avg
    if len(res) > n_elements:
        return [t for t in res if t.value <= res[n_elements - 1].value]
    return res


def guesses_from_diagonal_singles(matrix, n_guesses, spin_change=0, spin_block_symmetrisation="none",
                                  degeneracy_tolerance=1e-14, max_diagonal_value=1000):
    motrans = MoIndexTranslation(matrix.mospaces, matrix.axis_spaces["ph"])
    if n_guesses == 0:
        return []
    ret = [guess_zero(matrix, spin_change=spin_change,
                      spin_block_symmetrisation=spin_block_symmetrisation)
           for _ in range(n_guesses)]
    sym_ph = ret[0].ph.symmetry

    def pred_singles(telem):
        # with spin-block maps only the canonical (alpha) block is a free element
        canonical = sym_ph is None or sym_ph.empty or sym_ph.is_canonical(telem.index)
        return (ret[0].ph.is_allowed(telem.index) and canonical
                and telem.spin_change == spin_change and abs(telem.value) <= max_diagonal_value)

    elements = find_smallest_matching_elements(pred_singles, matrix.diagonal().ph, motrans,
                                               n_guesses, degeneracy_tolerance=degeneracy_tolerance)
    if len(elements) == 0:
        return []

    def telem_nospin(t):
        return (t.value, t.block_index_spatial, t.inblock_index)

    ivec = 0
    for _, group in groupby(elements, key=telem_nospin):
        if ivec >= len(ret):
            break
        group = list(group)
        if len(group) == 1:
            ret[ivec].ph[group[0].index] = 1.0
            ivec += 1
        elif len(group) == 2:
            ret[ivec].ph[group[0].index] = 1
            ret[ivec].ph[group[1].index] = 1
            ivec += 1
            if ivec < n_guesses:
                ret[ivec].ph[group[0].index] = +1
                ret[ivec].ph[group[1].index] = -1
                ivec += 1
        else:
            raise AssertionError("group size > 3 should not occur when setting up single guesse.")
    assert ivec <= n_guesses
    return [evaluate(v / np.sqrt(v @ v)) for v in ret[:ivec]]


# ------------------------------------------------------------------------------ doubles
def _fill_pp_doubles_guesses(guesses_d, mospaces, df02, df13, spin_change_twice, degeneracy_tolerance):
    """Port of adc_guess_d (see module docstring).  Returns the number of guesses built."""
    n_guesses = len(guesses_d)
    if n_guesses == 0:
        return 0
    if spin_change_twice not in (0, -2):
        raise NotImplementedError("spin_change == -1 has not been tested.")
    first = guesses_d[0]
    spaces = first.subspaces
    sym = first.symmetry
    motrans = MoIndexTranslation(mospaces, spaces)
    sym_o = spaces[0] == spaces[1] and sym is not None and any(
        axes == (1, 0, 2, 3) and s < 0 for axes, s in sym._perms)
    sym_v = spaces[2] == spaces[3] and sym is not None and any(
        axes == (0, 1, 3, 2) and s < 0 for axes, s in sym._perms)
    spin = 0
    if sym is not None:
        for frm, to, fac in sym.spin_block_maps:
            if {frm, to} == {"aaaa", "bbbb"}:
                spin = 1 if fac == 1 else 3
    forbidden = set(sym.spin_blocks_forbidden) if sym is not None else set()
    d1, d2 = df02.to_ndarray(), df13.to_ndarray()
    n_alpha = [mospaces.n_orbs_alpha(s) for s in spaces]
    n_ablocks = [sum(1 for x in mospaces.map_block_spin[s] if x == "a") for s in spaces]

    def smallest(d, ns):
        flat = np.argsort(d, axis=None, kind="stable")[:ns]
        return [(tuple(int(x) for x in np.unravel_index(f, d.shape)), float(d.flat[f])) for f in flat]

    def canon(spm, spi, idx):
        spm, spi, idx = list(spm), list(spi), list(idx)
        for (p, q), on in (((0, 1), sym_o), ((2, 3), sym_v)):
            if on and ((spi[p] == spi[q] and idx[p] > idx[q]) or spi[p] > spi[q]):
                for arr in (spm, spi, idx):
                    arr[p], arr[q] = arr[q], arr[p]
        return tuple(spm), tuple(spi), tuple(idx)

    ns, size, max_reached, groups = n_guesses, 0, False, []
    while size < n_guesses and not max_reached:
        ns *= 2
        ilx, ily = smallest(d1, ns), smallest(d2, ns)
        max_reached = len(ilx) < ns
        groups = []
        for (ia, va), (ib, vb) in itertools.product(ilx, ily):
            if sym_o and ia[0] == ib[0]:
                continue
            if sym_v and ia[1] == ib[1]:
                continue
            full = [ia[0], ib[0], ia[1], ib[1]]
            if sym_o and full[0] > full[1]:
                full[0], full[1] = full[1], full[0]
            if sym_v and full[2] > full[3]:
                full[2], full[3] = full[3], full[2]
            spin_str, spi, idx = motrans.split_spin(tuple(full))
            ms = 0
            for k in range(4):
                ms += 1 if ((k < 2) == (spin_str[k] == "b")) else -1
            if ms != spin_change_twice:
                continue
            if spin_str in forbidden:
                continue
            spm = [c == "b" for c in spin_str]
            if spin in (1, 3) and spm[0]:
                spm = [not x for x in spm]
            spm, spi, idx = canon(spm, spi, idx)
            val = va + vb
            for g in groups:
                if abs(val - g[0]) < degeneracy_tolerance and g[1] == spi and g[2] == idx:
                    g[3].add(spm)
                    break
            else:
                groups.append([val, spi, idx, {spm}])
        groups.sort(key=lambda g: g[0])
        size = 0
        for g in groups:
            idx = g[2]
            so = sym_o and idx[0] == idx[1]
            sv = sym_v and idx[2] == idx[3]
            if spin == 0:
                size += len(g[3])
            elif spin == 1:
                size += 1 if (so or sv) else 2
            elif not (so and sv):
                size += 1 if (so or sv) else 3

    def to_index(spm, spi, idx):
        out = []
        for k in range(4):
            blk = spi[k] + (n_ablocks[k] if spm[k] else 0)
            out.append(mospaces.map_block_start[spaces[k]][blk] + idx[k])
        return tuple(out)

    Al, Be = False, True
    n_built = 0
    for val, spi, idx, states in groups:
        if n_built >= n_guesses:
            break
        so_sp = sym_o and spi[0] == spi[1] and idx[0] == idx[1]
        sv_sp = sym_v and spi[2] == spi[3] and idx[2] == idx[3]
        if spin == 0:
            st = sorted(states, key=lambda m: sum(int(x) << (3 - k) for k, x in enumerate(m)))
            if len(st) == 2:
                lv = [[(st[0], 1.0), (st[1], 1.0)], [(st[0], 1.0), (st[1], -1.0)]]
            elif len(st) == 6:
                coeff = [[2, 1, 1, 1, 1, 2], [0, 1, -1, -1, 1, 0], [1, 0, 0, 0, 0, -1],
                         [0, 1, 1, -1, -1, 0], [0, 1, -1, 1, -1, 0], [1, -1, -1, -1, -1, 1]]
                lv = [[(st[j], float(c[j])) for j in range(6) if c[j] != 0] for c in coeff]
            else:
                lv = [[(s, 1.0)] for s in st]
        elif spin == 1:
            if so_sp or sv_sp:
                one = [((Al, Be, Al, Be), 1.0)]
                if spi[0] != spi[1] and idx[0] != idx[1]:
                    one.append(((Be, Al, Al, Be), -1.0))
                elif spi[2] != spi[3] and idx[2] != idx[3]:
                    one.append(((Al, Be, Be, Al), -1.0))
                lv = [one]
            else:
                lv = [[((Al, Al, Al, Al), 2.0), ((Al, Be, Al, Be), 1.0), ((Al, Be, Be, Al), 1.0),
                       ((Be, Al, Be, Al), 1.0), ((Be, Al, Al, Be), 1.0)],
                      [((Al, Be, Al, Be), 1.0), ((Al, Be, Be, Al), -1.0), ((Be, Al, Be, Al), 1.0),
                       ((Be, Al, Al, Be), -1.0)]]
        else:
            if so_sp and sv_sp:
                continue
            if so_sp or sv_sp:
                one = [((Al, Be, Al, Be), 1.0)]
                if spi[0] != spi[1] and idx[0] != idx[1]:
                    one.append(((Be, Al, Al, Be), 1.0))
                elif spi[2] != spi[3] and idx[2] != idx[3]:
                    one.append(((Al, Be, Be, Al), 1.0))
                lv = [one]
            else:
                lv = [[((Al, Al, Al, Al), 1.0)],
                      [((Al, Be, Al, Be), 1.0), ((Al, Be, Be, Al), -1.0), ((Be, Al, Be, Al), -1.0),
                       ((Be, Al, Al, Be), 1.0)],
                      [((Al, Be, Al, Be), 1.0), ((Al, Be, Be, Al), 1.0), ((Be, Al, Be, Al), -1.0),
                       ((Be, Al, Al, Be), -1.0)]]
        for elems in lv:
            if n_built >= n_guesses:
                break
            t = guesses_d[n_built]
            t.fill(0.0)
            for spm, c in elems:
                index = to_index(spm, spi, idx)
                if t.is_allowed(index):
                    t[index] = c
            norm = t.dot(t)
            if norm == 0.0:
                continue
            if norm != 1.0:
                t._scale_inplace(1.0 / np.sqrt(norm))     # the caller's tensor object is filled (t *= c would rebind)
            n_built += 1
    return n_built


def guesses_from_diagonal_doubles(matrix, n_guesses, spin_change=0, spin_block_symmetrisation="none",
                                  degeneracy_tolerance=1e-14, max_diagonal_value=1000):
    if n_guesses == 0:
        return []
    ret = [guess_zero(matrix, spin_change=spin_change,
                      spin_block_symmetrisation=spin_block_symmetrisation)
           for _ in range(n_guesses)]
    spaces_d = matrix.axis_spaces["pphh"]
    df02 = matrix.ground_state.df(spaces_d[0] + spaces_d[2])
    df13 = matrix.ground_state.df(spaces_d[1] + spaces_d[3])
    guesses_d = [gv.pphh for gv in ret]
    spin_change_twice = int(spin_change * 2)
    assert spin_change_twice / 2 == spin_change
    n_found = _fill_pp_doubles_guesses(guesses_d, matrix.mospaces, df02, df13, spin_change_twice,
                                       degeneracy_tolerance)
    ret = ret[:n_found]
    diagonal_elements = [r.pphh.dot(matrix.diagonal().pphh * r.pphh) for r in ret]
    return [ret[i] for (i, elem) in enumerate(diagonal_elements) if elem <= max_diagonal_value]
