"""Oracle guess vectors (test infrastructure, see oracle/__init__.py).

Restates
  * adcc/guess/guess_zero.py:113-171 (spin-block maps +-1, forbidden spin blocks,
    -jiab / -ijba permutational symmetry of trial vectors),
  * adcc/guess/guesses_from_diagonal.py:98-271 (singles guesses from the smallest diagonal
    entries, spin-adapted +- pairs for kind="any"),
  * adcc/guess/guesses_from_diagonal.py:274-305 -> libadcc_src/fill_pp_doubles_guesses.cc:26-71
    -> libadcc_src/guess/adc_guess_d.cc:205-543, guess/index_group_d.cc:38-94 (doubles guesses
    from the smallest df02+df13 sums, spin-adapted combinations),
  * libadcc_src/MoSpaces/construct_blocks.cc:25-76 (tiling each spin half into blocks <= 16).
Dense arrays: "setting an element" of a symmetric block tensor is modelled by writing the
whole symmetry orbit of the element (permutations with sign, alpha<->beta map with factor).
"""
import itertools
import numpy as np

MAX_BLOCK = 16  # libadcc_src/AdcMemory.hh:43


def block_starts(n_alpha, n_total, max_block=MAX_BLOCK):
    """construct_blocks.cc:25-76 for crude starts [0, n_alpha]."""
    starts = []
    for lo, hi in ((0, n_alpha), (n_alpha, n_total)):
        ln = hi - lo
        if ln == 0:
            continue
        nb = (ln + max_block - 1) // max_block
        bs, nlarge = ln // nb, ln % nb
        pos = lo
        for ib in range(nb):
            starts.append(pos)
            pos += bs + 1 if ib < nlarge else bs
    return starts


class AxisBlocks:
    def __init__(self, n_alpha, n_total):
        self.n_alpha, self.n = n_alpha, n_total
        self.starts = block_starts(n_alpha, n_total)
        self.n_alpha_blocks = sum(1 for s in self.starts if s < n_alpha)

    def split(self, i):
        """index -> (is_beta, spatial block index, in-block index)."""
        b = max(k for k, s in enumerate(self.starts) if s <= i)
        beta = i >= self.n_alpha
        return beta, (b - self.n_alpha_blocks if beta else b), i - self.starts[b]


def kind_to_sym(kind):
    return {"singlet": "symmetric", "triplet": "antisymmetric", "any": "none",
            "spin_flip": "none"}[kind]


def _axes(matrix):
    hf = matrix.hf
    occ = "o2" if matrix.cvs else "o1"
    return {"ph": [occ, "v1"], "pphh": ["o1", occ, "v1", "v1"]}


def orbit(index, value, spaces, n_alpha, n_tot, sbs):
    """All symmetry-equivalent (index, value) of a trial-vector element."""
    fac = {"symmetric": 1.0, "antisymmetric": -1.0}.get(sbs)
    gens = []
    if len(index) == 4:
        if spaces[0] == spaces[1]:
            gens.append(lambda ix: ((ix[1], ix[0], ix[2], ix[3]), -1.0))
        if spaces[2] == spaces[3]:
            gens.append(lambda ix: ((ix[0], ix[1], ix[3], ix[2]), -1.0))
    if fac is not None:
        def flip(ix):
            new = tuple(i + n_alpha[k] if i < n_alpha[k] else i - n_alpha[k]
                        for k, i in enumerate(ix))
            return new, fac
        gens.append(flip)
    out = {tuple(index): value}
    todo = [tuple(index)]
    while todo:
        ix = todo.pop()
        for g in gens:
            new, f = g(ix)
            v = out[ix] * f
            if new not in out:
                out[new] = v
                todo.append(new)
    return out


def _spin_change(index, spaces, n_alpha):
    sc = 0.0
    for k, i in enumerate(index):
        beta = i >= n_alpha[k]
        if spaces[k][0] == "o":
            sc += 0.5 if beta else -0.5
        else:
            sc += -0.5 if beta else 0.5
    return int(sc)


def guess_singles(matrix, n_guesses, kind, degeneracy_tolerance=1e-14, max_diagonal_value=1000):
    """guesses_from_diagonal.py:200-271."""
    hf = matrix.hf
    sbs = kind_to_sym(kind)
    spin_change = -1 if kind == "spin_flip" else 0
    spaces = _axes(matrix)["ph"]
    n_alpha = [hf.n_alpha_of[s] for s in spaces]
    n_tot = [hf.n_orbs[s] for s in spaces]
    ax = [AxisBlocks(n_alpha[k], n_tot[k]) for k in range(2)]
    diag = matrix.diagonal()["ph"]
    elems = []
    for i in range(n_tot[0]):
        for a in range(n_tot[1]):
            idx = (i, a)
            if _spin_change(idx, spaces, n_alpha) != spin_change:
                continue
            if abs(diag[idx]) > max_diagonal_value:
                continue
            if sbs != "none" and i >= n_alpha[0]:
                continue   # select_n_min is unique-by-symmetry: canonical (alpha) block only
            si, sa = ax[0].split(i), ax[1].split(a)
            elems.append({"index": idx, "value": float(diag[idx]),
                          "spatial": ((si[1], sa[1]), (si[2], sa[2]))})
    elems.sort(key=lambda e: (e["value"], e["spatial"]))
    if not elems:
        return []
    # average degenerate values (find_smallest_matching_elements :175-187)
    istart = 0
    for i in range(len(elems)):
        if abs(elems[istart]["value"] - elems[i]["value"]) > degeneracy_tolerance:
            avg = np.average([e["value"] for e in elems[istart:i]])
            for e in elems[istart:i]:
                e["value"] = avg
            istart = i
    avg = np.average([e["value"] for e in elems[istart:]])
    for e in elems[istart:]:
        e["value"] = avg
    if len(elems) > n_guesses:
        cut = elems[n_guesses - 1]["value"]
        elems = [e for e in elems if e["value"] <= cut]

    shapes = matrix.shapes
    ret = []
    pspaces = _axes(matrix)["pphh"]

    def zero():
        v = {"ph": np.zeros(shapes["ph"])}
        if "pphh" in shapes:
            v["pphh"] = np.zeros(shapes["pphh"])
        return v

    def setel(vec, idx, val):
        for ix, x in orbit(idx, val, spaces, n_alpha, n_tot, sbs).items():
            vec["ph"][ix] = x

    for _, grp in itertools.groupby(elems, key=lambda e: (e["value"], e["spatial"])):
        if len(ret) >= n_guesses:
            break
        grp = list(grp)
        if len(grp) == 1:
            v = zero()
            setel(v, grp[0]["index"], 1.0)
            ret.append(v)
        elif len(grp) == 2:
            v = zero()
            setel(v, grp[0]["index"], 1.0)
            setel(v, grp[1]["index"], 1.0)
            ret.append(v)
            if len(ret) < n_guesses:
                v = zero()
                setel(v, grp[0]["index"], 1.0)
                setel(v, grp[1]["index"], -1.0)
                ret.append(v)
        else:
            raise AssertionError("group size > 2")
    for v in ret:
        nrm = np.sqrt(sum(float(np.vdot(t, t)) for t in v.values()))
        for k in v:
            v[k] = v[k] / nrm
    return ret


def guess_doubles(matrix, n_guesses, kind, degeneracy_tolerance=1e-14, max_diagonal_value=1000):
    """guesses_from_diagonal.py:274-305 + adc_guess_d.cc:205-543."""
    if n_guesses == 0:
        return []
    hf, mp = matrix.hf, matrix.mp
    sbs = kind_to_sym(kind)
    dm_s = -2 if kind == "spin_flip" else 0
    spaces = _axes(matrix)["pphh"]
    n_alpha = [hf.n_alpha_of[s] for s in spaces]
    n_tot = [hf.n_orbs[s] for s in spaces]
    ax = [AxisBlocks(n_alpha[k], n_tot[k]) for k in range(4)]
    sym_o = spaces[0] == spaces[1]
    sym_v = spaces[2] == spaces[3]
    spin = {"symmetric": 1, "antisymmetric": 3, "none": 0}[sbs]
    d1 = mp.df(spaces[0] + spaces[2])
    d2 = mp.df(spaces[1] + spaces[3])

    def smallest(d, ns):
        flat = np.argsort(d, axis=None, kind="stable")[:ns]
        return [(tuple(int(x) for x in np.unravel_index(f, d.shape)), float(d.flat[f])) for f in flat]

    forbidden = set()
    if spin in (1, 3):
        forbidden = {"aabb", "bbaa", "aaab", "aaba", "abaa", "baaa", "abbb", "babb", "bbab", "bbba"}

    def canon(spm, spi, idx):
        """index_group_d.cc:66-94"""
        spm, spi, idx = list(spm), list(spi), list(idx)
        for (p, q), on in (((0, 1), sym_o), ((2, 3), sym_v)):
            if not on:
                continue
            if (spi[p] == spi[q] and idx[p] > idx[q]) or spi[p] > spi[q]:
                for arr in (spm, spi, idx):
                    arr[p], arr[q] = arr[q], arr[p]
        return tuple(spm), tuple(spi), tuple(idx)

    ns = n_guesses
    groups = []
    max_reached = False
    size = 0
    while size < n_guesses and not max_reached:
        ns *= 2
        ilx, ily = smallest(d1, ns), smallest(d2, ns)
        max_reached = len(ilx) < ns
        groups = []   # list of [value, spi, idx, set(spin states)] in multimap (value) order
        for (ia, va), (ib, vb) in itertools.product(ilx, ily):
            if sym_o and ia[0] == ib[0]:
                continue
            if sym_v and ia[1] == ib[1]:
                continue
            full = [ia[0], ib[0], ia[1], ib[1]]
            # order by (block, in-block) == by index within an axis (adc_guess_d.cc:224-237)
            if sym_o and full[0] > full[1]:
                full[0], full[1] = full[1], full[0]
            if sym_v and full[2] > full[3]:
                full[2], full[3] = full[3], full[2]
            sp = [ax[k].split(full[k]) for k in range(4)]
            ms = 0
            for k in range(4):
                occupied = k < 2
                ms += 1 if occupied == sp[k][0] else -1
            if ms != dm_s:
                continue
            sblock = "".join("b" if s[0] else "a" for s in sp)
            if sblock in forbidden:
                continue
            # canonical representative under the alpha<->beta map: first spin alpha
            if spin in (1, 3) and sp[0][0]:
                sp = [(not s[0], s[1], s[2]) for s in sp]
            spm, spi, idx = canon([s[0] for s in sp], [s[1] for s in sp], [s[2] for s in sp])
            val = va + vb
            added = False
            for g in groups:
                if abs(val - g[0]) < degeneracy_tolerance and g[1] == spi and g[2] == idx:
                    g[3].add(spm)
                    added = True
                    break
            if not added:
                groups.append([val, spi, idx, {spm}])
        groups.sort(key=lambda g: g[0])   # stable: equal values keep insertion order
        size = 0
        for g in groups:
            idx = g[2]
            so = sym_o and idx[0] == idx[1]
            sv = sym_v and idx[2] == idx[3]
            if spin == 0:
                size += len(g[3])
            elif spin == 1:
                size += 1 if (so or sv) else 2
            else:
                if so and sv:
                    continue
                size += 1 if (so or sv) else 3

    def to_index(spm, spi, idx):
        out = []
        for k in range(4):
            b = spi[k] + (ax[k].n_alpha_blocks if spm[k] else 0)
            out.append(ax[k].starts[b] + idx[k])
        return tuple(out)

    A, B = False, True
    guesses = []
    for val, spi, idx, states in groups:
        if len(guesses) >= n_guesses:
            break
        so_sp = sym_o and spi[0] == spi[1] and idx[0] == idx[1]
        sv_sp = sym_v and spi[2] == spi[3] and idx[2] == idx[3]
        lv = []
        if spin == 0:
            st = sorted(states, key=lambda m: sum(int(x) << (3 - k) for k, x in enumerate(m)))
            if len(st) == 2:
                lv = [[(st[0], 1.0), (st[1], 1.0)], [(st[0], 1.0), (st[1], -1.0)]]
            elif len(st) == 6:
                coeff = [[2, 1, 1, 1, 1, 2], [0, 1, -1, -1, 1, 0], [1, 0, 0, 0, 0, -1],
                         [0, 1, 1, -1, -1, 0], [0, 1, -1, 1, -1, 0], [1, -1, -1, -1, -1, 1]]
                lv = [[(st[j], float(c[j])) for j in range(6)] for c in coeff]
            else:
                lv = [[(s, 1.0)] for s in st]
        elif spin == 1:
            if so_sp or sv_sp:
                one = [((A, B, A, B), 1.0)]
                if spi[0] != spi[1] and idx[0] != idx[1]:
                    one.append(((B, A, A, B), -1.0))
                elif spi[2] != spi[3] and idx[2] != idx[3]:
                    one.append(((A, B, B, A), -1.0))
                lv = [one]
            else:
                lv = [[((A, A, A, A), 2.0), ((A, B, A, B), 1.0), ((A, B, B, A), 1.0),
                       ((B, A, B, A), 1.0), ((B, A, A, B), 1.0)],
                      [((A, B, A, B), 1.0), ((A, B, B, A), -1.0), ((B, A, B, A), 1.0),
                       ((B, A, A, B), -1.0)]]
        else:
            if so_sp and sv_sp:
                continue
            if so_sp or sv_sp:
                one = [((A, B, A, B), 1.0)]
                if spi[0] != spi[1] and idx[0] != idx[1]:
                    one.append(((B, A, A, B), 1.0))
                elif spi[2] != spi[3] and idx[2] != idx[3]:
                    one.append(((A, B, B, A), 1.0))
                lv = [one]
            else:
                lv = [[((A, A, A, A), 1.0)],
                      [((A, B, A, B), 1.0), ((A, B, B, A), -1.0), ((B, A, B, A), -1.0),
                       ((B, A, A, B), 1.0)],
                      [((A, B, A, B), 1.0), ((A, B, B, A), 1.0), ((B, A, B, A), -1.0),
                       ((B, A, A, B), -1.0)]]
        for elems in lv:
            if len(guesses) >= n_guesses:
                break
            vec = {"ph": np.zeros(matrix.shapes["ph"]), "pphh": np.zeros(matrix.shapes["pphh"])}
            for spm, c in elems:
                for ix, x in orbit(to_index(spm, spi, idx), c, spaces, n_alpha, n_tot, sbs).items():
                    vec["pphh"][ix] = x
            nrm = float(np.vdot(vec["pphh"], vec["pphh"]))
            if nrm == 0.0:
                continue
            vec["pphh"] /= np.sqrt(nrm)
            guesses.append(vec)
    diag = matrix.diagonal()["pphh"]
    return [g for g in guesses
            if float(np.vdot(g["pphh"], diag * g["pphh"])) <= max_diagonal_value]


def obtain_guesses(matrix, n_guesses, kind, n_guesses_doubles=None):
    """adcc/workflow.py:460-499."""
    n_singles = n_guesses if n_guesses_doubles is None else n_guesses - n_guesses_doubles
    singles = guess_singles(matrix, n_singles, kind)
    doubles = []
    if "pphh" in matrix.shapes:
        if n_guesses_doubles is None:
            n_guesses_doubles = n_guesses - len(singles)
        if n_guesses_doubles > 0:
            doubles = guess_doubles(matrix, n_guesses_doubles, kind)
    total = singles + doubles
    if len(total) < n_guesses:
        raise ValueError(f"Less guesses found than requested: {len(total)} found, "
                         f"{n_guesses} requested")
    return total


def estimate_n_guesses(matrix, n_states, n_guesses_per_state=2):
    """adcc/workflow.py:425-457."""
    n = n_guesses_per_state * max(2, n_states)
    if "pphh" not in matrix.shapes:
        occ = "o2" if matrix.cvs else "o1"
        n = min(n, matrix.hf.n_alpha_of[occ] * matrix.hf.n_alpha_of["v1"])
    return max(n_states, n)
