"""MO subspace bookkeeping (host side).

Mirrors the reference's ``adcc.MoSpaces`` / ``libadcc.MoSpaces`` for the C1 point group:
  * adcc/MoSpaces.py:51-156 (expansion of core_orbitals / frozen_core / frozen_virtual),
  * libadcc_src/MoSpaces.cc:189-312 (distribution into o1,o2,o3,v1,v2), :379-405 (internal
    order: spin a<b, then subspace o3<o2<o1<v1<v2, then orbital energy; every tensor axis is
    [alpha | beta]), :449-498 (block starts / spins per subspace),
  * libadcc_src/MoSpaces/construct_blocks.cc:25-76 (tiling into blocks of <= max_block_size).
Blocks are bookkeeping only here (guess ordering, spin metadata): the device store is dense.
"""
from collections.abc import Iterable
import numpy as np

MAX_BLOCK_SIZE = 16   # libadcc_src/AdcMemory.hh:43
_VALID = ("o1", "o2", "o3", "v1", "v2")
_LETTER = {"o": "o1", "v": "v1", "c": "o2"}


def split_spaces(spacestr):
    """'o1v1' -> ['o1','v1']; also understands the o/v/c shorthand of adcc/block.py:30-38."""
    if len(spacestr) % 2 == 0 and all(spacestr[i:i + 2] in _VALID for i in range(0, len(spacestr), 2)):
        return [spacestr[i:i + 2] for i in range(0, len(spacestr), 2)]
    try:
        return [_LETTER[ch] for ch in spacestr]
    except KeyError:
        raise ValueError(f"Invalid space string '{spacestr}'.")


def expand_spaceargs(n_orbs_alpha, **spaceargs):
    """adcc/MoSpaces.py:51-121: numbers / ranges / lists / (alpha, beta) pairs -> index lists."""
    noa = n_orbs_alpha

    def to_list(entry, from_min=None, from_max=None):
        if entry is None:
            return np.array([], dtype=int)
        if isinstance(entry, (int, np.integer)) and not isinstance(entry, bool):
            if from_min is not None:
                return from_min + np.arange(entry)
            return np.arange(from_max - entry, from_max)
        if isinstance(entry, Iterable):
            return np.array(list(entry), dtype=int)
        raise TypeError(f"Unsupported type {type(entry)} for an orbital space argument")

    out = {}
    for key in spaceargs:
        if isinstance(spaceargs[key], bool) and spaceargs[key]:
            raise NotImplementedError("Automatic determination of frozen-core electrons not implemented.")
        if not isinstance(spaceargs[key], tuple):
            spaceargs[key] = (spaceargs[key], spaceargs[key])
    n_done = [0, 0]
    for key in ("frozen_core", "core_orbitals"):
        val = spaceargs.get(key)
        if not val or (val[0] is None and val[1] is None):
            out[key] = []
            continue
        la = to_list(val[0], from_min=n_done[0])
        lb = noa + to_list(val[1], from_min=n_done[1])
        out[key] = np.concatenate((la, lb)).tolist()
        n_done[0] += len(la)
        n_done[1] += len(lb)
    val = spaceargs.get("frozen_virtual")
    if val and not (val[0] is None and val[1] is None):
        out["frozen_virtual"] = np.concatenate((to_list(val[0], from_max=noa),
                                                to_list(val[1], from_max=noa) + noa)).tolist()
    else:
        out["frozen_virtual"] = []
    return out


def construct_blocks(crude_starts, length, max_block_size=MAX_BLOCK_SIZE):
    """libadcc_src/MoSpaces/construct_blocks.cc:25-76."""
    crude = sorted(set(crude_starts)) or [0]
    edges = crude + [length]
    ret = []
    for lo, hi in zip(edges[:-1], edges[1:]):
        ln = hi - lo
        if ln <= 0:
            continue
        nb = (ln + max_block_size - 1) // max_block_size
        bs, nlarge = ln // nb, ln % nb
        pos = lo
        for ib in range(nb):
            ret.append(pos)
            pos += bs + 1 if ib < nlarge else bs
    return ret


class MoSpaces:
    def __init__(self, hfdata, core_orbitals=None, frozen_core=None, frozen_virtual=None,
                 max_block_size=MAX_BLOCK_SIZE, explicit_index_lists=False):
        """``explicit_index_lists``: the three space arguments are already lists of spin-orbital
        indices in the host provider's convention (the form libadcc.MoSpaces takes,
        libadcc_src/MoSpaces.hh:60-75), not the user-level numbers / ranges / pairs."""
        from .DataHfProvider import import_scf_results
        hf = import_scf_results(hfdata)
        noa = hf.n_orbs_alpha
        nf = 2 * noa
        if explicit_index_lists:
            args = {"core_orbitals": [int(i) for i in core_orbitals or []],
                    "frozen_core": [int(i) for i in frozen_core or []],
                    "frozen_virtual": [int(i) for i in frozen_virtual or []]}
        else:
            args = expand_spaceargs(noa, core_orbitals=core_orbitals, frozen_core=frozen_core,
                                    frozen_virtual=frozen_virtual)
        self.point_group = "C1"
        self.irreps = ["A"]
        self.irrep_totsym = "A"
        self.restricted = hf.restricted
        self.core_orbitals = args["core_orbitals"]
        self.frozen_core = args["frozen_core"]
        self.frozen_virtual = args["frozen_virtual"]
        occ = hf.occupation_f
        orben = hf.orben_f
        spin = ["a"] * noa + ["b"] * noa
        sub = ["f"] * nf

        def distribute(name, indices):
            na = nb = 0
            for i in indices:
                if i >= nf:
                    raise ValueError(f"Orbital index {i} overshoots range of valid indices.")
                if sub[i] != "f":
                    raise ValueError(f"MO index {i} is already distributed to subspace {sub[i]}; "
                                     "the orbital index lists must be disjoint.")
                sub[i] = name
                if spin[i] == "a":
                    na += 1
                else:
                    nb += 1
            if na != nb:
                raise ValueError("An unequal number of alpha and beta orbitals has been selected "
                                 f"for the orbital subspace {name}.")
        distribute("o2", self.core_orbitals)
        distribute("o3", self.frozen_core)
        distribute("v2", self.frozen_virtual)

        n_alpha = int(round(float(np.sum(occ[:noa]))))
        n_beta = int(round(float(np.sum(occ[noa:]))))
        done_a = sum(1 for i in range(noa) if sub[i] in ("o2", "o3"))
        if n_alpha <= done_a or n_beta <= done_a:
            raise ValueError("Too many orbitals selected to be core-occupied / frozen core: no "
                             "active electrons remain.")
        todo = {"a": n_alpha - done_a, "b": n_beta - done_a}
        order = sorted(range(nf), key=lambda i: (-occ[i], orben[i]))
        for i in order:
            if sub[i] != "f":
                continue
            if todo[spin[i]] > 0:
                sub[i] = "o1"
                todo[spin[i]] -= 1
            else:
                sub[i] = "v1"
        if todo["a"] > 0 or todo["b"] > 0 or not any(s == "v1" for s in sub):
            raise ValueError("Too many orbitals selected to be frozen virtuals.")

        sort_ss = ["o3", "o2", "o1", "v1", "v2"]
        full = sorted(range(nf), key=lambda i: (spin[i], sort_ss.index(sub[i]), orben[i]))
        self.map_index_hf_provider = {"f": list(full)}
        # absent subspaces count zero orbitals (libadcc_src/MoSpaces.cc:190-194, 265-268 initialise every
        # subspace entry; adcc/workflow.py:262-270 asks for n_orbs_alpha("o3") of a state without frozen core)
        self._n_alpha = {"f": noa, "o1": 0, "o2": 0, "o3": 0, "v1": 0, "v2": 0}
        self._n = {"f": nf, "o1": 0, "o2": 0, "o3": 0, "v1": 0, "v2": 0}
        self.subspaces_occupied = [s for s in ("o1", "o2", "o3") if any(x == s for x in sub)]
        self.subspaces_virtual = [s for s in ("v1", "v2") if any(x == s for x in sub)]
        self.subspaces = self.subspaces_occupied + self.subspaces_virtual
        for s in self.subspaces:
            idx = [i for i in full if sub[i] == s]
            self.map_index_hf_provider[s] = idx
            self._n_alpha[s] = sum(1 for i in idx if spin[i] == "a")
            self._n[s] = len(idx)
        self.map_block_start, self.map_block_spin, self.map_block_irrep = {}, {}, {}
        for s in ["f"] + self.subspaces:
            # a block starts wherever spin or subspace changes (libadcc_src/MoSpaces.cc:462-481)
            idx = self.map_index_hf_provider[s]
            crude = [k for k in range(len(idx))
                     if k == 0 or (spin[idx[k]], sub[idx[k]]) != (spin[idx[k - 1]], sub[idx[k - 1]])]
            starts = construct_blocks(crude, self._n[s], max_block_size)
            self.map_block_start[s] = starts
            self.map_block_spin[s] = ["a" if st < self._n_alpha[s] else "b" for st in starts]
            self.map_block_irrep[s] = ["A"] * len(starts)

    @property
    def occupied_orbitals(self):
        """Active valence-occupied spin orbitals in host-provider indexing (adcc/MoSpaces.py:173-179)."""
        return self.map_index_hf_provider["o1"]

    @property
    def virtual_orbitals(self):
        """Active virtual spin orbitals in host-provider indexing (adcc/MoSpaces.py:181-187)."""
        return self.map_index_hf_provider["v1"]

    @property
    def has_core_occupied_space(self):
        return "o2" in self.subspaces_occupied

    def has_subspace(self, space):
        return space in self.subspaces

    def n_orbs(self, space="f"):
        return self._n[space]

    def n_orbs_alpha(self, space="f"):
        return self._n_alpha[space]

    def n_orbs_beta(self, space="f"):
        return self._n[space] - self._n_alpha[space]

    def subspace_name(self, space):
        return {"o1": "active occupied orbitals", "o2": "core-occupied orbitals",
                "o3": "frozen occupied orbitals", "v1": "active virtual orbitals",
                "v2": "frozen virtual orbitals", "f": "molecular orbitals"}[space]

    def describe(self):
        return "\n".join(f"{s}: {self.n_orbs(s)} ({self.subspace_name(s)})" for s in self.subspaces)


class MoIndexTranslation:
    """libadcc_src/MoIndexTranslation.hh: index <-> (spin block, spatial block, in-block)."""

    def __init__(self, mospaces, space):
        self.mospaces = mospaces
        self.subspaces = split_spaces(space) if isinstance(space, str) else list(space)
        self.space = "".join(self.subspaces)
        self.ndim = len(self.subspaces)
        self.shape = tuple(mospaces.n_orbs(s) for s in self.subspaces)

    def _block_of(self, k, i):
        starts = self.mospaces.map_block_start[self.subspaces[k]]
        return max(b for b, s in enumerate(starts) if s <= i)

    def block_index_of(self, index):
        return tuple(self._block_of(k, i) for k, i in enumerate(index))

    def inblock_index_of(self, index):
        return tuple(i - self.mospaces.map_block_start[self.subspaces[k]][self._block_of(k, i)]
                     for k, i in enumerate(index))

    def spin_of(self, index):
        return "".join("a" if i < self.mospaces.n_orbs_alpha(self.subspaces[k]) else "b"
                       for k, i in enumerate(index))

    def block_index_spatial_of(self, index):
        out = []
        for k, i in enumerate(index):
            sp = self.subspaces[k]
            b = self._block_of(k, i)
            n_a_blocks = sum(1 for s in self.mospaces.map_block_spin[sp] if s == "a")
            out.append(b - n_a_blocks if self.mospaces.map_block_spin[sp][b] == "b" else b)
        return tuple(out)

    def split_spin(self, index):
        return self.spin_of(index), self.block_index_spatial_of(index), self.inblock_index_of(index)

    def split(self, index):
        """(block index, in-block index)  (libadcc_src/MoIndexTranslation.cc:218-251)."""
        return self.block_index_of(index), self.inblock_index_of(index)

    def combine(self, *args):
        """combine(block_index, inblock_index) or combine(spin_block, block_index_spatial, inblock_index):
        the inverse of split / split_spin  (MoIndexTranslation.cc:253-335)."""
        if len(args) == 3:
            spin_block, spatial, inblock = args
            if len(spin_block) != self.ndim or len(spatial) != self.ndim:
                raise ValueError(f"MoIndexTranslation is for subspace ({self.space}), which is of dimension "
                                 f"{self.ndim}, but passed spin block / spatial block index do not match.")
            block = []
            for k, (c, b) in enumerate(zip(spin_block, spatial)):
                if c not in "ab":
                    raise ValueError(f"Invalid spin-block identifier '{spin_block}'")
                n_a = sum(1 for s in self.mospaces.map_block_spin[self.subspaces[k]] if s == "a")
                block.append(int(b) + (n_a if c == "b" else 0))
            return self.combine(tuple(block), inblock)
        block, inblock = args
        if len(block) != self.ndim or len(inblock) != self.ndim:
            raise ValueError(f"MoIndexTranslation is for subspace ({self.space}), which is of dimension "
                             f"{self.ndim}, but passed block / in-block index do not match.")
        out = []
        for k, (b, i) in enumerate(zip(block, inblock)):
            starts = list(self.mospaces.map_block_start[self.subspaces[k]]) + [self.shape[k]]
            if b >= len(starts) - 1:
                raise ValueError(f"Block index {tuple(block)} overshoots the number of blocks at dimension {k}.")
            if starts[b] + i >= starts[b + 1]:
                raise ValueError(f"In-block index {tuple(inblock)} overshoots the block size at dimension {k}.")
            out.append(int(starts[b] + i))
        return tuple(out)

    def full_index_of(self, index):
        raise NotImplementedError("full_index_of not yet implemented.")      # MoIndexTranslation.cc:161-164

    def map_range_to_hf_provider(self, ranges):
        """[(subspace ranges, host ranges), ...]: the index box ``ranges`` (one (start, end) per axis) cut
        into the boxes that map contiguously onto the SCF provider's orbital order
        (MoIndexTranslation.cc:361-434)."""
        ranges = [tuple(int(x) for x in r) for r in ranges]
        if len(ranges) != self.ndim:
            raise ValueError(f"MoIndexTranslation is for subspace ({self.space}), which is of dimension "
                             f"{self.ndim}, but passed index range has a dimension of {len(ranges)}.")
        per_axis = []
        for k, (start, end) in enumerate(ranges):
            if start > end:
                raise ValueError(f"Passed index range at dimension {k}(== [{start},{end})) is invalid: "
                                 "start larger than end.")
            if end > self.shape[k]:
                raise ValueError(f"Passed index range at dimension {k}(== [{start},{end})) overshoots shape "
                                 f"{self.shape}.")
            hmap = self.mospaces.map_index_hf_provider[self.subspaces[k]]
            pieces = []
            for i in range(start, end):
                if pieces and hmap[i] == hmap[i - 1] + 1:
                    pieces[-1][1] = i + 1
                else:
                    pieces.append([i, i + 1])
            per_axis.append([((a, b), (hmap[a], hmap[a] + b - a)) for a, b in pieces])
        out = [((), ())]
        for pieces in per_axis:
            out = [(src + (s,), dst + (d,)) for src, dst in out for s, d in pieces]
        return [] if any(not p for p in per_axis) else out

    def hf_provider_index_of(self, index):
        return tuple(self.mospaces.map_index_hf_provider[self.subspaces[k]][i]
                     for k, i in enumerate(index))
