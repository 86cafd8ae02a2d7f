"""Tensor symmetry metadata (host side).

Mirrors libadcc_src/Symmetry.hh:89-172 / adcc/Symmetry.py:26-38: permutations are strings
like "-jiab" (sign + permuted index letters relative to the identity string), spin-block maps
are (from, to, factor) with spin strings like "aaaa", forbidden spin blocks a list of spin
strings.  The point group is C1 (libadcc_src/MoSpaces.cc:136), so irreps_allowed is ["A"].
"""
from .MoSpaces import split_spaces


class Symmetry:
    def __init__(self, mospaces, space, permutations=None, spin_block_maps=None,
                 spin_blocks_forbidden=None):
        self.mospaces = mospaces
        self.subspaces = split_spaces(space) if isinstance(space, str) else list(space)
        self.space = "".join(self.subspaces)
        self.ndim = len(self.subspaces)
        self.shape = tuple(mospaces.n_orbs(s) for s in self.subspaces)
        self.irreps_allowed = ["A"]
        self._perms = []          # list of (axis permutation tuple, sign)
        self._perm_strings = []
        self.spin_block_maps = list(spin_block_maps or [])
        self.spin_blocks_forbidden = list(spin_blocks_forbidden or [])
        if permutations:
            self.permutations = permutations

    @property
    def permutations(self):
        return list(self._perm_strings)

    @permutations.setter
    def permutations(self, perms):
        perms = list(perms)
        if not perms:
            self._perms, self._perm_strings = [], []
            return
        ident = perms[0].lstrip("+-")
        if len(ident) != self.ndim or len(set(ident)) != self.ndim:
            raise ValueError("First permutation must be the identity index string")
        parsed = []
        for p in perms:
            sign = -1.0 if p.startswith("-") else 1.0
            letters = p.lstrip("+-")
            if sorted(letters) != sorted(ident):
                raise ValueError(f"Invalid permutation string {p}")
            axes = tuple(ident.index(c) for c in letters)
            for k, ax in enumerate(axes):
                if self.subspaces[k] != self.subspaces[ax]:
                    raise ValueError(f"Permutation {p} mixes different subspaces")
            parsed.append((axes, sign))
        self._perms, self._perm_strings = parsed, perms

    def permutation_group(self):
        """Closure of the listed (axes, sign) permutations under composition."""
        group = {tuple(range(self.ndim)): 1.0}
        gens = [g for g in self._perms[1:]]
        todo = list(group.items())
        while todo:
            axes, sign = todo.pop()
            for gaxes, gsign in gens:
                new = tuple(axes[a] for a in gaxes)
                if new not in group:
                    group[new] = sign * gsign
                    todo.append((new, sign * gsign))
        return list(group.items())

    @property
    def has_permutations(self):
        return len(self._perms) > 1

    # libadcc.pyi:541-631 / libadcc_src/Symmetry.hh:109-172
    @property
    def has_irreps_allowed(self):
        return bool(getattr(self, "irreps_allowed", None))

    @property
    def has_spin_block_maps(self):
        return bool(self.spin_block_maps)

    @property
    def has_spin_blocks_forbidden(self):
        return bool(self.spin_blocks_forbidden)

    @property
    def empty(self):
        return not (self.has_permutations or self.spin_block_maps or self.spin_blocks_forbidden)

    def clear(self):
        self._perms, self._perm_strings = [], []
        self.spin_block_maps, self.spin_blocks_forbidden = [], []

    # ---- element-level semantics (used by guesses, __setitem__, select_n_*)
    def spin_string(self, index):
        return "".join("a" if i < self.mospaces.n_orbs_alpha(s) else "b"
                       for i, s in zip(index, self.subspaces))

    def orbit(self, index, value=1.0):
        """dict {equivalent index: value} under permutations and spin-block maps;
        returns None if the element is forced to zero."""
        index = tuple(int(i) for i in index)
        if self.spin_string(index) in self.spin_blocks_forbidden:
            return None
        na = [self.mospaces.n_orbs_alpha(s) for s in self.subspaces]
        out = {index: value}
        todo = [index]
        while todo:
            ix = todo.pop()
            images = []
            for axes, sign in self._perms[1:]:
                images.append((tuple(ix[a] for a in axes), sign))
            spin = self.spin_string(ix)
            for frm, to, fac in self.spin_block_maps:
                for src, dst in ((frm, to), (to, frm)):
                    if spin == src:
                        new = tuple(i - na[k] if src[k] == "b" and dst[k] == "a" else
                                    i + na[k] if src[k] == "a" and dst[k] == "b" else i
                                    for k, i in enumerate(ix))
                        images.append((new, fac))
            for new, f in images:
                v = out[ix] * f
                if new in out:
                    if abs(out[new] - v) > 1e-14 * max(1.0, abs(v)):
                        return None          # e.g. antisymmetric diagonal: must vanish
                else:
                    out[new] = v
                    todo.append(new)
        return out

    def is_allowed(self, index):
        return self.orbit(index) is not None

    def is_canonical(self, index):
        orb = self.orbit(index)
        return orb is not None and min(orb) == tuple(index)

    def describe(self):
        return (f"Symmetry({self.space}, permutations={self._perm_strings}, "
                f"spin_block_maps={self.spin_block_maps}, "
                f"spin_blocks_forbidden={self.spin_blocks_forbidden})")

    __repr__ = describe
