"""Oracle reference state: MO subspaces, Fock and antisymmetrised ERI blocks.

Restates (test infrastructure, see oracle/__init__.py)
  * libadcc_src/MoSpaces.cc:189-312,379-405  (orbital -> subspace distribution, ordering
    spin < subspace (o3<o2<o1<v1<v2) < energy; each tensor axis is [alpha | beta]),
  * adcc/MoSpaces.py:51-121 (integer core_orbitals/frozen_core count from below,
    frozen_virtual from above, same number for alpha and beta),
  * libadcc_src/ReferenceState.cc:418-509 + import_eri.cc:28-266 (fock / <pq||rs> blocks),
  * adcc/ReferenceState.py:166-172 + adcc/block.py:30-38 (``hf.oovv``/``hf.fvv`` sugar:
    o->o1, v->v1, c->o2).
"""
import numpy as np

_LETTER = {"o": "o1", "v": "v1", "c": "o2"}


def split_space(s):
    if len(s) % 2 == 0 and all(s[i] in "ov" and s[i + 1] in "123" for i in range(0, len(s), 2)):
        return [s[i:i + 2] for i in range(0, len(s), 2)]
    return [_LETTER[ch] for ch in s]


class OracleRefState:
    def __init__(self, hf, core_orbitals=0, frozen_core=0, frozen_virtual=0):
        self.hf = hf
        na = hf.n_orbs_alpha
        self.restricted = hf.restricted
        self.conv_tol = hf.conv_tol
        self.energy_scf = hf.energy_scf
        n_alpha = int(round(hf.occupation_f[:na].sum()))
        n_beta = int(round(hf.occupation_f[na:].sum()))
        core_orbitals = core_orbitals or 0
        frozen_core = frozen_core or 0
        frozen_virtual = frozen_virtual or 0

        sub = {}
        for spin, off, nocc in ((0, 0, n_alpha), (1, na, n_beta)):
            # sort by (occupied first, energy) within the spin (MoSpaces.cc sorted_moidcs)
            idx = np.arange(na) + off
            order = sorted(idx, key=lambda i: (-hf.occupation_f[i], hf.orben_f[i]))
            occ, virt = order[:nocc], order[nocc:]
            # explicit index lists: lowest host indices first (adcc/MoSpaces.py:104-121)
            o3 = [off + k for k in range(frozen_core)]
            o2 = [off + frozen_core + k for k in range(core_orbitals)]
            v2 = [off + na - frozen_virtual + k for k in range(frozen_virtual)]
            o1 = [i for i in occ if i not in o3 and i not in o2]
            v1 = [i for i in virt if i not in v2]
            for name, lst in (("o3", o3), ("o2", o2), ("o1", o1), ("v1", v1), ("v2", v2)):
                lst = sorted(lst, key=lambda i: hf.orben_f[i])
                sub.setdefault(name, [[], []])[spin] = lst
        self.host_index = {k: np.array(v[0] + v[1], dtype=int) for k, v in sub.items()}
        self.n_alpha_of = {k: len(v[0]) for k, v in sub.items()}
        self.n_orbs = {k: len(v[0]) + len(v[1]) for k, v in sub.items()}
        self.has_core_occupied_space = self.n_orbs["o2"] > 0
        self._cache = {}

    # ---- helpers
    def spin_of_axis(self, space):
        """0 for alpha, 1 for beta along a subspace axis."""
        return np.array([0] * self.n_alpha_of[space] + [1] * (self.n_orbs[space] - self.n_alpha_of[space]))

    def fock(self, space):
        s = split_space(space)
        key = "f" + "".join(s)
        if key not in self._cache:
            self._cache[key] = self.hf.fock_ff[np.ix_(self.host_index[s[0]], self.host_index[s[1]])].copy()
        return self._cache[key]

    def eri(self, space):
        s = split_space(space)
        key = "".join(s)
        if key not in self._cache:
            idx = [self.host_index[x] for x in s]
            self._cache[key] = np.ascontiguousarray(self.hf.asym_block(*idx))
        return self._cache[key]

    def orben(self, space):
        return self.hf.orben_f[self.host_index[space]]

    def orbital_coefficients(self, space):
        """C[space index, basis function] (alpha and beta rows share the spatial AO basis)."""
        return self.hf.orbcoeff_fb[self.host_index[space]]

    def __getattr__(self, attr):
        if attr.startswith("_") or attr in ("hf",):
            raise AttributeError(attr)
        if attr.startswith("f") and len(attr) == 3:
            return self.fock(attr[1:])
        if len(attr) == 4 and all(c in "ovc" for c in attr):
            return self.eri(attr)
        raise AttributeError(attr)
