"""Oracle reference state of the synthetic workload (test infrastructure, see oracle/__init__.py):
dense blocks from the input definition adcc_b200/synthetic.py::synthetic_eri_value (pure NumPy),
same orbital energies as adcc_b200.packed.synthetic_reference_state."""
import numpy as np

from adcc_b200.synthetic import synthetic_eri_block, synthetic_orbital_energies
from .refstate import OracleRefState, split_space


class OracleSyntheticRef(OracleRefState):
    def __init__(self, no, nv, seed=20261019, scale=0.02):
        self.no, self.nv, self.seed, self.scale = no, nv, seed, scale
        e_o, e_v = synthetic_orbital_energies(no, nv)
        self._eps = {"o1": np.concatenate((e_o[0::2], e_o[1::2])),
                     "v1": np.concatenate((e_v[0::2], e_v[1::2]))}
        self.restricted = False
        self.conv_tol = 1e-12
        self.energy_scf = 0.0
        self.n_orbs = {"o1": no, "v1": nv, "o2": 0, "o3": 0, "v2": 0}
        self.n_alpha_of = {"o1": no // 2, "v1": nv // 2, "o2": 0, "o3": 0, "v2": 0}
        self.has_core_occupied_space = False
        self._cache = {}

    def fock(self, space):
        s = split_space(space)
        if s[0] != s[1]:
            return np.zeros((self.n_orbs[s[0]], self.n_orbs[s[1]]))
        return np.diag(self._eps[s[0]])

    def eri(self, space):
        s = split_space(space)
        key = "".join("o" if x == "o1" else "v" for x in s)
        if key not in self._cache:
            self._cache[key] = synthetic_eri_block(self.seed, key, self.no, self.nv, self.scale)
        return self._cache[key]

    def orben(self, space):
        return self._eps[space]
