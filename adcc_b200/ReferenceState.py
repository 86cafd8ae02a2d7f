"""Reference state: SCF data staged into HBM as dense spin-orbital Fock / <pq||rs> blocks.

Mirrors adcc/ReferenceState.py:39-172 (constructor arguments, ``hf.oovv`` / ``hf.fvv`` sugar,
``import_all`` for tiny systems) and libadcc_src/ReferenceState.cc:56-99 (orbital energies),
:418-476 (lazy cached ``fock(space)``), :478-509 (lazy cached ``eri(space)``, immutable),
:511-528 (``import_all``).  ERI staging follows libadcc_src/import_eri.cc:28-266: direct
import if the provider offers <pq||rs>, otherwise chemists' blocks are copied to the device
and antisymmetrised there, <pq||rs> = (pr|qs) - (ps|qr), by two strided-gather launches.
"""
import numpy as np

from . import backend as be
from .DataHfProvider import import_scf_results
from .MoSpaces import MoSpaces, split_spaces
from .Tensor import Tensor
from .timings import Timer


class ReferenceState:
    def __init__(self, hfdata, core_orbitals=None, frozen_core=None, frozen_virtual=None,
                 symmetry_check_on_import=False, import_all_below_n_orbs=10, mospaces=None):
        hf = import_scf_results(hfdata)
        self._hf = hf
        self.mospaces = mospaces if mospaces is not None else MoSpaces(
            hf, core_orbitals=core_orbitals, frozen_core=frozen_core, frozen_virtual=frozen_virtual)
        self.timer = Timer()
        self.symmetry_check_on_import = symmetry_check_on_import
        self.restricted = hf.restricted
        # 0 ("unknown") for unrestricted references, whatever the provider says (libadcc_src/ReferenceState.cc:333-339)
        self.spin_multiplicity = hf.spin_multiplicity if hf.restricted else 0
        self.conv_tol = hf.conv_tol
        self.energy_scf = hf.energy_scf
        self.backend = hf.backend
        self.irreducible_representation = "A"
        self.n_orbs_alpha = hf.n_orbs_alpha
        self.n_orbs_beta = hf.n_orbs_alpha
        self.n_orbs = hf.n_orbs
        occ = hf.occupation_f
        self.n_alpha = int(round(float(np.sum(occ[:hf.n_orbs_alpha]))))
        self.n_beta = int(round(float(np.sum(occ[hf.n_orbs_alpha:]))))
        self._orben_f = hf.orben_f
        self._fock_ff = hf.fock_ff
        self._orbcoeff_fb = hf.orbcoeff_fb
        self._fock = {}
        self._eri = {}
        self.environment = None
        self.operator_integral_provider = getattr(hf, "operator_integral_provider", None)
        if import_all_below_n_orbs is not None and hf.n_orbs < import_all_below_n_orbs:
            self.import_all()

    # ------------------------------------------------------------------ basic data
    @property
    def has_core_occupied_space(self):
        return self.mospaces.has_core_occupied_space

    @property
    def is_aufbau_occupation(self):
        """Lowest-energy orbitals occupied? (adcc/ReferenceState.py:184-194)"""
        e_homo = max(float(np.max(self.orbital_energies_host(s))) for s in self.mospaces.subspaces_occupied)
        e_lumo = min(float(np.min(self.orbital_energies_host(s))) for s in self.mospaces.subspaces_virtual)
        return e_homo < e_lumo

    @property
    def hf_provider(self):
        return self._hf

    def _host_index(self, space):
        return np.asarray(self.mospaces.map_index_hf_provider[space], dtype=int)

    def orbital_energies(self, space):
        sp, = split_spaces(space)
        return Tensor.from_ndarray(self.mospaces, [sp], self._orben_f[self._host_index(sp)])

    def orbital_energies_host(self, space):
        return self._orben_f[self._host_index(space)]

    def orbital_coefficients_host(self, space):
        """numpy (n_space, n_bas): rows follow the adcc order of `space`."""
        return self._orbcoeff_fb[self._host_index(space)]

    def _coefficients(self, space, keep):
        if not (isinstance(space, str) and space.endswith("b") and len(space) == 3):
            raise ValueError(f"Invalid space string {space}: an object orbital_coefficients({space}) "
                             "is not known (expected e.g. 'o1b').")
        sub = space[:2]
        idx = self._host_index(sub)
        coeff = np.array(self._orbcoeff_fb[idx])
        is_alpha = idx < self._hf.n_orbs_alpha
        if keep == "a":
            coeff[~is_alpha] = 0.0
        elif keep == "b":
            coeff[is_alpha] = 0.0
        return coeff

    def orbital_coefficients(self, space):
        """MO coefficients C[p, mu] of subspace ``space[:2]`` as a host array (n_space, n_bas)
        (libadcc_src/ReferenceState.cc:380-390; the basis axis "b" is not a tensor space here)."""
        return self._coefficients(space, "ab")

    def orbital_coefficients_alpha(self, space):
        """Like orbital_coefficients with the rows of beta orbitals zeroed (ReferenceState.cc:392-403)."""
        return self._coefficients(space, "a")

    def orbital_coefficients_beta(self, space):
        return self._coefficients(space, "b")

    # ------------------------------------------------------------------ Fock / ERI
    def fock(self, space):
        sp = split_spaces(space)
        if len(sp) != 2:
            raise ValueError("fock needs a two-index space string")
        key = "".join(sp)
        if key not in self._fock:
            with self.timer.record("import/fock/" + key):
                blk = self._fock_ff[np.ix_(self._host_index(sp[0]), self._host_index(sp[1]))]
                t = Tensor.from_ndarray(self.mospaces, sp, blk)
                t.set_immutable()
                self._fock[key] = t
        return self._fock[key]

    def eri(self, space):
        sp = split_spaces(space)
        if len(sp) != 4:
            raise ValueError("eri needs a four-index space string")
        key = "".join(sp)
        if key not in self._eri:
            with self.timer.record("import/eri/" + key):
                t = Tensor._wrap(self.mospaces, sp, self._stage_eri(sp))
                t.set_immutable()
                self._eri[key] = t
                self._hf.flush_cache()
        return self._eri[key]

    def _stage_eri(self, sp):
        hf = self._hf
        P, Q, R, S = (self._host_index(s) for s in sp)
        if hf.has_eri_phys_asym_ffff_inner():
            blk = hf.eri_asym_block(P, Q, R, S)
            if self.symmetry_check_on_import:
                self._check_eri_symmetry(sp, blk)
            return be.from_numpy(blk)
        # (pr|qs) goes to the device once; the exchange part re-uses it when an index pair
        # shares its subspace (import_eri.cc:123-144), otherwise (ps|qr) is staged as well
        on_device = hasattr(hf, "eri_chem_block_device")     # AO->MO transformed in HBM (8(f4))
        chem = hf.eri_chem_block_device if on_device else (lambda *ix: be.from_numpy(hf.eri_chem_block(*ix)))
        c1 = chem(P, R, Q, S)                                        # [p,r,q,s]
        out = be.permute(c1, (0, 2, 1, 3))                           # <pq|rs>
        if sp[2] == sp[3]:
            be.permute(c1, (0, 2, 3, 1), alpha=-1.0, out=out, beta=1.0)       # c1[p,s,q,r]
        elif sp[0] == sp[1]:
            be.permute(c1, (2, 0, 1, 3), alpha=-1.0, out=out, beta=1.0)       # c1[q,r,p,s]
        else:
            c2 = chem(P, S, Q, R)                                    # [p,s,q,r]
            be.permute(c2, (0, 2, 3, 1), alpha=-1.0, out=out, beta=1.0)
        if self.symmetry_check_on_import:
            self._check_eri_symmetry(sp, out.cpu().numpy())
        return out

    def _check_eri_symmetry(self, sp, blk):
        tol = 10 * self.conv_tol
        if sp[0] == sp[1] and np.max(np.abs(blk + blk.transpose(1, 0, 2, 3))) > tol:
            raise ValueError("imported ERI block is not antisymmetric in its bra indices")
        if sp[2] == sp[3] and np.max(np.abs(blk + blk.transpose(0, 1, 3, 2))) > tol:
            raise ValueError("imported ERI block is not antisymmetric in its ket indices")

    def import_all(self):
        ss = self.mospaces.subspaces
        for a in ss:
            for b in ss:
                self.fock(a + b)
        occ = self.mospaces.subspaces_occupied
        wanted = ["o1o1o1o1", "o1o1o1v1", "o1o1v1v1", "o1v1o1v1", "o1v1v1v1", "v1v1v1v1"]
        if "o2" in occ:
            wanted += ["o1o2o1o2", "o1o2o2v1", "o2v1o2v1", "o1o2v1v1", "o2o2v1v1", "o2o2o2v1",
                       "o1v1o2o2", "o1o1o2v1", "o2v1v1v1", "o1o2o1v1"]
        for w in wanted:
            self.eri(w)
        self.flush_hf_cache()

    @property
    def cached_eri_blocks(self):
        return sorted(self._eri)

    @cached_eri_blocks.setter
    def cached_eri_blocks(self, keep):
        self._eri = {k: v for k, v in self._eri.items() if k in set(keep)}

    @property
    def cached_fock_blocks(self):
        return sorted(self._fock)

    @cached_fock_blocks.setter
    def cached_fock_blocks(self, keep):
        self._fock = {k: v for k, v in self._fock.items() if k in set(keep)}

    def flush_hf_cache(self):
        self._hf.flush_cache()
        if hasattr(self._hf, "release_device_cache"):
            self._hf.release_device_cache()

    def __getattr__(self, attr):
        if attr.startswith("_"):
            raise AttributeError(attr)
        try:
            if attr.startswith("f") and len(attr) == 3:
                return self.fock(attr[1:])
            if len(attr) == 4:
                return self.eri(attr)
        except ValueError:
            pass
        raise AttributeError(attr)

    @property
    def nuclear_dipole(self):
        # (order, gauge_origin): libadcc_src/ReferenceState.cc:353-358; the dipole does not depend on the origin argument
        return np.asarray(self._hf.get_nuclear_multipole(1, "origin"))
