"""SCF input side: HartreeFockProvider interface and the dict provider.

Mirrors adcc/DataHfProvider.py:79-325 (required/recommended keys, shape checks, ``fill_*``
protocol) and the abstract interface libadcc_src/HartreeFockSolution_i.hh:36-252 (spin
orbitals ordered [alpha | beta]; ``eri_ffff`` chemists' (pq|rs); ``eri_phys_asym_ffff``
<pq||rs>).  Extension: key ``eri_ffff_spatial`` holds the spatial chemists' ERI of a
restricted reference and is expanded per requested block with the spin pattern of
adcc/backends/EriBuilder.py:117-119 (never materialising nf^4 on the host).
"""
import numpy as np


class HartreeFockProvider:
    """Interface a host program implements (same method names as adcc.HartreeFockProvider)."""

    # -- required
    def get_restricted(self): raise NotImplementedError
    def get_conv_tol(self): raise NotImplementedError
    def get_n_orbs_alpha(self): raise NotImplementedError
    def get_n_bas(self): raise NotImplementedError
    def fill_occupation_f(self, out): raise NotImplementedError
    def fill_orben_f(self, out): raise NotImplementedError
    def fill_orbcoeff_fb(self, out): raise NotImplementedError
    def fill_fock_ff(self, slices, out): raise NotImplementedError
    def fill_eri_ffff(self, slices, out): raise NotImplementedError

    # -- optional
    def fill_eri_phys_asym_ffff(self, slices, out): raise NotImplementedError
    def has_eri_phys_asym_ffff(self): return False        # the hook host back-ends override (libadcc.pyi:136)
    def has_eri_phys_asym_ffff_inner(self): return bool(self.has_eri_phys_asym_ffff())

    def transform_gauge_origin_to_xyz(self, gauge_origin):
        """libadcc.pyi:140; back-ends that know the molecule override it."""
        if isinstance(gauge_origin, str):
            if gauge_origin == "origin":
                return (0.0, 0.0, 0.0)
            raise NotImplementedError(f"gauge origin {gauge_origin!r} is not known to this provider")
        return tuple(float(x) for x in gauge_origin)

    def get_energy_scf(self): return 0.0
    def get_spin_multiplicity(self): return 1 if self.get_restricted() else 0
    def get_backend(self): return "generic"
    def flush_cache(self): pass

    # -- derived conveniences used by ReferenceState
    @property
    def restricted(self): return bool(self.get_restricted())
    @property
    def conv_tol(self): return float(self.get_conv_tol())
    @property
    def n_orbs_alpha(self): return int(self.get_n_orbs_alpha())
    @property
    def n_orbs(self): return 2 * self.n_orbs_alpha
    @property
    def n_bas(self): return int(self.get_n_bas())
    @property
    def energy_scf(self): return float(self.get_energy_scf())
    @property
    def spin_multiplicity(self): return int(self.get_spin_multiplicity())
    @property
    def backend(self): return self.get_backend()

    def _vec(self, fill):
        out = np.empty(self.n_orbs)
        fill(out)
        return out

    @property
    def occupation_f(self): return self._vec(self.fill_occupation_f)
    @property
    def orben_f(self): return self._vec(self.fill_orben_f)

    @property
    def orbcoeff_fb(self):
        out = np.empty((self.n_orbs, self.n_bas))
        self.fill_orbcoeff_fb(out)
        return out

    @property
    def fock_ff(self):
        n = self.n_orbs
        out = np.empty((n, n))
        self.fill_fock_ff((slice(0, n), slice(0, n)), out)
        return out

    def _block_via_slices(self, fill, idx):
        idx = [np.asarray(i, dtype=int) for i in idx]
        lo = [int(i.min()) for i in idx]
        hi = [int(i.max()) + 1 for i in idx]
        buf = np.empty(tuple(h - l for l, h in zip(lo, hi)))
        fill(tuple(slice(l, h) for l, h in zip(lo, hi)), buf)
        return buf[np.ix_(*[i - l for i, l in zip(idx, lo)])]

    def eri_chem_block(self, P, Q, R, S):
        """(pq|rs) for host index arrays."""
        return self._block_via_slices(self.fill_eri_ffff, (P, Q, R, S))

    def eri_asym_block(self, P, Q, R, S):
        """<pq||rs> = (pr|qs) - (ps|qr)  (libadcc_src/import_eri.cc:62-68, 197-240)."""
        if self.has_eri_phys_asym_ffff_inner():
            return self._block_via_slices(self.fill_eri_phys_asym_ffff, (P, Q, R, S))
        a = self.eri_chem_block(P, R, Q, S).transpose(0, 2, 1, 3)
        b = self.eri_chem_block(P, S, Q, R).transpose(0, 2, 3, 1)
        return a - b


class DataOperatorIntegralProvider:
    def __init__(self, backend="data"):
        self.backend = backend

    @property
    def available(self):
        return tuple(k for k in dir(self) if not k.startswith("_") and k not in ("backend", "available"))


class DataHfProvider(HartreeFockProvider):
    def __init__(self, data):
        if not isinstance(data, dict):
            raise TypeError("Can only deal with data objects of type dict.")
        self.data = data
        self._backend = data.get("backend", "dict")
        if np.asarray(data["orbcoeff_fb"]).shape[0] % 2 != 0:
            raise ValueError("orbcoeff_fb first axis should have even length")
        nb = self.get_n_bas()
        nf = 2 * self.get_n_orbs_alpha()
        na = nf // 2
        checks = [("orbcoeff_fb", (nf, nb)), ("occupation_f", (nf,)), ("orben_f", (nf,)),
                  ("fock_ff", (nf, nf)), ("eri_ffff", (nf, nf, nf, nf)),
                  ("eri_phys_asym_ffff", (nf, nf, nf, nf)), ("eri_ffff_spatial", (na, na, na, na))]
        for key, exshape in checks:
            if key in data and tuple(np.shape(data[key])) != exshape:
                raise ValueError(f"Shape mismatch for key {key}: Expected {exshape}, but got "
                                 f"{np.shape(data[key])}.")
        if not any(k in data for k in ("eri_ffff", "eri_phys_asym_ffff", "eri_ffff_spatial")):
            raise KeyError("one of eri_ffff, eri_phys_asym_ffff, eri_ffff_spatial is required")
        opprov = DataOperatorIntegralProvider(self._backend)
        mmp = data.get("multipoles", {})
        if "elec_1" in mmp:
            if np.shape(mmp["elec_1"]) != (3, nb, nb):
                raise ValueError(f"multipoles/elec_1 is expected to have shape {(3, nb, nb)} not "
                                 f"{np.shape(mmp['elec_1'])}")
            opprov.electric_dipole = np.asarray(mmp["elec_1"])
        self.operator_integral_provider = opprov
        self._spin = np.array([0] * na + [1] * na)
        self._spatial_dev = None
        if ("eri_ffff_spatial" in data and "eri_ffff" not in data
                and not hasattr(type(self), "eri_chem_block_device")):
            # restricted spatial ERI: blocks are cut and spin-expanded in HBM (ReferenceState
            # looks for this attribute), the host only uploads the spatial array once
            self.eri_chem_block_device = self._spatial_block_device

    def get_restricted(self): return bool(np.asarray(self.data["restricted"]))

    def get_conv_tol(self):
        key = "conv_tol" if "conv_tol" in self.data else "threshold"
        return float(np.asarray(self.data[key]))

    def fill_occupation_f(self, out): out[:] = self.data["occupation_f"]
    def fill_orbcoeff_fb(self, out): out[:] = self.data["orbcoeff_fb"]
    def fill_orben_f(self, out): out[:] = self.data["orben_f"]
    def fill_fock_ff(self, slices, out): out[:] = np.asarray(self.data["fock_ff"])[slices]

    def fill_eri_ffff(self, slices, out):
        if "eri_ffff" in self.data:
            out[:] = np.asarray(self.data["eri_ffff"])[slices]
        else:
            idx = [np.arange(*s.indices(self.n_orbs)) for s in slices]
            out[:] = self.eri_chem_block(*idx)

    def fill_eri_phys_asym_ffff(self, slices, out):
        out[:] = np.asarray(self.data["eri_phys_asym_ffff"])[slices]

    def has_eri_phys_asym_ffff_inner(self): return "eri_phys_asym_ffff" in self.data
    def get_backend(self): return self._backend
    def get_energy_scf(self): return float(np.asarray(self.data.get("energy_scf", 0.0)))

    def get_spin_multiplicity(self):
        if "spin_multiplicity" in self.data:
            return int(np.asarray(self.data["spin_multiplicity"]))
        if not self.get_restricted():
            return 0
        noa = self.get_n_orbs_alpha()
        occ = np.asarray(self.data["occupation_f"])
        return int(np.sum(occ[:noa])) - int(np.sum(occ[noa:])) + 1

    def get_n_orbs_alpha(self): return np.asarray(self.data["orbcoeff_fb"]).shape[0] // 2
    def get_n_bas(self): return np.asarray(self.data["orbcoeff_fb"]).shape[1]

    def get_nuclear_multipole(self, order, gauge_origin="origin"):
        mm = self.data.get("multipoles", {})
        key = f"nuclear_{order}"
        if key not in mm:
            raise NotImplementedError(f"Nuclear multipole with order {order} is not available.")
        return np.atleast_1d(np.asarray(mm[key], dtype=float))

    def eri_chem_block(self, P, Q, R, S):
        P, Q, R, S = (np.asarray(x, dtype=int) for x in (P, Q, R, S))
        if "eri_ffff" in self.data:
            return np.asarray(self.data["eri_ffff"])[np.ix_(P, Q, R, S)]
        if "eri_ffff_spatial" in self.data:
            na = self.n_orbs_alpha
            sp = np.asarray(self.data["eri_ffff_spatial"])
            blk = sp[np.ix_(P % na, Q % na, R % na, S % na)]
            m1 = (self._spin[P][:, None] == self._spin[Q][None, :]).astype(float)
            m2 = (self._spin[R][:, None] == self._spin[S][None, :]).astype(float)
            return blk * m1[:, :, None, None] * m2[None, None, :, :]
        raise ValueError("chemists' ERI not available from this data")

    def _spatial_block_device(self, P, Q, R, S):
        """(pq|rs) for spin-orbital index lists of the form [alpha subset | the same beta subset]:
        the spatial block is cut out of the device copy of ``eri_ffff_spatial`` and expanded with
        the spin pattern of adcc/backends/EriBuilder.py:117-119 by one kernel."""
        from . import backend as be
        na = self.n_orbs_alpha
        spatial_idx = []
        for x in (np.asarray(v, dtype=int) for v in (P, Q, R, S)):
            a, b = x[x < na], x[x >= na]
            if len(a) != len(b) or not np.array_equal(a, b - na) or not np.array_equal(x, np.concatenate((a, b))):
                return be.from_numpy(self.eri_chem_block(P, Q, R, S))       # ragged: host path
            spatial_idx.append(a)
        if all(len(a) > 0 and np.array_equal(a, np.arange(a[0], a[0] + len(a))) for a in spatial_idx):
            if self._spatial_dev is None:
                self._spatial_dev = be.from_numpy(np.asarray(self.data["eri_ffff_spatial"]))
            view = self._spatial_dev
            for ax, a in enumerate(spatial_idx):
                view = view.narrow(ax, int(a[0]), len(a))
            block = be.contiguous(view)
        else:
            block = be.from_numpy(np.asarray(self.data["eri_ffff_spatial"])[np.ix_(*spatial_idx)])
        return be.spin_expand4(block)

    def release_device_cache(self):
        self._spatial_dev = None

    def eri_asym_block(self, P, Q, R, S):
        if "eri_phys_asym_ffff" in self.data:
            return np.asarray(self.data["eri_phys_asym_ffff"])[np.ix_(P, Q, R, S)]
        return super().eri_asym_block(P, Q, R, S)


def import_scf_results(res):
    """adcc/backends/__init__.py:58-111 restricted to what exists offline: dict or provider."""
    if isinstance(res, HartreeFockProvider):
        return res
    if isinstance(res, dict):
        if "eri_bbbb" in res:
            from .AoEriProvider import AoDataHfProvider
            return AoDataHfProvider(res)
        return DataHfProvider(res)
    if isinstance(res, str):
        # on-disk hfdata dictionary (adcc/backends/__init__.py:100-106; read without h5py)
        import os
        if os.path.isfile(res) and (res.endswith(".h5") or res.endswith(".hdf5")):
            from .hdf5io import load
            return import_scf_results(load(res))
        raise ValueError(f"Unrecognised path or file extension: {res}")
    raise NotImplementedError("No means to import an SCF result of type " + str(type(res)) +
                              " (supported here: dict, HartreeFockProvider).")
