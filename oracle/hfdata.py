"""SCF input container of the oracle (test infrastructure, see oracle/__init__.py).

Follows the data contract of the reference's dict provider
(adcc/DataHfProvider.py:79-141): spin orbitals are ordered [all alpha | all beta]
(libadcc_src/HartreeFockSolution_i.hh:36-42), ``eri_ffff`` is in chemists' notation
(pq|rs), ``eri_phys_asym_ffff`` is <pq||rs>.  One extension key is understood:
``eri_ffff_spatial`` (n,n,n,n) = spatial chemists' ERI of a restricted reference, expanded
on the fly with the (aa|aa),(aa|bb),(bb|aa),(bb|bb) pattern of
adcc/backends/EriBuilder.py:117-119, so that large synthetic cases never materialise nf^4.
"""
import os
import numpy as np


class OracleHf:
    def __init__(self, data):
        self.data = data
        self.restricted = bool(data["restricted"])
        self.conv_tol = float(data.get("conv_tol", data.get("threshold", 1e-12)))
        self.orbcoeff_fb = np.asarray(data["orbcoeff_fb"], dtype=float)
        self.nf = self.orbcoeff_fb.shape[0]
        self.n_orbs_alpha = self.nf // 2
        self.n_bas = self.orbcoeff_fb.shape[1]
        self.occupation_f = np.asarray(data["occupation_f"], dtype=float)
        self.orben_f = np.asarray(data["orben_f"], dtype=float)
        self.fock_ff = np.asarray(data["fock_ff"], dtype=float)
        self.energy_scf = float(data.get("energy_scf", 0.0))
        self.spin_of = np.array([0] * self.n_orbs_alpha + [1] * self.n_orbs_alpha)
        mm = data.get("multipoles", {})
        self.dipole_ao = np.asarray(mm["elec_1"]) if "elec_1" in mm else None
        self.nuclear_dipole = np.asarray(mm.get("nuclear_1", np.zeros(3)))
        if "eri_ffff_spatial" in data:
            self._mode = "spatial"
            self._eri = np.asarray(data["eri_ffff_spatial"], dtype=float)
        elif "eri_phys_asym_ffff" in data:
            self._mode = "asym"
            self._eri = np.asarray(data["eri_phys_asym_ffff"], dtype=float)
        else:
            self._mode = "chem"
            self._eri = np.asarray(data["eri_ffff"], dtype=float)

    def chem_block(self, P, Q, R, S):
        """(pq|rs) for host spin-orbital index arrays P,Q,R,S."""
        P, Q, R, S = (np.asarray(x, dtype=int) for x in (P, Q, R, S))
        if self._mode == "chem":
            return self._eri[np.ix_(P, Q, R, S)]
        if self._mode == "spatial":
            na = self.n_orbs_alpha
            sp, sq, sr, ss = (self.spin_of[x] for x in (P, Q, R, S))
            blk = self._eri[np.ix_(P % na, Q % na, R % na, S % na)]
            m1 = (sp[:, None] == sq[None, :]).astype(float)
            m2 = (sr[:, None] == ss[None, :]).astype(float)
            return blk * m1[:, :, None, None] * m2[None, None, :, :]
        raise ValueError("chemists' ERI not available")

    def asym_block(self, P, Q, R, S):
        """<pq||rs> = (pr|qs) - (ps|qr)   (libadcc_src/import_eri.cc:62-68,197-240)."""
        if self._mode == "asym":
            return self._eri[np.ix_(P, Q, R, S)]
        a = self.chem_block(P, R, Q, S).transpose(0, 2, 1, 3)
        b = self.chem_block(P, S, Q, R).transpose(0, 2, 3, 1)
        return a - b


def restricted_dict_from_spatial(orben, n_occ, eri_spatial_chem, orbcoeff=None, dipole_ao=None,
                                 energy_scf=0.0, conv_tol=1e-12, nuclear_dipole=None):
    """Expand restricted spatial data to the dict layout of adcc/DataHfProvider.py:84-141
    (keeping the ERI spatial under the extension key ``eri_ffff_spatial``)."""
    orben = np.asarray(orben, dtype=float)
    n = orben.shape[0]
    if orbcoeff is None:
        orbcoeff = np.eye(n)
    occ = np.array([1.0] * n_occ + [0.0] * (n - n_occ))
    data = {
        "restricted": True,
        "conv_tol": conv_tol,
        "energy_scf": energy_scf,
        "spin_multiplicity": 1,
        "orbcoeff_fb": np.vstack([orbcoeff, orbcoeff]),
        "occupation_f": np.concatenate([occ, occ]),
        "orben_f": np.concatenate([orben, orben]),
        "fock_ff": np.diag(np.concatenate([orben, orben])),
        "eri_ffff_spatial": np.asarray(eri_spatial_chem, dtype=float),
    }
    mm = {}
    if dipole_ao is not None:
        mm["elec_1"] = np.asarray(dipole_ao)
    if nuclear_dipole is not None:
        mm["nuclear_1"] = np.asarray(nuclear_dipole)
    data["multipoles"] = mm
    return data


def load_h2o_sto3g(path=None):
    """The reference's offline H2O/STO-3G fixture (examples/water/data.py) as a dict."""
    if path is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests",
                            "golden", "h2o_sto3g.npz")
    z = np.load(path)
    return restricted_dict_from_spatial(
        z["orben"], int(z["n_occ_alpha"]), z["eri_spatial_chem"], orbcoeff=z["orbcoeff"],
        dipole_ao=z["dipole_ao"], energy_scf=float(z["energy_scf"]),
        conv_tol=float(z["conv_tol"]), nuclear_dipole=z["nuclear_dipole"])


def load_h2o_sto3g_pyscf(path=None):
    """The reference's second offline fixture (examples/0_basics/h2o_sto3g_hfdata.hdf5, PySCF RHF)
    as the nested hfdata dict adcc's DataHfProvider takes (tests/golden/make_h2o_sto3g_pyscf.py)."""
    if path is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests",
                            "golden", "h2o_sto3g_pyscf_hfdata.npz")
    z = np.load(path, allow_pickle=False)
    out = {}
    for key in z.files:
        val = z[key]
        val = val[()] if val.shape == () else val
        if isinstance(val, np.str_):
            val = str(val)
        node = out
        parts = key.split("/")
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = val
    return out
