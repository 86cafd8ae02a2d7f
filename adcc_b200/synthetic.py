"""Synthetic SCF inputs (pure NumPy; inputs only, no arithmetic of the hot path).

``restricted_scf_dict`` builds a restricted closed-shell data dict with the key layout of
adcc/DataHfProvider.py:84-141 (ERI kept spatial under ``eri_ffff_spatial``) with the shapes of
BASELINE.json's molecular configs: canonical orbitals (diagonal Fock), a
random 8-fold-symmetric chemists' ERI scaled so that the ADC matrix stays diagonally dominant,
identity MO coefficients and random symmetric dipole integrals.
"""
import numpy as np

# (n_orb spatial, n_occ spatial, core orbitals for CVS, seed): shapes of SURVEY.md 8.0
SHAPES = {
    "h2o_ccpvdz": (24, 5, 0, 1001),      # C1: o=10, v=38 spin orbitals
    "h2o_ccpvtz": (58, 5, 0, 1002),      # C2: o=10, v=106
    "methyloxirane_ccpvdz": (86, 16, 0, 1003),   # C3: o=32, v=140
    "h2o_ccpvtz_cvs": (58, 5, 1, 1002),  # C4: c=2, o=8, v=106
    "ne_ccpvtz_cvs": (30, 5, 1, 1004),   # C4': c=2, o=8, v=50
}


def molecule_like_levels(lo, hi, n, dense_end="hi"):
    """n levels in [lo, hi] with square-root spacing: sparse near the frontier end, dense far from
    it, so that the lowest excitations are well separated like in a molecule (a uniform ladder
    gives hundreds of near-degenerate e_a - e_i and a Davidson that crawls)."""
    if n == 1:
        return np.array([lo if dense_end == "hi" else hi])
    t = np.sqrt(np.arange(n) / (n - 1.0))
    return lo + (hi - lo) * t if dense_end == "hi" else hi - (hi - lo) * t[::-1]


def restricted_scf_dict(n_orb, n_occ, seed=1001, scale=None, conv_tol=1e-12):
    """scale=None: 0.35 / (number of virtual spin orbitals), which keeps the spectral radius of the
    random <ab||cd> block (~0.85 * scale * n_virt_spin) a perturbation of the orbital-energy gaps."""
    rng = np.random.default_rng(seed)
    if scale is None:
        scale = 0.35 / (2 * (n_orb - n_occ))
    if n_occ > 1:
        occ_e = np.concatenate(([-20.0], molecule_like_levels(-1.6, -0.4, n_occ - 1, dense_end="lo")))
    else:
        occ_e = np.array([-0.5])
    virt_e = molecule_like_levels(0.15, 4.0, n_orb - n_occ, dense_end="hi")
    orben = np.concatenate((occ_e, virt_e))
    g = rng.standard_normal((n_orb,) * 4)
    # 8-fold symmetry of real chemists' integrals (pq|rs)
    g = g + g.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    g *= scale / 8.0
    dip = rng.standard_normal((3, n_orb, n_orb))
    dip = 0.5 * (dip + dip.transpose(0, 2, 1))
    occ = np.array([1.0] * n_occ + [0.0] * (n_orb - n_occ))
    eye = np.eye(n_orb)
    return {
        "restricted": True, "conv_tol": conv_tol, "energy_scf": 0.0, "spin_multiplicity": 1,
        "orbcoeff_fb": np.vstack([eye, eye]),
        "occupation_f": np.concatenate([occ, occ]),
        "orben_f": np.concatenate([orben, orben]),
        "fock_ff": np.diag(np.concatenate([orben, orben])),
        "eri_ffff_spatial": g,
        "multipoles": {"elec_1": dip, "nuclear_0": 2.0 * n_occ, "nuclear_1": np.zeros(3)},
        "backend": f"synthetic(n_orb={n_orb},n_occ={n_occ},seed={seed})",
    }


def named_case(name):
    n_orb, n_occ, core, seed = SHAPES[name]
    return restricted_scf_dict(n_orb, n_occ, seed), core


# ----------------------------------------------------------------------------------------------
# BASELINE config 5: counter-based synthetic antisymmetrised ERIs.
# This is the DEFINITION of the inputs; the CUDA generator (csrc/packed_ops.cu::synthetic_eri)
# implements the same integer hash, so the CPU oracle can evaluate any <pq||rs> entry.
# ----------------------------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
ERI_BLOCKS = {"oooo": 0, "ooov": 1, "oovv": 2, "ovov": 3, "ovvv": 4, "vvvv": 5}


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def synthetic_eri_value(seed, block, p, q, r, s, scale):
    """<pq||rs> of the synthetic workload for integer index arrays (broadcastable)."""
    blk = ERI_BLOCKS[block]
    p, q, r, s = (np.asarray(x, dtype=np.int64) for x in np.broadcast_arrays(p, q, r, s))
    p, q, r, s = p.copy(), q.copy(), r.copy(), s.copy()
    sign = np.ones(p.shape)
    zero = np.zeros(p.shape, dtype=bool)
    if block in ("oooo", "ooov", "oovv", "vvvv"):
        zero |= p == q
        sw = p > q
        p, q = np.where(sw, q, p), np.where(sw, p, q)
        sign = np.where(sw, -sign, sign)
    if block in ("oooo", "oovv", "ovvv", "vvvv"):
        zero |= r == s
        sw = r > s
        r, s = np.where(sw, s, r), np.where(sw, r, s)
        sign = np.where(sw, -sign, sign)
    if block in ("oooo", "vvvv", "ovov"):
        sw = (p > r) | ((p == r) & (q > s))
        p, r = np.where(sw, r, p), np.where(sw, p, r)
        q, s = np.where(sw, s, q), np.where(sw, q, s)
    with np.errstate(over="ignore"):
        key = ((p.astype(np.uint64) * np.uint64(1024) + q.astype(np.uint64)) * np.uint64(1024)
               + r.astype(np.uint64)) * np.uint64(1024) + s.astype(np.uint64)
        h = _splitmix64(np.uint64(seed) ^ _splitmix64(key * np.uint64(8) + np.uint64(blk)))
    u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    val = sign * (2.0 * u - 1.0) * 1.7320508075688772 * scale
    return np.where(zero, 0.0, val)


def synthetic_eri_block(seed, block, no, nv, scale):
    """Dense spin-orbital block (small sizes only; used by the oracle in parity tests)."""
    ext = [no if c == "o" else nv for c in block]
    idx = np.meshgrid(*[np.arange(e) for e in ext], indexing="ij")
    return synthetic_eri_value(seed, block, *idx, scale)


def synthetic_orbital_energies(no, nv):
    """Spin-orbital energies of the synthetic workload: occupied in [-2.0, -0.3], virtual in
    [0.2, 4.0] (the ranges of SURVEY.md 8(d)) with molecule-like square-root spacing."""
    return (molecule_like_levels(-2.0, -0.3, no, dense_end="lo"),
            molecule_like_levels(0.2, 4.0, nv, dense_end="hi"))
