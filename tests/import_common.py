"""Index-encoded synthetic SCF data for import tests, in the spirit of adcc/HfCounterData.py:30-220
and adcc/tests/ReferenceState_counterdata_test.py: every Fock / ERI element encodes its own index
so that each subspace block imported by ReferenceState can be checked element by element, for
every subspace partitioning of adcc/tests/testcases.py:199-301 (gen, cvs, fc, fv, ...)."""
import numpy as np

CASES = {   # name: (core_orbitals, frozen_core, frozen_virtual)
    "gen": (None, None, None), "cvs": (1, None, None), "fc": (None, 1, None), "fv": (None, None, 1),
    "fc-cvs": (1, 1, None), "fv-cvs": (1, None, 2), "fc-fv": (None, 1, 1), "fc-fv-cvs": (1, 1, 1),
}


def counter_data(n_orb=7, n_occ=4, restricted=True):
    """Spatial chemists' ERI with (pq|rs) = code(sorted pairs): 8-fold symmetric by construction."""
    g = np.zeros((n_orb,) * 4)
    for p in range(n_orb):
        for q in range(n_orb):
            for r in range(n_orb):
                for s in range(n_orb):
                    a, b = max(p, q), min(p, q)
                    c, d = max(r, s), min(r, s)
                    if (a, b) < (c, d):
                        a, b, c, d = c, d, a, b
                    g[p, q, r, s] = 1e-3 * (((a * 10 + b) * 10 + c) * 10 + d + 1)
    orben = np.arange(1, n_orb + 1, dtype=float) - n_occ - 0.5
    nf = 2 * n_orb
    spin = np.array([0] * n_orb + [1] * n_orb)
    sp = np.arange(nf) % n_orb
    eri = g[np.ix_(sp, sp, sp, sp)]
    eri = eri * (spin[:, None, None, None] == spin[None, :, None, None]) \
        * (spin[None, None, :, None] == spin[None, None, None, :])
    fock = np.diag(np.concatenate((orben, orben)))
    # off-diagonal Fock elements inside one spin so that fock blocks are checked too
    for i in range(nf):
        for j in range(nf):
            if i != j and spin[i] == spin[j]:
                fock[i, j] = 1e-4 * (10 * max(sp[i], sp[j]) + min(sp[i], sp[j]) + 1)
    occ = np.array([1.0] * n_occ + [0.0] * (n_orb - n_occ))
    return {
        "restricted": restricted, "conv_tol": 1e-10, "energy_scf": -1.0, "spin_multiplicity": 1,
        "orbcoeff_fb": np.vstack([np.eye(n_orb)] * 2), "occupation_f": np.concatenate([occ, occ]),
        "orben_f": np.concatenate((orben, orben)), "fock_ff": fock, "eri_ffff": eri,
    }, n_orb, n_occ


def expected_spaces(n_orb, n_occ, core, fc, fv):
    """host indices of each subspace, [alpha | beta] (adcc/MoSpaces.py:51-121, MoSpaces.cc:379-405)."""
    core, fc, fv = core or 0, fc or 0, fv or 0
    out = {}
    for name, rng in (("o3", range(0, fc)), ("o2", range(fc, fc + core)), ("o1", range(fc + core, n_occ)),
                      ("v1", range(n_occ, n_orb - fv)), ("v2", range(n_orb - fv, n_orb))):
        idx = list(rng)
        if idx:
            out[name] = np.array(idx + [i + n_orb for i in idx])
    return out


def check_imports(case, spatial=False):
    """spatial=True hands over the restricted spatial ERI (key eri_ffff_spatial): the blocks are then
    cut and spin-expanded on the device (DataHfProvider._spatial_block_device)."""
    import adcc_b200 as adcc
    core, fc, fv = CASES[case]
    data, n_orb, n_occ = counter_data()
    if spatial:
        given = dict(data)
        given["eri_ffff_spatial"] = np.ascontiguousarray(data["eri_ffff"][:n_orb, :n_orb, :n_orb, :n_orb])
        del given["eri_ffff"]
        ref = adcc.ReferenceState(given, core_orbitals=core, frozen_core=fc, frozen_virtual=fv)
        assert hasattr(ref._hf, "eri_chem_block_device")
    else:
        ref = adcc.ReferenceState(data, core_orbitals=core, frozen_core=fc, frozen_virtual=fv)
    spaces = expected_spaces(n_orb, n_occ, core, fc, fv)
    ms = ref.mospaces
    assert sorted(ms.subspaces) == sorted(spaces)
    assert ms.has_core_occupied_space == ("o2" in spaces)
    for s, idx in spaces.items():
        assert ms.n_orbs(s) == len(idx) and ms.n_orbs_alpha(s) == len(idx) // 2
        assert list(ms.map_index_hf_provider[s]) == list(idx)
        np.testing.assert_array_equal(ref.orbital_energies(s).to_ndarray(), data["orben_f"][idx])
    chem = data["eri_ffff"]
    full_asym = chem.transpose(0, 2, 1, 3) - chem.transpose(0, 2, 3, 1)     # <pq||rs> = (pr|qs)-(ps|qr)
    names = sorted(spaces)
    for s1 in names:
        for s2 in names:
            blk = ref.fock(s1 + s2)
            np.testing.assert_array_equal(blk.to_ndarray(), data["fock_ff"][np.ix_(spaces[s1], spaces[s2])])
            assert not blk.mutable
    active = [s for s in names if s in ("o1", "o2", "v1")]
    n_checked = 0
    for s1 in active:
        for s2 in active:
            for s3 in active:
                for s4 in active:
                    if (n_checked % 3) and len(active) > 2:      # thin out: every third combination
                        n_checked += 1
                        continue
                    n_checked += 1
                    t = ref.eri(s1 + s2 + s3 + s4)
                    exp = full_asym[np.ix_(spaces[s1], spaces[s2], spaces[s3], spaces[s4])]
                    np.testing.assert_allclose(t.to_ndarray(), exp, rtol=0, atol=1e-15)
                    assert t.subspaces == [s1, s2, s3, s4]
    # cached + sugar
    assert ref.eri("o1o1v1v1") is ref.oovv
    assert "o1o1v1v1" in ref.cached_eri_blocks
    ref.cached_eri_blocks = []
    assert ref.cached_eri_blocks == []
    return ref
