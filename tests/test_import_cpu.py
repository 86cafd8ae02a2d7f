"""SCF import (MoSpaces + ReferenceState.fock/eri staging) on index-encoded data, CPU shim."""
import pytest

import cpu_shim
from import_common import CASES, check_imports


@pytest.fixture(autouse=True)
def shim(monkeypatch):
    cpu_shim.install(monkeypatch)


@pytest.mark.parametrize("case", sorted(CASES))
def test_import_blocks(case):
    check_imports(case)


@pytest.mark.parametrize("case", ["gen", "fc-fv-cvs", "fv-cvs"])
def test_import_blocks_from_spatial_eri(case):
    check_imports(case, spatial=True)


def test_invalid_spaces():
    import adcc_b200 as adcc
    from import_common import counter_data
    data, n_orb, n_occ = counter_data()
    with pytest.raises(ValueError):
        adcc.ReferenceState(data, core_orbitals=n_occ)            # no active electrons left
    with pytest.raises(ValueError):
        adcc.ReferenceState(data, frozen_virtual=n_orb - n_occ)   # no active virtuals left
    with pytest.raises(ValueError):
        adcc.ReferenceState(data, core_orbitals=([0], [0, 1 + n_orb]))   # unequal alpha / beta
    with pytest.raises(ValueError):
        adcc.ReferenceState(data, core_orbitals=[0, n_orb], frozen_core=[0, n_orb])   # overlapping lists
    ref = adcc.ReferenceState(data, core_orbitals=range(1, 2))                        # explicit range
    assert list(ref.mospaces.map_index_hf_provider["o2"]) == [1, 1 + n_orb]
    bad = dict(data)
    bad["fock_ff"] = data["fock_ff"][:3, :3]
    with pytest.raises(ValueError):
        adcc.ReferenceState(bad)


def _mospaces_goldens():
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "mospaces_goldens.json")) as fp:
        return json.load(fp)


@pytest.mark.parametrize("icase", range(10))
def test_mospaces_reference_expectations(icase):
    """The reference's own hand-written MoSpaces expectations (libadcc_src/tests/MoSpacesTest.cc)."""
    import numpy as np
    from adcc_b200.MoSpaces import MoSpaces
    gold = _mospaces_goldens()
    case = gold["cases"][icase]
    nf = len(gold["occupation_f"])
    data = {"restricted": True, "conv_tol": 1e-10, "orbcoeff_fb": np.zeros((nf, nf // 2)),
            "occupation_f": np.array(gold["occupation_f"], dtype=float),
            "orben_f": np.array(gold["orben_f"]), "fock_ff": np.diag(gold["orben_f"]),
            "eri_ffff_spatial": np.zeros((nf // 2,) * 4)}
    def split(lst):      # C++ layer: full [alpha | beta] indices; Python layer: (alpha, beta) lists
        return ([i for i in lst if i < nf // 2], [i - nf // 2 for i in lst if i >= nf // 2])
    mo = MoSpaces(data, core_orbitals=split(case["core_orbitals"]),
                  frozen_core=split(case["frozen_core"]),
                  frozen_virtual=split(case["frozen_virtual"]), max_block_size=case["block_size"])
    assert mo.point_group == "C1" and mo.irreps == ["A"]
    for key in ("subspaces_occupied", "subspaces_virtual", "subspaces"):
        assert getattr(mo, key) == case[key], key
    assert mo.has_core_occupied_space == bool(case["core_orbitals"])
    for space, n in case["n_orbs_alpha"].items():
        assert mo.n_orbs_alpha(space) == n and mo.n_orbs_beta(space) == n
        assert mo.n_orbs(space) == 2 * n
    for key in ("map_index_hf_provider", "map_block_start", "map_block_spin"):
        got = getattr(mo, key)
        assert set(got) == set(case[key]), key
        for space, expect in case[key].items():
            assert list(got[space]) == expect, (key, space)


def test_mo_index_translation_reference_expectations():
    """Scalar checks of libadcc_src/tests/MoIndexTranslationTest.cc (84 index translations)."""
    import numpy as np
    from adcc_b200.MoSpaces import MoSpaces, MoIndexTranslation
    gold = _mospaces_goldens()
    nf = len(gold["occupation_f"])
    data = {"restricted": True, "conv_tol": 1e-10, "orbcoeff_fb": np.zeros((nf, nf // 2)),
            "occupation_f": np.array(gold["occupation_f"], dtype=float),
            "orben_f": np.array(gold["orben_f"]), "fock_ff": np.diag(gold["orben_f"]),
            "eri_ffff_spatial": np.zeros((nf // 2,) * 4)}

    def split(lst):
        return ([i for i in lst if i < nf // 2], [i - nf // 2 for i in lst if i >= nf // 2])
    n_checks = n_ranges = 0
    for case in gold["motrans"]:
        mo = MoSpaces(data, core_orbitals=split(case["core_orbitals"]),
                      frozen_core=split(case["frozen_core"]),
                      frozen_virtual=split(case["frozen_virtual"]),
                      max_block_size=case["block_size"])
        tr = MoIndexTranslation(mo, case["space"])
        for fn, idx, expect in case["checks"]:
            got = getattr(tr, fn)(idx)
            got = got if isinstance(got, str) else list(got)
            assert got == expect, (case["space"], case["block_size"], fn, idx)
            n_checks += 1
            # split / combine invert each other (MoIndexTranslationTest.cc:68-77)
            assert tr.combine(*tr.split(idx)) == tuple(idx)
            assert tr.combine(tr.spin_of(idx), tr.block_index_spatial_of(idx), tr.inblock_index_of(idx)) == tuple(idx)
        for ranges, expect in case.get("range_checks", []):
            # contiguous boxes of an index range in the SCF provider's order (MoIndexTranslationTest.cc:83-119, 166-189)
            got = tr.map_range_to_hf_provider(ranges)
            assert [[[list(p) for p in src], [list(p) for p in dst]] for src, dst in got] == expect, \
                (case["space"], case["block_size"], ranges)
            n_ranges += 1
    assert n_checks == 84 and n_ranges == 12
