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
