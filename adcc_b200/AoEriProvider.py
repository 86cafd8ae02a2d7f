"""SCF data in the ATOMIC-ORBITAL basis with the AO -> MO integral transformation on the device
(the step before the path).

Counterpart of adcc/backends/EriBuilder.py:41-135 (per-block AO->MO transformation with the spin
pattern (aa|aa),(aa|bb),(bb|aa),(bb|bb), :117-119) and the host-program providers that feed
libadcc::ReferenceState::eri (libadcc_src/ReferenceState.cc:478-509).  The AO tensor (mu nu|la si)
is copied to HBM once; every requested MO block is four quarter transformations = four DMMA GEMMs
(smallest orbital space first), then for restricted references one spin-expansion kernel.  Nothing
of size v^4 ever passes through the host.
"""
import numpy as np

from . import backend as be
from .DataHfProvider import DataHfProvider
from .functions import einsum
from .Tensor import Tensor


class AoDataHfProvider(DataHfProvider):
    """dict provider whose ERIs are given as ``eri_bbbb`` (chemists' AO integrals, shape (nb,)*4).

    Other keys as in adcc/DataHfProvider.py:84-141 (``orbcoeff_fb``, ``occupation_f``, ``orben_f``,
    ``fock_ff``, ``restricted``, ``conv_tol``, ``multipoles`` ...)."""

    def __init__(self, data):
        if "eri_bbbb" not in data:
            raise KeyError("eri_bbbb (AO electron-repulsion integrals) is required")
        nb = np.asarray(data["orbcoeff_fb"]).shape[1]
        if tuple(np.shape(data["eri_bbbb"])) != (nb,) * 4:
            raise ValueError(f"Shape mismatch for key eri_bbbb: Expected {(nb,) * 4}, but got "
                             f"{np.shape(data['eri_bbbb'])}.")
        n = np.asarray(data["orbcoeff_fb"]).shape[0] // 2
        base = {k: v for k, v in data.items() if k != "eri_bbbb"}
        # DataHfProvider insists on one ERI key; hand it a zero-stride placeholder that is never read
        base["eri_ffff_spatial"] = np.broadcast_to(np.zeros(1), (n, n, n, n))
        super().__init__(base)
        del self.data["eri_ffff_spatial"]
        self.data["eri_bbbb"] = data["eri_bbbb"]
        self._ao = None
        self._coeff = np.asarray(data["orbcoeff_fb"], dtype=float)

    def has_eri_phys_asym_ffff_inner(self):
        return False

    def get_backend(self):
        return self.data.get("backend", "ao-dict")

    # ------------------------------------------------------------------ device transformation
    def _ao_tensor(self):
        if self._ao is None:
            self._ao = Tensor._wrap(None, ["b"] * 4, be.from_numpy(np.asarray(self.data["eri_bbbb"])))
            self._ao.set_immutable()
        return self._ao

    def _transform(self, rows):
        """(pq|rs) for four lists of coefficient rows; contracts the smallest space first."""
        cs = [Tensor._wrap(None, [f"x{k}", "b"], be.from_numpy(self._coeff[r])) for k, r in enumerate(rows)]
        letters = ["m", "n", "l", "s"]
        out_letters = ["P", "Q", "R", "S"]
        cur = self._ao_tensor()
        idx = list(letters)
        for k in sorted(range(4), key=lambda k: len(rows[k])):
            new = list(idx)
            new[k] = out_letters[k]
            cur = einsum("".join(idx) + "," + out_letters[k] + letters[k] + "->" + "".join(new), cur, cs[k])
            idx = new
        return cur._contig()

    def eri_chem_block_device(self, P, Q, R, S):
        na = self.n_orbs_alpha
        idx = [np.asarray(x, dtype=int) for x in (P, Q, R, S)]
        halves = []
        for x in idx:
            a, b = x[x < na], x[x >= na]
            halves.append((a, b))
        same_spatial = all(len(a) == len(b) and np.array_equal(a, b - na) for a, b in halves)
        if self.restricted and same_spatial and \
                np.array_equal(self._coeff[:na], self._coeff[na:]):
            spatial = self._transform([a for a, _ in halves])
            return be.spin_expand4(spatial)
        # general case: transform with spin-orbital coefficient rows, mask spin-forbidden elements
        full = self._transform(idx)
        spin = [(x >= na).astype(float) for x in idx]
        m1 = (spin[0][:, None] == spin[1][None, :]).astype(float).reshape(-1, 1)
        m2 = (spin[2][:, None] == spin[3][None, :]).astype(float).reshape(-1, 1)
        mask = be.gemm(be.from_numpy(m1), be.from_numpy(m2))
        return be.mul(full, mask.reshape(full.shape))

    def eri_chem_block(self, P, Q, R, S):
        return self.eri_chem_block_device(P, Q, R, S).cpu().numpy()

    def fill_eri_ffff(self, slices, out):
        idx = [np.arange(*s.indices(self.n_orbs)) for s in slices]
        out[:] = self.eri_chem_block(*idx)
