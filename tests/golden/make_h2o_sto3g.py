#!/usr/bin/env python3
"""Generate tests/golden/h2o_sto3g.npz from the reference's in-repo H2O/STO-3G RHF data.

Source: /root/reference/examples/water/data.py (+ import_data.py:8-35 for the meaning of
the arrays).  The reference stores the full 14^4 spin-orbital chemists' ERI; for a restricted
reference that is the 7^4 spatial tensor replicated over the (aa|aa),(aa|bb),(bb|aa),(bb|bb)
spin cases, so only the spatial part is kept (the script checks the replication is exact).
Run in the build container only (the reference tree does not exist on the GPU box).
"""
import importlib.util
import os
import numpy as np

REF = "/root/reference/examples/water/data.py"
spec = importlib.util.spec_from_file_location("refdata", REF)
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

na = nb = 7
nf = 2 * na
coeff = np.array(mod.coeff_data).reshape(nf, nb)
orben = np.array(mod.orben_data).reshape(nf)
eri = np.array(mod.eri_data).reshape(nf, nf, nf, nf)
dip = np.array(mod.dip_data).reshape(3, nb, nb)

a = slice(0, na)
b = slice(na, nf)
eri_sp = eri[a, a, a, a].copy()
# exact replication check
for s1 in (a, b):
    for s2 in (a, b):
        assert np.array_equal(eri[s1, s1, s2, s2], eri_sp)
full = np.zeros_like(eri)
for s1 in (a, b):
    for s2 in (a, b):
        full[s1, s1, s2, s2] = eri_sp
assert np.array_equal(full, eri)
assert np.array_equal(orben[:na], orben[na:])
assert np.array_equal(coeff[:na], coeff[na:])

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "h2o_sto3g.npz")
np.savez_compressed(
    out,
    n_orbs_alpha=na, n_bas=nb, n_occ_alpha=5,
    energy_scf=-7.4959319286025718e+01, conv_tol=1e-12,
    orben=orben[:na], orbcoeff=coeff[:na], eri_spatial_chem=eri_sp,
    dipole_ao=dip, nuclear_dipole=np.array([1.693194615993441, 0., 1.196196642772152]),
    nuclear_charge=10.0,
)
print("wrote", out, os.path.getsize(out), "bytes")
