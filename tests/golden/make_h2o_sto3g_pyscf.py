#!/usr/bin/env python3
"""Generate tests/golden/h2o_sto3g_pyscf_hfdata.npz from the reference's second offline fixture,
/root/reference/examples/0_basics/h2o_sto3g_hfdata.hdf5 (PySCF RHF H2O/STO-3G in adcc's hfdata
layout, adcc/DataHfProvider.py:84-141), read with the repository's own HDF5 reader
(adcc_b200/hdf5io.py; h5py is not available offline).  Nested keys are flattened with '/'.
Run in the build container only (the reference tree does not exist on the GPU box)."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/examples/0_basics/h2o_sto3g_hfdata.hdf5"
spec = importlib.util.spec_from_file_location("hdf5io", os.path.join(HERE, "..", "..", "adcc_b200", "hdf5io.py"))
hdf5io = importlib.util.module_from_spec(spec)
spec.loader.exec_module(hdf5io)


def flatten(d, prefix=""):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(flatten(v, prefix + k + "/"))
        else:
            out[prefix + k] = np.asarray(v)
    return out


data = hdf5io.load(SRC)
flat = flatten(data)
out = os.path.join(HERE, "h2o_sto3g_pyscf_hfdata.npz")
np.savez_compressed(out, **flat)
print("wrote", out, os.path.getsize(out), "bytes;", sorted(flat))
