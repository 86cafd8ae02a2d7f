"""(Anti)symmetrisation against the reference's vendored golden arrays
(libadcc_src/tests/TensorTestData.cc, checks of libadcc_src/tests/TensorTest.cc:480-572)."""
import os
import numpy as np

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tensor_symm_goldens.npz"))


def check_tensor_goldens():
    from adcc_b200.Tensor import Tensor
    from adcc_b200 import backend as be
    a = Tensor._wrap(None, ["x"] * 4, be.from_numpy(GOLD["a"]))
    b = Tensor._wrap(None, ["x"] * 6, be.from_numpy(GOLD["b"]))
    cases = [(a.symmetrise((0, 1)), "a_sym_01"), (a.symmetrise([(0, 1), (2, 3)]), "a_sym_01_23"),
             (a.symmetrise((0, 1, 2)), "a_sym_012"), (b.symmetrise([(0, 1, 2), (3, 4, 5)]), "b_sym_012_345"),
             (a.antisymmetrise((0, 1)), "a_asym_01"), (a.antisymmetrise([(0, 1), (2, 3)]), "a_asym_01_23"),
             (a.antisymmetrise((0, 1, 2)), "a_asym_012"),
             (b.antisymmetrise([(0, 1, 2), (3, 4, 5)]), "b_asym_012_345")]
    for got, name in cases:
        np.testing.assert_allclose(got.to_ndarray(), GOLD[name], rtol=0, atol=1e-14, err_msg=name)
