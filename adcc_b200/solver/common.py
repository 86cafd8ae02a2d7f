"""adcc/solver/common.py:26-46."""
import numpy as np


def select_eigenpairs(eigenvalues, n_ep, which):
    """Indices of the n_ep eigenpairs selected by `which` (eigenvalues sorted ascending)."""
    mask = np.zeros(len(eigenvalues), dtype=bool)
    if which == "LA":
        mask[-n_ep:] = True
    elif which == "SA":
        mask[:n_ep] = True
    elif which == "LM":
        mask[np.argsort(np.abs(eigenvalues))[-n_ep:]] = True
    elif which == "SM":
        mask[np.argsort(np.abs(eigenvalues))[:n_ep]] = True
    else:
        raise ValueError("For now only the values 'LM', 'LA', 'SM' and 'SA' are understood for 'which'.")
    return mask.nonzero()[0]
