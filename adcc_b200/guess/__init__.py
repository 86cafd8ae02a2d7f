"""Guess vectors for the Davidson solver (adcc/guess/__init__.py:30-100)."""
from .guess_zero import guess_zero, guess_symmetries  # noqa: F401
from .guesses_from_diagonal import guesses_from_diagonal  # noqa: F401


def guess_kwargs_kind(kind):
    kwargsmap = dict(
        singlet=dict(spin_block_symmetrisation="symmetric", spin_change=0),
        triplet=dict(spin_block_symmetrisation="antisymmetric", spin_change=0),
        spin_flip=dict(spin_block_symmetrisation="none", spin_change=-1),
        any=dict(spin_block_symmetrisation="none", spin_change=0),
    )
    try:
        return kwargsmap[kind]
    except KeyError:
        raise ValueError(f"Kind not known: {kind}")


def guesses_singlet(matrix, n_guesses, block="ph", **kwargs):
    return guesses_from_diagonal(matrix, n_guesses, block=block, **guess_kwargs_kind("singlet"), **kwargs)


def guesses_triplet(matrix, n_guesses, block="ph", **kwargs):
    return guesses_from_diagonal(matrix, n_guesses, block=block, **guess_kwargs_kind("triplet"), **kwargs)


guesses_any = guesses_from_diagonal


def guesses_spin_flip(matrix, n_guesses, block="ph", **kwargs):
    return guesses_from_diagonal(matrix, n_guesses, block=block, **guess_kwargs_kind("spin_flip"), **kwargs)
