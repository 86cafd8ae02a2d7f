"""Iterative solvers that call the matvec: Jacobi-Davidson and Lanczos (adcc/solver/__init__.py)."""
from .davidson import jacobi_davidson, davidson, eigsh, DavidsonState  # noqa: F401
from .lanczos import lanczos, lanczos_iterations, LanczosIterator, LanczosState  # noqa: F401
from .preconditioner import JacobiPreconditioner, PreconditionerIdentity  # noqa: F401
from .explicit_symmetrisation import IndexSymmetrisation, IndexSpinSymmetrisation  # noqa: F401
from .SolverStateBase import EigenSolverStateBase  # noqa: F401
