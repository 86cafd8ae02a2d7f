"""Exception types of the user-facing workflow (adcc/exceptions.py)."""


class InputError(ValueError):
    """The user supplied inconsistent or invalid input."""
