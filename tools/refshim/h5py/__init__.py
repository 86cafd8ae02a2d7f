"""Stand-in for the absent third-party package h5py (TEST INFRASTRUCTURE): only the names adcc
touches at import time.  HDF5 I/O is outside the ADC matvec / Davidson path."""


class Group:
    pass


class Dataset:
    pass


class File:
    def __init__(self, *a, **k):
        raise ImportError("h5py is not installed; this is an import-time stand-in only")


def special_dtype(**kwargs):
    return object


def Empty(dtype):
    return None
