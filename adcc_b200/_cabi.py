"""ctypes binding of include/adccb.h (the C ABI of libadccb.so).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no CPU
fallback: a missing library or a failing call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libadccb.so")

_lib = None

c_i64 = ctypes.c_int64
c_dbl = ctypes.c_double
c_ptr = ctypes.c_void_p


def _declare(lib):
    lib.adccb_version.restype = ctypes.c_int
    lib.adccb_last_error.restype = ctypes.c_char_p
    lib.adccb_launch_count.restype = ctypes.c_longlong
    lib.adccb_device_info.argtypes = [ctypes.POINTER(ctypes.c_int)] * 3
    lib.adccb_dgemm.argtypes = [c_ptr, c_i64, c_i64, c_i64, c_dbl, c_ptr, c_i64, c_i64, c_ptr,
                                c_i64, c_i64, c_dbl, c_ptr, c_i64, c_ptr, c_i64, ctypes.c_int]
    lib.adccb_dgemm_peers.argtypes = [c_ptr, c_i64, c_i64, c_i64, c_dbl, c_ptr, c_i64, c_ptr, c_i64,
                                      c_ptr, ctypes.c_int, ctypes.POINTER(c_ptr), c_i64, c_ptr, c_i64,
                                      ctypes.c_int]
    lib.adccb_peer_alloc.argtypes = [c_i64, ctypes.POINTER(c_ptr), c_ptr]
    lib.adccb_peer_open.argtypes = [c_ptr, ctypes.POINTER(c_ptr)]
    lib.adccb_peer_close.argtypes = [c_ptr]
    lib.adccb_peer_free.argtypes = [c_ptr]
    lib.adccb_strided_gather.argtypes = [c_ptr, ctypes.c_int, ctypes.POINTER(c_i64),
                                         ctypes.POINTER(c_i64), c_dbl, c_ptr, c_dbl, c_ptr, c_dbl,
                                         c_ptr]
    lib.adccb_elementwise.argtypes = [c_ptr, ctypes.c_int, c_i64, c_dbl, c_ptr, c_ptr, c_dbl,
                                      c_dbl, c_ptr]
    lib.adccb_direct_sum.argtypes = [c_ptr, c_i64, c_i64, c_dbl, c_ptr, c_dbl, c_ptr, c_ptr]
    lib.adccb_add_columns.argtypes = [c_ptr, c_i64, c_i64, c_ptr, c_i64, c_i64, c_ptr, c_i64]
    lib.adccb_lincomb.argtypes = [c_ptr, ctypes.c_int, ctypes.POINTER(c_dbl),
                                  ctypes.POINTER(c_ptr), c_i64, c_dbl, c_ptr]
    lib.adccb_lincomb_multi.argtypes = [c_ptr, ctypes.c_int, ctypes.c_int, ctypes.POINTER(c_dbl),
                                        ctypes.POINTER(c_ptr), c_i64, c_dbl, ctypes.POINTER(c_ptr)]
    lib.adccb_dot_many.argtypes = [c_ptr, ctypes.c_int, c_ptr, ctypes.POINTER(c_ptr), c_i64,
                                   c_dbl, c_dbl, c_ptr, c_ptr, c_i64]
    lib.adccb_spin_expand4.argtypes = [c_ptr, ctypes.POINTER(ctypes.c_int), c_ptr, c_ptr]
    lib.adccb_spin_project.argtypes = [c_ptr, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                       ctypes.POINTER(ctypes.c_int), c_dbl, c_ptr, c_ptr]
    lib.adccb_spin_singlet.argtypes = [c_ptr, ctypes.POINTER(ctypes.c_int),
                                       ctypes.POINTER(ctypes.c_int), c_ptr]
    c_int = ctypes.c_int
    lib.adccb_precond_apply.argtypes = [c_ptr, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_ptr, c_ptr,
                                        c_dbl, c_int, c_dbl, c_int, c_ptr]
    lib.adccb_packed_unpack.argtypes = [c_ptr, c_int, c_int, c_ptr, c_ptr]
    lib.adccb_packed_pack.argtypes = [c_ptr, c_int, c_int, c_ptr, c_ptr]
    lib.adccb_packed_expand.argtypes = [c_ptr, c_int, c_int, c_int, c_ptr, c_ptr]
    lib.adccb_packed_expand_slab.argtypes = [c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr]
    lib.adccb_slab_layout.argtypes = [c_ptr, c_int, c_i64, c_i64, c_int, ctypes.POINTER(c_i64), c_i64,
                                      c_ptr, c_ptr]
    lib.adccb_copy2d.argtypes = [c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_i64, c_i64]
    lib.adccb_block_gather.argtypes = [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]
    lib.adccb_block_scatter_add.argtypes = [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_i64]
    lib.adccb_block_scatter.argtypes = [c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_dbl, c_ptr, c_ptr, c_i64]
    lib.adccb_vvvv_exchange_slab.argtypes = [c_ptr, c_int, c_int, c_int, c_ptr, c_ptr]
    lib.adccb_ovvv_exchange_slab.argtypes = [c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr]
    lib.adccb_pair_matvec.argtypes = [c_ptr, c_int, c_int, c_int, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_int,
                                      c_dbl, c_dbl, c_ptr, c_i64, c_ptr, c_i64]
    lib.adccb_packed_fold_virt.argtypes = [c_ptr, c_i64, c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr,
                                           c_ptr, c_dbl, c_ptr]
    lib.adccb_packed_fold_occ.argtypes = [c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_dbl, c_ptr]
    lib.adccb_packed_diagonal.argtypes = [c_ptr, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]
    lib.adccb_pair_select.argtypes = [c_ptr, c_i64, c_int, c_i64, c_ptr, c_ptr]
    lib.adccb_synth_fill.argtypes = [c_ptr, c_int, c_int, c_int, c_int, ctypes.c_uint64, c_dbl,
                                     c_i64, c_i64, c_int, c_int, c_ptr]
    for name in EXPORTS:
        if name not in ("adccb_last_error", "adccb_launch_count"):
            getattr(lib, name).restype = ctypes.c_int


EXPORTS = [
    "adccb_version", "adccb_last_error", "adccb_launch_count", "adccb_device_info",
    "adccb_dgemm", "adccb_dgemm_peers", "adccb_peer_alloc", "adccb_peer_open", "adccb_peer_close",
    "adccb_peer_free", "adccb_strided_gather", "adccb_elementwise", "adccb_direct_sum",
    "adccb_add_columns", "adccb_lincomb", "adccb_lincomb_multi", "adccb_dot_many", "adccb_spin_expand4", "adccb_spin_project", "adccb_spin_singlet", "adccb_precond_apply",
    "adccb_packed_unpack", "adccb_packed_pack", "adccb_packed_expand", "adccb_packed_expand_slab",
    "adccb_slab_layout", "adccb_copy2d", "adccb_block_gather", "adccb_block_scatter_add", "adccb_block_scatter", "adccb_vvvv_exchange_slab", "adccb_ovvv_exchange_slab", "adccb_pair_matvec", "adccb_packed_fold_virt",
    "adccb_packed_fold_occ", "adccb_packed_diagonal", "adccb_pair_select", "adccb_synth_fill",
]


def lib():
    """Load libadccb.so (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA backend has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
                "adcc_b200 has no CPU fallback.")
        loaded = ctypes.CDLL(LIB_PATH)
        _declare(loaded)
        _lib = loaded
    return _lib


def check(status):
    if status != 0:
        msg = lib().adccb_last_error().decode()
        raise RuntimeError(f"adccb: {msg}")


def launch_count():
    return int(lib().adccb_launch_count())
