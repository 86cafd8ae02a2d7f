"""Device-side primitives: thin Python wrappers around the C ABI (include/adccb.h).

torch is used for device memory and streams only; every arithmetic operation below is one of
the hand-written sm_100a kernels in adcc_b200/csrc.  No CPU fallback exists.
"""
import ctypes
import numpy as np
import torch

from . import _cabi

F64 = torch.float64
EW_AXPBY, EW_MUL, EW_DIV, EW_FILL, EW_DIV_SHIFT, EW_ADD_SCALAR = range(6)

_ws = {}


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("adcc_b200 needs a CUDA device (B200, sm_100a); no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    if _raw_stream is not None:        # same handle as current_stream().cuda_stream, ~10x cheaper
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def workspace(nbytes=None):
    """Per-device scratch (split-K partials, reduction partials)."""
    dev = torch.cuda.current_device()
    want = max(nbytes or 0, 64 << 20)
    ws = _ws.get(dev)
    if ws is None or ws.numel() * 8 < want:
        ws = torch.empty(want // 8, dtype=F64, device=device())
        _ws[dev] = ws
    return ws


def empty(shape):
    return torch.empty(tuple(int(s) for s in shape), dtype=F64, device=device())


def zeros(shape):
    t = empty(shape)
    fill(t, 0.0)
    return t


def from_numpy(arr):
    arr = np.ascontiguousarray(arr, dtype=np.float64)
    return torch.from_numpy(arr).to(device())


def upload(host_tensor):
    """Host (ideally pinned) FP64 torch tensor -> device copy on the current stream."""
    return host_tensor.to(device(), non_blocking=True)


def download(dev_tensor, host_tensor):
    """Device array -> existing host tensor; returns after the copy has finished."""
    host_tensor.copy_(dev_tensor, non_blocking=True)
    if dev_tensor.is_cuda:
        torch.cuda.current_stream().synchronize()
    return host_tensor


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _check_dev(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda or t.dtype != F64):
            raise ValueError("expected CUDA float64 tensors")


# --------------------------------------------------------------------------------- graphs
def graphs_supported():
    return torch.cuda.is_available()


class ReplayGraph:
    """A fixed launch sequence captured once into a CUDA graph (stream capture) and replayed:
    ``fn(inputs) -> outputs`` over static input buffers.  For small systems the matvec is a few
    dozen microsecond-sized kernels and the host cost of issuing them dominates; one graph launch
    replaces them.  ``fn`` must not synchronise or read device data on the host."""

    def __init__(self, shapes, fill, fn):
        self.inputs = {k: empty(shape) for k, shape in shapes.items()}
        fill(self.inputs)
        workspace()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            fn(self.inputs)            # eager pass: operand layout caches fill outside the capture
        cur.wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn(self.inputs)

    def run(self, fill):
        fill(self.inputs)
        self.graph.replay()
        return self.outputs


def release_cached_memory():
    """Return the allocator's cached, unused device blocks to the driver (after large builds)."""
    if torch.cuda.is_available():
        torch.cuda.empty_cache()


def free_memory():
    """Bytes the device can still hand out: free device memory plus the allocator's cached blocks."""
    free, _ = torch.cuda.mem_get_info()
    return free + torch.cuda.memory_reserved() - torch.cuda.memory_allocated()


def clone(t):
    """Device-to-device copy of a contiguous array (cudaMemcpyAsync on the current stream)."""
    return t.clone()


# --------------------------------------------------------------------------------- GEMM
#: largest operand (elements) that gemm() re-pitches on the fly; bigger constant operands with an
#: odd row pitch should be stored through even_pitch() once
REPITCH_MAX = 1 << 27


def even_pitch(t):
    """``t`` if its rows are 16-byte multiples apart, else a copy with one spare column per row
    (returned as the [:, :n] view).  The TMA loads of adccb_dgemm need such rows; operands with an
    odd pitch (e.g. 5565 = pairs of 106 virtuals) would otherwise take the plain-load fallback
    kernel, measured 5-10x slower."""
    if t.dim() != 2 or t.stride(1) != 1 or t.shape[0] <= 1 or t.stride(0) % 2 == 0:
        return t
    buf = empty((t.shape[0], t.shape[1] + 1))
    view = buf[:, :t.shape[1]]
    copy2d(view, t, t.shape[0], t.shape[1])
    return view


def _tma_operand(t, is_b):
    if t.numel() > REPITCH_MAX or t.numel() == 0:
        return t
    if t.stride(1) == 1:
        return even_pitch(t)
    if is_b and t.stride(0) == 1 and t.stride(1) % 2 == 1 and t.shape[1] > 1:     # N-major B
        return even_pitch(t.t()).t()
    return t


def gemm(A, B, out=None, alpha=1.0, beta=0.0, tile_hint=0):
    """out[m,n] = alpha * sum_k A[m,k] B[n,k] + beta*out.  A, B: 2-D views (any strides)."""
    _check_dev(A, B, out)
    M, K = A.shape
    N, K2 = B.shape
    if K != K2:
        raise ValueError("gemm: inner extents differ")
    if M * N * K >= 32 ** 3:
        A, B = _tma_operand(A, False), _tma_operand(B, True)
    if out is None:
        out = empty((M, N))
        beta = 0.0
    if out.stride(1) != 1 and out.numel() > 0:
        raise ValueError("gemm: output must be row-major")
    ldc = out.stride(0) if M > 1 else max(N, 1)
    ws = workspace()
    _cabi.check(_cabi.lib().adccb_dgemm(
        _stream(), M, N, K, float(alpha), _ptr(A), A.stride(0), A.stride(1), _ptr(B),
        B.stride(0), B.stride(1), float(beta), _ptr(out), ldc, _ptr(ws), ws.numel() * 8,
        int(tile_hint)))
    return out


# --------------------------------------------------------------------------------- peer exchange
class _RawDeviceArray:
    """__cuda_array_interface__ view of memory this library allocated (wrapped as a torch tensor)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def peer_exchange_supported():
    return torch.cuda.is_available()


class PeerBuffer:
    """One FP64 buffer of ``n`` elements per rank, each mapped into every other rank's address space
    (CUDA IPC; one process per GPU on one NVSwitch box).  ``local`` is this rank's buffer as a
    torch tensor, ``peer_ptrs[r]`` the device address of rank r's buffer as seen from here.
    Construction is collective; ``PeerBuffer.create`` returns None on EVERY rank if any rank could
    not allocate, export or map (the caller then keeps the NCCL exchange)."""

    def __init__(self):
        self.n, self._own, self._opened, self.peer_ptrs, self.local = 0, None, [], [], None
        self.error = None

    @classmethod
    def create(cls, n, group=None):
        import torch.distributed as dist
        self = cls()
        self.n = int(n)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        handle = ctypes.create_string_buffer(64)
        try:
            ptr = ctypes.c_void_p()
            _cabi.check(_cabi.lib().adccb_peer_alloc(self.n * 8, ctypes.byref(ptr), handle))
            self._own = ptr.value
            self.local = torch.as_tensor(_RawDeviceArray(self._own, self.n), device=device())
        except Exception as exc:      # noqa: BLE001 - reported, decided collectively below
            self.error = f"rank {self.rank}: {exc}"
        handles = [None] * self.world
        dist.all_gather_object(handles, (self.error, handle.raw), group=group)
        if all(err is None for err, _ in handles):
            try:
                for r, (_, h) in enumerate(handles):
                    if r == self.rank:
                        self.peer_ptrs.append(self._own)
                        continue
                    p = ctypes.c_void_p()
                    _cabi.check(_cabi.lib().adccb_peer_open(ctypes.create_string_buffer(h, 64),
                                                            ctypes.byref(p)))
                    self.peer_ptrs.append(p.value)
                    self._opened.append(p.value)
            except Exception as exc:  # noqa: BLE001
                self.error = f"rank {self.rank}: {exc}"
        errors = [None] * self.world
        dist.all_gather_object(errors, self.error or next((e for e, _ in handles if e), None),
                               group=group)
        failed = [e for e in errors if e]
        if failed:
            self.close()
            if self.rank == 0:
                import sys
                print(f"adcc_b200: peer-memory exchange unavailable ({failed[0]}); using the NCCL "
                      "all-gather", file=sys.stderr)
            return None
        return self

    def close(self):
        for p in self._opened:
            _cabi.lib().adccb_peer_close(ctypes.c_void_p(p))
        self._opened, self.peer_ptrs = [], []
        if self._own:
            self.local = None
            _cabi.lib().adccb_peer_free(ctypes.c_void_p(self._own))
            self._own = None


def gemm_peers(A, B, local_out, peer_ptrs, col_offset, ldc, tile_hint=0):
    """local_out[:, :] = A B^T, also written to address peer_ptrs[d] + col_offset elements (same
    leading dimension) for every d: the column-slab GEMM whose epilogue does the slab exchange
    (col_offset = first_row * ldc + first_column of the slab inside the peers' buffers)."""
    _check_dev(A, B, local_out)
    M, K = A.shape
    N, K2 = B.shape
    if K != K2 or tuple(local_out.shape) != (M, N) or local_out.stride(1) != 1 or local_out.stride(0) != ldc:
        raise ValueError("gemm_peers: shape mismatch")
    if A.stride(1) != 1 or B.stride(1) != 1:
        raise ValueError("gemm_peers: operands must be K-contiguous")
    k = len(peer_ptrs)
    ps = (ctypes.c_void_p * max(k, 1))(*[int(p) + 8 * int(col_offset) for p in peer_ptrs])
    ws = workspace()
    _cabi.check(_cabi.lib().adccb_dgemm_peers(
        _stream(), M, N, K, 1.0, _ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(local_out), k, ps,
        int(ldc), _ptr(ws), ws.numel() * 8, int(tile_hint)))
    return local_out


# --------------------------------------------------------------------------------- gather
def gather(inp, shape, in_strides, alpha=1.0, add=None, gamma=0.0, out=None, beta=0.0):
    """out[o] = alpha*inp.flat[sum o_d*in_strides[d]] + gamma*add[o] + beta*out[o]."""
    _check_dev(inp, add, out)
    shape = [int(s) for s in shape]
    if out is None:
        out = empty(shape)
        beta = 0.0
    elif not out.is_contiguous():
        raise ValueError("gather: out must be contiguous")
    if add is not None and not add.is_contiguous():
        raise ValueError("gather: add must be contiguous")
    nd = len(shape)
    sh = (ctypes.c_int64 * max(nd, 1))(*shape)
    st = (ctypes.c_int64 * max(nd, 1))(*[int(s) for s in in_strides])
    _cabi.check(_cabi.lib().adccb_strided_gather(
        _stream(), nd, sh, st, float(alpha), _ptr(inp), float(gamma), _ptr(add), float(beta),
        _ptr(out)))
    return out


def permute(t, perm, alpha=1.0, add=None, gamma=0.0, out=None, beta=0.0):
    """Contiguous alpha * t.permute(perm) (+ gamma*add + beta*out)."""
    shape = [t.shape[p] for p in perm]
    strides = [t.stride(p) for p in perm]
    return gather(t, shape, strides, alpha=alpha, add=add, gamma=gamma, out=out, beta=beta)


def contiguous(t):
    if t.is_contiguous():
        return t
    return gather(t, list(t.shape), list(t.stride()))


# --------------------------------------------------------------------------------- elementwise
def _ew(op, n, alpha, x, y, c, beta, out):
    _cabi.check(_cabi.lib().adccb_elementwise(_stream(), op, n, float(alpha), _ptr(x), _ptr(y),
                                              float(c), float(beta), _ptr(out)))
    return out


def fill(t, value):
    return _ew(EW_FILL, t.numel(), 1.0, None, None, value, 0.0, t)


def axpby(alpha, x, beta, y):
    """y = alpha*x + beta*y (contiguous, same numel)."""
    return _ew(EW_AXPBY, x.numel(), alpha, x, None, 0.0, beta, y)


def scaled_copy(x, alpha=1.0):
    x = contiguous(x)
    return _ew(EW_AXPBY, x.numel(), alpha, x, None, 0.0, 0.0, torch.empty_like(x))


def mul(x, y, alpha=1.0, out=None, beta=0.0):
    """out = alpha * x * y + beta * out (new tensor if out is None)."""
    if out is None:
        out, beta = torch.empty_like(x), 0.0
    return _ew(EW_MUL, x.numel(), alpha, x, y, 0.0, beta, out)


def div(x, y, alpha=1.0):
    return _ew(EW_DIV, x.numel(), alpha, x, y, 0.0, 0.0, torch.empty_like(x))


def div_shift(x, d, shift):
    """x / (d - shift): Jacobi preconditioner."""
    return _ew(EW_DIV_SHIFT, x.numel(), 1.0, x, d, shift, 0.0, torch.empty_like(x))


def add_scalar(x, c):
    return _ew(EW_ADD_SCALAR, x.numel(), 1.0, x, None, c, 0.0, torch.empty_like(x))


def direct_sum(a, b, sa=1.0, sb=1.0):
    """out[p..., q...] = sa*a[p...] + sb*b[q...]"""
    a, b = contiguous(a), contiguous(b)
    out = empty(tuple(a.shape) + tuple(b.shape))
    _cabi.check(_cabi.lib().adccb_direct_sum(_stream(), a.numel(), b.numel(), float(sa), _ptr(a),
                                             float(sb), _ptr(b), _ptr(out)))
    return out


# --------------------------------------------------------------------------------- vectors
def add_columns(S, col0, G, ncols):
    """S[:, col0:col0+ncols] += G[:, :ncols] (2-D row-major, unit column stride)."""
    _cabi.check(_cabi.lib().adccb_add_columns(_stream(), S.shape[0], int(ncols), _ptr(S), S.stride(0),
                                              int(col0), _ptr(G), G.stride(0)))
    return S


def lincomb(coeffs, tensors, out=None, beta=0.0):
    n = tensors[0].numel()
    for t in tensors:
        if not t.is_contiguous() or t.numel() != n:
            raise ValueError("lincomb: tensors must be contiguous and equally sized")
    if out is None:
        out = torch.empty_like(tensors[0])
        beta = 0.0
    k = len(tensors)
    cs = (ctypes.c_double * k)(*[float(c) for c in coeffs])
    ps = (ctypes.c_void_p * k)(*[t.data_ptr() for t in tensors])
    _cabi.check(_cabi.lib().adccb_lincomb(_stream(), k, cs, ps, n, float(beta), _ptr(out)))
    return out


def lincomb_multi(coeff_rows, tensors):
    """[sum_k coeff_rows[j][k] * tensors[k] for j]: all outputs from ONE pass over the inputs."""
    n = tensors[0].numel()
    for t in tensors:
        if not t.is_contiguous() or t.numel() != n:
            raise ValueError("lincomb_multi: tensors must be contiguous and equally sized")
    C = np.ascontiguousarray(coeff_rows, dtype=np.float64)
    m, k = C.shape
    if k != len(tensors):
        raise ValueError("lincomb_multi: one coefficient per tensor and output expected")
    outs = [torch.empty_like(tensors[0]) for _ in range(m)]
    cs = C.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    ps = (ctypes.c_void_p * k)(*[t.data_ptr() for t in tensors])
    os_ = (ctypes.c_void_p * m)(*[o.data_ptr() for o in outs])
    _cabi.check(_cabi.lib().adccb_lincomb_multi(_stream(), k, m, cs, ps, n, 0.0, os_))
    return outs


def dot_many(x, ys, scale=1.0):
    """numpy array of <x, y_j>; one pass over x and all y_j; one D2H copy."""
    n = x.numel()
    k = len(ys)
    if k == 0:
        return np.zeros(0)
    for t in [x] + list(ys):
        if not t.is_contiguous() or t.numel() != n:
            raise ValueError("dot_many: tensors must be contiguous and equally sized")
    out = torch.empty(k, dtype=F64, device=x.device)
    ws = workspace()
    ps = (ctypes.c_void_p * k)(*[t.data_ptr() for t in ys])
    _cabi.check(_cabi.lib().adccb_dot_many(_stream(), k, _ptr(x), ps, n, float(scale), 0.0,
                                           _ptr(out), _ptr(ws), ws.numel() * 8))
    return out.cpu().numpy()


def dot_many_blocks(pairs, scale=1.0, out=None, scales=None):
    """sum over blocks b of <x_b, y_bj> for pairs = [(x_b, [y_b0 .. y_bk-1]), ...]: the partial sums
    accumulate on the device (beta = 1 from the second block on), ONE D2H copy at the end.
    With ``out`` (device array of length k) the result stays on the device and nothing is read.
    ``scales``: one weight per block (default: ``scale`` for all)."""
    k = len(pairs[0][1])
    if k == 0:
        return np.zeros(0)
    keep = out is not None
    if not keep:
        out = torch.empty(k, dtype=F64, device=pairs[0][0].device)
    ws = workspace()
    for ib, (x, ys) in enumerate(pairs):
        n = x.numel()
        if len(ys) != k:
            raise ValueError("dot_many_blocks: every block needs the same number of vectors")
        for t in [x] + list(ys):
            if not t.is_contiguous() or t.numel() != n:
                raise ValueError("dot_many: tensors must be contiguous and equally sized")
        ps = (ctypes.c_void_p * k)(*[t.data_ptr() for t in ys])
        _cabi.check(_cabi.lib().adccb_dot_many(_stream(), k, _ptr(x), ps, n,
                                               float(scale if scales is None else scales[ib]),
                                               0.0 if ib == 0 else 1.0, _ptr(out), _ptr(ws),
                                               ws.numel() * 8))
    return out if keep else out.cpu().numpy()


def to_host(t):
    return t.cpu().numpy()


def dot(x, y):
    return float(dot_many(x, [y])[0])


# --------------------------------------------------------------------------------- spin
def spin_expand4(T):
    """Spatial chemists' block (p'q'|r's') -> spin-orbital block, axes [alpha|beta], spin-forbidden = 0."""
    T = contiguous(T)
    n = (ctypes.c_int * 4)(*[int(x) for x in T.shape])
    out = empty(tuple(2 * int(x) for x in T.shape))
    _cabi.check(_cabi.lib().adccb_spin_expand4(_stream(), n, _ptr(T), _ptr(out)))
    return out


def spin_project(t, n_alpha, fac):
    t = contiguous(t)
    nd = t.dim()
    out = torch.empty_like(t)
    sh = (ctypes.c_int * nd)(*[int(s) for s in t.shape])
    na = (ctypes.c_int * nd)(*[int(s) for s in n_alpha])
    _cabi.check(_cabi.lib().adccb_spin_project(_stream(), nd, sh, na, float(fac), _ptr(t), _ptr(out)))
    return out


def spin_singlet_(t, n_alpha):
    if not t.is_contiguous() or t.dim() != 4:
        raise ValueError("spin_singlet_: contiguous rank-4 tensor expected")
    sh = (ctypes.c_int * 4)(*[int(s) for s in t.shape])
    na = (ctypes.c_int * 4)(*[int(s) for s in n_alpha])
    _cabi.check(_cabi.lib().adccb_spin_singlet(_stream(), sh, na, _ptr(t)))
    return t


def precond_apply(x, D, shift, n_alpha, sym_mode=0, spin_fac=0.0, purify=False):
    """Purify(Sym(SpinProject(x / (D - shift)))) in one launch (see include/adccb.h); x, D: contiguous
    dense blocks of rank 2 or 4 (D may be None)."""
    if not x.is_contiguous() or (D is not None and not D.is_contiguous()):
        raise ValueError("precond_apply: contiguous blocks expected")
    nd = x.dim()
    sh = (ctypes.c_int * nd)(*[int(s) for s in x.shape])
    na = (ctypes.c_int * nd)(*[int(s) for s in n_alpha])
    out = torch.empty_like(x)
    _cabi.check(_cabi.lib().adccb_precond_apply(_stream(), nd, sh, na, _ptr(x), _ptr(D), float(shift),
                                                int(sym_mode), float(spin_fac), 1 if purify else 0, _ptr(out)))
    return out


# --------------------------------------------------------------------------------- packed doubles
def n_pairs(n):
    return n * (n - 1) // 2


def packed_unpack(U, no, nv):
    out = empty((no, no, nv, nv))
    _cabi.check(_cabi.lib().adccb_packed_unpack(_stream(), no, nv, _ptr(U), _ptr(out)))
    return out


def packed_pack(T, no, nv):
    T = contiguous(T)
    U = empty((n_pairs(no), n_pairs(nv)))
    _cabi.check(_cabi.lib().adccb_packed_pack(_stream(), no, nv, _ptr(T), _ptr(U)))
    return U


def packed_expand(which, U, no, nv, out=None, n_pairs_o=None):
    """which: 0 X[i,a,k,c]; 1 Ue[i,j,AB]; 2 Uh[a,JK,b] (n_pairs_o: number of JK rows in U when a row
    slab of the packed vector is expanded)."""
    npo = n_pairs(no) if n_pairs_o is None else n_pairs_o
    shape = {0: (no * nv, no * nv), 1: (no, no * n_pairs(nv)), 2: (nv, npo * nv)}[which]
    if out is None:
        out = empty(shape)
    _cabi.check(_cabi.lib().adccb_packed_expand(_stream(), which, no if n_pairs_o is None else -npo, nv,
                                                _ptr(U), _ptr(out)))
    return out


def packed_expand_slab(which, U, no, nv, first, count):
    """Slab forms of packed_expand: 0: X rows (il, a), i = first + il < first + count -> [count*nv, no*nv];
    1: Ue[i, jl, AB], j = first + jl -> [no, count*npv]; 2: Uh[a, r, b] for the `count` packed rows
    of the row slab U -> [nv, count*nv]."""
    shape = {0: (count * nv, no * nv), 1: (no, count * n_pairs(nv)), 2: (nv, count * nv)}[which]
    out = empty(shape)
    _cabi.check(_cabi.lib().adccb_packed_expand_slab(_stream(), which, no, nv, int(first), int(count),
                                                     _ptr(U), _ptr(out)))
    return out


def slab_layout(direction, full, slabs, cuts, wmax):
    """direction 0: full [nrows, npv] -> slabs [world, nrows, wmax] (zero padded); 1: slabs -> full."""
    world = len(cuts) - 1
    nrows, npv = full.shape
    if tuple(slabs.shape) != (world, nrows, wmax) or not full.is_contiguous() or not slabs.is_contiguous():
        raise ValueError("slab_layout: shape mismatch")
    c = (ctypes.c_int64 * (world + 1))(*[int(x) for x in cuts])
    _cabi.check(_cabi.lib().adccb_slab_layout(_stream(), int(direction), nrows, npv, world, c, int(wmax),
                                              _ptr(full), _ptr(slabs)))
    return slabs if direction == 0 else full


def copy2d(dst, src, nrows, ncols):
    """dst[:nrows, :ncols] = src[:nrows, :ncols] between 2-D FP64 views with unit column stride that
    live in device or pinned host memory, on the current stream (one cudaMemcpy2DAsync)."""
    if dst.stride(1) != 1 or src.stride(1) != 1:
        raise ValueError("copy2d: unit column stride expected")
    _cabi.check(_cabi.lib().adccb_copy2d(_stream(), ctypes.c_void_p(dst.data_ptr()), dst.stride(0) * 8,
                                         ctypes.c_void_p(src.data_ptr()), src.stride(0) * 8,
                                         int(ncols) * 8, int(nrows)))
    return dst


def copy2d_raw(dst_ptr, dst_pitch, src_ptr, src_pitch, nrows, ncols):
    """copy2d between raw device addresses (pitches in elements): pulls from peer-mapped memory."""
    _cabi.check(_cabi.lib().adccb_copy2d(_stream(), ctypes.c_void_p(int(dst_ptr)), int(dst_pitch) * 8,
                                         ctypes.c_void_p(int(src_ptr)), int(src_pitch) * 8,
                                         int(ncols) * 8, int(nrows)))


# --------------------------------------------------------------------------------- collectives
# NCCL over NVLink / NVSwitch through torch.distributed (plumbing): equal-sized slabs only.
def comm_all_gather(out, inp, group=None):
    import torch.distributed as dist
    dist.all_gather_into_tensor(out, inp, group=group)
    return out


def comm_reduce_scatter(out, inp, group=None):
    import torch.distributed as dist
    dist.reduce_scatter_tensor(out, inp, group=group)
    return out


def comm_all_reduce(t, group=None):
    import torch.distributed as dist
    dist.all_reduce(t, group=group)
    return t


def _idx_ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def block_gather(inp, rows=None, cols=None):
    """Compact copy inp[rows][:, cols] of a row-major 2-D array (rows / cols: device int64 lists)."""
    if inp.stride(1) != 1:
        raise ValueError("block_gather: unit column stride expected")
    nr = rows.numel() if rows is not None else inp.shape[0]
    nc = cols.numel() if cols is not None else inp.shape[1]
    out = empty((nr, nc))
    _cabi.check(_cabi.lib().adccb_block_gather(_stream(), nr, nc, _idx_ptr(rows), _idx_ptr(cols), _ptr(inp),
                                               inp.stride(0), _ptr(out)))
    return out


def block_scatter_add(out, blk, rows=None, cols=None, alpha=1.0, assign=False):
    """out[rows][:, cols] += alpha * blk  (assign: = instead of +=)."""
    if out.stride(1) != 1 or not blk.is_contiguous():
        raise ValueError("block_scatter_add: unit column stride expected")
    nr, nc = blk.shape
    fn = _cabi.lib().adccb_block_scatter if assign else _cabi.lib().adccb_block_scatter_add
    _cabi.check(fn(_stream(), nr, nc, _idx_ptr(rows), _idx_ptr(cols), float(alpha), _ptr(blk), _ptr(out),
                   out.stride(0)))
    return out


def vvvv_exchange_slab(W, nv, a_begin, na):
    """[na, nv, nv, nv] array out[al,b,c,d] = <a c||b d>, a = a_begin + al, from packed W[AB,CD]."""
    out = empty((na, nv, nv, nv))
    _cabi.check(_cabi.lib().adccb_vvvv_exchange_slab(_stream(), int(nv), int(a_begin), int(na), _ptr(W),
                                                     _ptr(out)))
    return out


def ovvv_exchange_slab(V2, no, nv, a_begin, na):
    """[na, nv, no, nv] array out[al,c,j,d] = <j c||a d>, a = a_begin + al, from V2[c,(j,AD)]."""
    out = empty((na, nv, no, nv))
    _cabi.check(_cabi.lib().adccb_ovvv_exchange_slab(_stream(), int(no), int(nv), int(a_begin), int(na),
                                                     _ptr(V2), _ptr(out)))
    return out


def pair_matvec(V, nv, n_out, n_red, row_stride_out, row_stride_red, X, x_by_out, out=None, alpha=1.0, beta=0.0):
    """out[o, a] = beta*out + alpha * sum_r sum_c A_rho[a,c] x[c] over pair-packed antisymmetric rows
    rho = o*row_stride_out + r*row_stride_red of V; x = X[o] (x_by_out) or X[r] (see include/adccb.h)."""
    _check_dev(V, X, out)
    if out is None:
        out, beta = empty((n_out, nv)), 0.0
    if not V.is_contiguous() or X.stride(-1) != 1 or out.stride(-1) != 1:
        raise ValueError("pair_matvec: contiguous operands expected")
    ws = workspace()
    _cabi.check(_cabi.lib().adccb_pair_matvec(
        _stream(), int(nv), int(n_out), int(n_red), int(row_stride_out), int(row_stride_red), _ptr(V), _ptr(X),
        X.stride(0), 1 if x_by_out else 0, float(alpha), float(beta), _ptr(out), out.stride(0), _ptr(ws),
        ws.numel() * 8))
    return out


def packed_fold_virt(S, nv, sa, P1, off1, P2=None, off2=None, alpha=1.0, rows=None):
    """S[rows[r],AB] += alpha (F_r[a,b] - F_r[b,a]); off1/off2/rows: int64 device tensors."""
    nrows = off1.numel()
    if nrows == 0:
        return S
    _cabi.check(_cabi.lib().adccb_packed_fold_virt(
        _stream(), nrows, nv, int(sa), _ptr(P1), ctypes.c_void_p(off1.data_ptr()), _ptr(P2),
        ctypes.c_void_p(off2.data_ptr()) if off2 is not None else ctypes.c_void_p(0),
        ctypes.c_void_p(rows.data_ptr()) if rows is not None else ctypes.c_void_p(0),
        float(alpha), _ptr(S)))
    return S


def packed_fold_occ(S, C, no, nv, alpha=1.0, j_begin=0, nj=None):
    nj = no if nj is None else nj
    _cabi.check(_cabi.lib().adccb_packed_fold_occ(_stream(), no, nv, int(j_begin), int(nj), _ptr(C),
                                                  float(alpha), _ptr(S)))
    return S


def packed_diagonal(no, nv, dov, doo=None, dvv=None):
    D = empty((n_pairs(no), n_pairs(nv)))
    _cabi.check(_cabi.lib().adccb_packed_diagonal(_stream(), no, nv, _ptr(dov), _ptr(doo), _ptr(dvv),
                                                  _ptr(D)))
    return D


def pair_select(t, axis):
    """Keep p<q of the adjacent equal-extent axes (axis, axis+1) of a contiguous tensor."""
    t = contiguous(t)
    n = t.shape[axis]
    assert t.shape[axis + 1] == n
    L = int(np.prod(t.shape[:axis], dtype=np.int64)) if axis > 0 else 1
    R = int(np.prod(t.shape[axis + 2:], dtype=np.int64)) if axis + 2 < t.dim() else 1
    out = empty(tuple(t.shape[:axis]) + (n_pairs(n),) + tuple(t.shape[axis + 2:]))
    _cabi.check(_cabi.lib().adccb_pair_select(_stream(), L, n, R, _ptr(t), _ptr(out)))
    return out


def int64_tensor(values):
    return torch.tensor([int(v) for v in values], dtype=torch.int64, device=device())


SYNTH_LAYOUTS = {"vvvv_packed": 0, "oooo_packed": 1, "ovov_jbkc": 2, "ovvv_jabc": 3, "ovvv_ajbc": 4,
                 "ooov_ijkb": 5, "ooov_ijak": 6, "dense": 7, "vvvv_exchange": 8, "ovvv_exchange": 9}
SYNTH_BLOCKS = {"oooo": 0, "ooov": 1, "oovv": 2, "ovov": 3, "ovvv": 4, "vvvv": 5}


def synth_fill(out, layout, no, nv, seed, scale, block="vvvv", row_begin=0, j_begin=0, nj=0):
    _cabi.check(_cabi.lib().adccb_synth_fill(_stream(), SYNTH_LAYOUTS[layout], SYNTH_BLOCKS[block], no, nv,
                                             int(seed), float(scale), out.numel(), int(row_begin),
                                             int(j_begin), int(nj), _ptr(out)))
    return out
