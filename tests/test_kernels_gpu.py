"""Kernel-level parity: every C-ABI entry point against NumPy on the same seeded inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    import torch
    from adcc_b200 import backend
    torch.cuda.set_device(0)
    return backend


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (780, 1000, 2048), (100, 36, 10), (1, 7, 50),
                                   (257, 130, 4098), (40, 400, 40000), (16, 16, 2), (33, 70, 96),
                                   (64, 1000, 512), (12, 300, 100000)])
def test_gemm_tn_tma(be, M, N, K):
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K))
    B = rng.standard_normal((N, K))
    C = be.gemm(be.from_numpy(A), be.from_numpy(B)).cpu().numpy()
    ref = A @ B.T
    assert _rel(C, ref) < 1e-13
    # alpha / beta
    C0 = rng.standard_normal((M, N))
    out = be.from_numpy(C0)
    be.gemm(be.from_numpy(A), be.from_numpy(B), out=out, alpha=-0.5, beta=2.0)
    assert _rel(out.cpu().numpy(), -0.5 * ref + 2.0 * C0) < 1e-13


@pytest.mark.parametrize("hint", [16, 40, 64, 112, 128, 400])
def test_gemm_tile_variants(be, hint):
    rng = np.random.default_rng(5)
    A = rng.standard_normal((300, 514))
    B = rng.standard_normal((333, 514))
    C = be.gemm(be.from_numpy(A), be.from_numpy(B), tile_hint=hint).cpu().numpy()
    assert _rel(C, A @ B.T) < 1e-13


@pytest.mark.parametrize("M,N,K,hint,ndest", [(780, 640, 2050, 0, 3), (100, 36, 130, 0, 1), (6, 28, 28, 0, 7),
                                              (300, 333, 514, 128, 2), (300, 333, 514, 64, 2),
                                              (45, 128, 40000, 0, 2)])
def test_gemm_peer_destinations(be, M, N, K, hint, ndest):
    """adccb_dgemm_peers: every tile (direct and stream-K fix-up) lands in the local slab and in the
    same columns of every additional destination; bit-identical to the plain GEMM.  The extra
    destinations here are ordinary buffers on this GPU (the kernel only sees addresses)."""
    import torch
    rng = np.random.default_rng(M + N + K + ndest)
    A, B = be.from_numpy(rng.standard_normal((M, K))), be.from_numpy(rng.standard_normal((N, K)))
    plain = be.gemm(A, B, tile_hint=hint)
    ld, c0 = N + 256, 128                          # slab columns [c0, c0+N) of wider result buffers
    bufs = [be.from_numpy(np.full((M, ld), -7.0)) for _ in range(ndest + 1)]
    be.gemm_peers(A, B, bufs[0][:, c0:c0 + N], [b.data_ptr() for b in bufs[1:]], col_offset=c0,
                  ldc=ld, tile_hint=hint)
    same_path = M * N * K >= 32 ** 3               # below that the plain GEMM takes the generic kernel
    for b in bufs:
        if same_path:
            assert torch.equal(b[:, c0:c0 + N], plain)
        else:
            assert _rel(b[:, c0:c0 + N].cpu().numpy(), plain.cpu().numpy()) < 1e-13
        assert bool((b[:, :c0] == -7.0).all()) and bool((b[:, c0 + N:] == -7.0).all())
    with pytest.raises(RuntimeError):              # more than 7 peers
        be.gemm_peers(A, B, bufs[0][:, c0:c0 + N], [bufs[0].data_ptr()] * 8, col_offset=c0, ldc=ld)


@pytest.mark.parametrize("M,N,K,hint", [(40, 3000, 400, 0), (8, 130, 37, 0), (64, 1000, 512, 0), (200, 258, 1000, 0),
                                        (40, 100000, 400, 0), (300, 334, 514, 16), (300, 334, 514, 40),
                                        (300, 334, 514, 64), (16, 48, 40000, 0), (2, 79800, 400, 0)])
def test_gemm_n_major_b(be, M, N, K, hint):
    """C = A . B with B stored [K][N] (N contiguous): the TMA kernel's N-major B loads, no transposed
    copy.  Same sums in the same order as the K-contiguous kernel on B^T -> bit-identical to it."""
    import torch
    rng = np.random.default_rng(M + N + K + hint)
    A = rng.standard_normal((M, K))
    Bkn = rng.standard_normal((K, N))
    dA, dB = be.from_numpy(A), be.from_numpy(Bkn)
    C = be.gemm(dA, dB.t(), tile_hint=hint)
    assert _rel(C.cpu().numpy(), A @ Bkn) < 1e-13
    ref = be.gemm(dA, dB.t().contiguous(), tile_hint=hint)          # K-contiguous copy of B
    assert torch.equal(C, ref)
    C0 = rng.standard_normal((M, N))
    out = be.from_numpy(C0)
    be.gemm(dA, dB.t(), out=out, alpha=-0.5, beta=2.0, tile_hint=hint)
    assert _rel(out.cpu().numpy(), -0.5 * (A @ Bkn) + 2.0 * C0) < 1e-13
    # column window of a wider buffer (the j-slab view of V2[a,(j,BC)])
    if N >= 64:
        view = dB[:, 16:16 + N // 2 * 2 - 16]
        Cv = be.gemm(dA, view.t(), tile_hint=hint)
        assert _rel(Cv.cpu().numpy(), A @ Bkn[:, 16:16 + N // 2 * 2 - 16]) < 1e-13


def test_gemm_generic_strides(be):
    rng = np.random.default_rng(7)
    # odd leading dimension -> generic path; transposed operands
    A = rng.standard_normal((37, 51))
    B = rng.standard_normal((29, 51))
    C = be.gemm(be.from_numpy(A), be.from_numpy(B)).cpu().numpy()
    assert _rel(C, A @ B.T) < 1e-13
    At = be.from_numpy(A.T.copy()).t()      # (37,51) view with strides (1,37)
    Bt = be.from_numpy(B.T.copy()).t()
    C = be.gemm(At, Bt).cpu().numpy()
    assert _rel(C, A @ B.T) < 1e-13
    # empty K
    out = be.from_numpy(np.ones((4, 5)))
    be.gemm(be.empty((4, 0)), be.empty((5, 0)), out=out, beta=3.0)
    assert np.allclose(out.cpu().numpy(), 3.0)


@pytest.mark.parametrize("shape,perm", [((6, 5, 4, 3), (1, 0, 3, 2)), ((10, 10, 38, 38), (2, 3, 0, 1)),
                                        ((33, 65), (1, 0)), ((4, 9, 7), (2, 0, 1)), ((5,), (0,)),
                                        ((3, 4, 5, 6, 7), (4, 2, 0, 3, 1)), ((40, 40, 64, 64), (0, 2, 1, 3))])
def test_permute(be, shape, perm):
    rng = np.random.default_rng(3)
    T = rng.standard_normal(shape)
    out = be.permute(be.from_numpy(T), perm).cpu().numpy()
    assert np.array_equal(out, T.transpose(perm))


def test_gather_symmetrise_and_diagonal(be):
    rng = np.random.default_rng(11)
    T = rng.standard_normal((6, 6, 8, 8))
    t = be.from_numpy(T)
    # antisymmetrise(0,1) = 0.5*(T - T.swap) in one pass
    out = be.permute(t, (1, 0, 2, 3), alpha=-0.5, add=t, gamma=0.5).cpu().numpy()
    assert np.allclose(out, 0.5 * (T - T.transpose(1, 0, 2, 3)), atol=1e-15)
    # diagonal "iaia->ia"
    X = rng.standard_normal((5, 7, 5, 7))
    x = be.from_numpy(X)
    d = be.gather(x, (5, 7), (x.stride(0) + x.stride(2), x.stride(1) + x.stride(3))).cpu().numpy()
    assert np.array_equal(d, np.einsum("iaia->ia", X))


def test_elementwise_family(be):
    rng = np.random.default_rng(2)
    x, y = rng.standard_normal(10007), rng.standard_normal(10007) + 3.0
    dx, dy = be.from_numpy(x), be.from_numpy(y)
    assert np.array_equal(be.mul(dx, dy).cpu().numpy(), x * y)
    assert np.array_equal(be.div(dx, dy).cpu().numpy(), x / y)
    assert np.array_equal(be.div_shift(dx, dy, 0.25).cpu().numpy(), x / (y - 0.25))
    assert np.array_equal(be.add_scalar(dx, 1.5).cpu().numpy(), x + 1.5)
    z = be.from_numpy(y)
    be.axpby(2.0, dx, -1.0, z)
    assert np.allclose(z.cpu().numpy(), 2.0 * x - y, rtol=1e-15)
    a, b = rng.standard_normal((4, 5)), rng.standard_normal((3,))
    ds = be.direct_sum(be.from_numpy(a), be.from_numpy(b), 1.0, -1.0).cpu().numpy()
    assert np.array_equal(ds, a[:, :, None] - b[None, None, :])


def test_lincomb_and_dots(be):
    rng = np.random.default_rng(4)
    n, k = 123457, 37
    X = rng.standard_normal((k, n))
    c = rng.standard_normal(k)
    ts = [be.from_numpy(X[i]) for i in range(k)]
    out = be.lincomb(c, ts).cpu().numpy()
    assert _rel(out, c @ X) < 1e-14
    d = be.dot_many(ts[0], ts)
    assert _rel(d, X @ X[0]) < 1e-13
    # deterministic
    assert np.array_equal(d, be.dot_many(ts[0], ts))
    # empty / tiny
    e = be.from_numpy(np.array([2.0]))
    assert be.dot(e, e) == 4.0


@pytest.mark.parametrize("k,m,n", [(3, 2, 1000), (70, 11, 4099), (32, 8, 262144), (1, 1, 5), (40, 5, 100003)])
def test_lincomb_multi(be, k, m, n):
    """adccb_lincomb_multi: m combinations of the same k inputs in one pass, each bit-identical to
    adccb_lincomb on the same coefficients (the Davidson loop relies on that)."""
    import torch
    rng = np.random.default_rng(k + m + n)
    xs = [be.from_numpy(rng.standard_normal(n)) for _ in range(k)]
    C = rng.standard_normal((m, k))
    outs = be.lincomb_multi(C, xs)
    assert len(outs) == m
    for j in range(m):
        assert torch.equal(outs[j], be.lincomb(C[j], xs))
    ref = C @ np.stack([x.cpu().numpy() for x in xs])
    assert _rel(np.stack([o.cpu().numpy() for o in outs]), ref) < 1e-13


def test_spin_kernels(be):
    from oracle.davidson import enforce_spin_maps, enforce_spin_kind
    rng = np.random.default_rng(9)
    na = [3, 3, 4, 4]
    T = rng.standard_normal((6, 6, 8, 8))
    S = rng.standard_normal((6, 8))
    for fac in (1.0, -1.0):
        ref = enforce_spin_maps({"ph": S.copy(), "pphh": T.copy()}, na, fac)
        assert np.allclose(be.spin_project(be.from_numpy(T), na, fac).cpu().numpy(), ref["pphh"], atol=1e-15)
        assert np.allclose(be.spin_project(be.from_numpy(S), [3, 4], fac).cpu().numpy(), ref["ph"], atol=1e-15)
    ref = enforce_spin_kind(T.copy(), na, "singlet")
    t = be.from_numpy(T)
    be.spin_singlet_(t, na)
    assert np.array_equal(t.cpu().numpy(), ref)


def test_gemm_randomised_shapes(be):
    """Ragged / prime extents, tile-boundary cases and stream-K splits, incl. beta accumulation."""
    rng = np.random.default_rng(123)
    shapes = [(1, 1, 1), (7, 3, 5), (113, 127, 131), (112, 128, 16), (129, 257, 1000), (780, 2048, 4096),
              (40, 4001, 402), (2, 100003, 64), (300, 300, 30002), (17, 19, 100000), (1000, 9, 2050)]
    for M, N, K in shapes:
        A = rng.standard_normal((M, K))
        B = rng.standard_normal((N, K))
        C0 = rng.standard_normal((M, N))
        out = be.from_numpy(C0)
        be.gemm(be.from_numpy(A), be.from_numpy(B), out=out, alpha=0.7, beta=-1.3)
        ref = 0.7 * (A @ B.T) - 1.3 * C0
        assert _rel(out.cpu().numpy(), ref) < 1e-12, (M, N, K)


def test_gemm_views_and_unaligned(be):
    import torch
    rng = np.random.default_rng(5)
    big = be.from_numpy(rng.standard_normal((70, 300)))
    A = big[3:60, 10:210]          # row-strided view, 16B-aligned start (offset 10 doubles -> 80 B)
    B = big[5:33, 11:211]          # start offset 11 doubles = 88 B: not 16B aligned -> generic path
    C = be.gemm(A, B)
    assert _rel(C.cpu().numpy(), A.cpu().numpy() @ B.cpu().numpy().T) < 1e-13
    # output into a column slab of a wider matrix (ldc > N)
    wide = be.from_numpy(np.zeros((57, 100)))
    be.gemm(A, big[0:28, 10:210], out=wide[:, 30:58])
    assert _rel(wide.cpu().numpy()[:, 30:58], A.cpu().numpy() @ big[0:28, 10:210].cpu().numpy().T) < 1e-13
    assert torch.count_nonzero(wide[:, :30]).item() == 0 and torch.count_nonzero(wide[:, 58:]).item() == 0


def test_vector_kernels_odd_sizes_and_offsets(be):
    rng = np.random.default_rng(8)
    for n in (1, 31, 2049, 1_000_003, 3_000_000):
        k = 11
        flat = be.from_numpy(rng.standard_normal(k * (n + 1) + 3))
        # vectors that start at odd element offsets (8-byte but not 16-byte aligned) and aligned ones
        for start in (0, 1):
            vs = [flat[start + j * (n + 1): start + j * (n + 1) + n] for j in range(k)]
            ref = np.stack([v.cpu().numpy() for v in vs])
            d = be.dot_many(vs[0], vs)
            assert _rel(d, ref @ ref[0]) < 1e-12, (n, start)
            c = rng.standard_normal(k)
            out = be.lincomb(c, vs)
            assert _rel(out.cpu().numpy(), c @ ref) < 1e-13


def test_gather_randomised_permutations(be):
    rng = np.random.default_rng(21)
    for trial in range(12):
        nd = int(rng.integers(1, 7))
        shape = tuple(int(x) for x in rng.integers(1, 9 if nd > 4 else 40, size=nd))
        perm = tuple(int(x) for x in rng.permutation(nd))
        T = rng.standard_normal(shape)
        add = rng.standard_normal(tuple(shape[p] for p in perm))
        out = be.permute(be.from_numpy(T), perm, alpha=2.0, add=be.from_numpy(add), gamma=-0.5).cpu().numpy()
        assert np.allclose(out, 2.0 * T.transpose(perm) - 0.5 * add, atol=1e-14), (shape, perm)


def test_tensor_symmetrise_reference_goldens_device(be):
    from tensor_golden_common import check_tensor_goldens
    check_tensor_goldens()


@pytest.mark.parametrize("M,N,K,hint", [(780, 4096, 9600, 0), (2048, 2048, 2048, 128), (2048, 2048, 2048, 112),
                                        (40, 240, 96000, 0), (2000, 1600, 1600, 0),
                                        (780, 8192, 12800, 0),        # 2 data-parallel rounds + two-tile stream-K tail
                                        (40, 400, 96000, 0), (37, 390, 50000, 0)])   # one 40 x 400 tile, two B boxes
def test_gemm_pipeline_stress_bitwise(be, M, N, K, hint):
    """Regression guard for the shared-memory WAR race the TMA pipeline once had (a refill
    overtaking delayed fragment reads, about once per 1e4 tiles with L2-resident operands): many
    back-to-back launches on operands that stay in L2, every result compared BITWISE with the first
    one, and the first one with the generic (non-TMA) kernel path."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g, dtype=torch.float64).cuda()
    B = torch.randn(N, K, generator=g, dtype=torch.float64).cuda()
    C0 = be.gemm(A, B, tile_hint=hint).clone()
    # generic kernel: odd leading dimension defeats the TMA alignment rule
    Ap = torch.zeros(M, K + 1, dtype=torch.float64, device="cuda")
    Ap[:, :K] = A
    Cg = be.gemm(Ap[:, :K], B)
    assert float((Cg - C0).abs().max() / C0.abs().max()) < 1e-13
    bad = torch.zeros((), dtype=torch.int64, device="cuda")
    for _ in range(60):
        C = be.gemm(A, B, tile_hint=hint)
        bad += (C != C0).any()
    assert int(bad.item()) == 0


@pytest.mark.parametrize("shape,sym_mode,fac,purify", [((6, 6, 8, 8), 1, 1.0, True), ((6, 6, 8, 8), 1, -1.0, False),
                                                       ((6, 6, 8, 8), 1, 0.0, False), ((6, 4, 8, 8), 2, 1.0, True),
                                                       ((10, 14), 0, 1.0, False), ((10, 14), 0, 0.0, False)])
def test_fused_precondition_symmetrise_spin(be, shape, sym_mode, fac, purify):
    """adccb_precond_apply against the chain of separate steps in NumPy."""
    rng = np.random.default_rng(len(shape) + sym_mode)
    x = rng.standard_normal(shape)
    D = rng.uniform(1.0, 3.0, shape)
    shift = 0.37
    na = [s // 2 for s in shape]
    got = be.precond_apply(be.from_numpy(x), be.from_numpy(D), shift, na, sym_mode=sym_mode, spin_fac=fac,
                           purify=purify).cpu().numpy()
    p = x / (D - shift)
    if fac != 0.0:
        f = p
        for ax, n in enumerate(na):
            f = np.concatenate((np.take(f, range(n, 2 * n), axis=ax), np.take(f, range(n), axis=ax)), axis=ax)
        spin = [np.array([0] * n + [1] * n) for n in na]
        if len(shape) == 2:
            keep = spin[0][:, None] == spin[1][None, :]
        else:
            keep = (spin[0][:, None, None, None] + spin[1][None, :, None, None]
                    == spin[2][None, None, :, None] + spin[3][None, None, None, :])
        p = 0.5 * (p + fac * f) * keep
    if sym_mode == 1:
        p = 0.5 * (p - p.transpose(1, 0, 2, 3))
        p = 0.5 * (p - p.transpose(0, 1, 3, 2))
        p = 0.5 * (p + p.transpose(1, 0, 3, 2))
    elif sym_mode == 2:
        p = 0.5 * (p - p.transpose(0, 1, 3, 2))
    if purify:
        ni, nj, a_, b_ = na
        new = p[:ni, nj:, :a_, b_:] + p[:ni, nj:, a_:, :b_]
        p = p.copy()
        p[:ni, :nj, :a_, :b_] = new
        p[ni:, nj:, a_:, b_:] = new
    assert np.max(np.abs(got - p)) < 1e-14 * max(1.0, np.max(np.abs(p)))


def test_block_gather_scatter(be):
    import torch
    rng = np.random.default_rng(3)
    M = rng.standard_normal((37, 300))
    rows = np.array([3, 0, 36, 17, 5])
    cols = np.sort(rng.choice(300, 70, replace=False))
    d = be.from_numpy(M)
    r, c = (torch.as_tensor(v, dtype=torch.int64, device=d.device) for v in (rows, cols))
    assert np.array_equal(be.block_gather(d, r, c).cpu().numpy(), M[np.ix_(rows, cols)])
    assert np.array_equal(be.block_gather(d, None, c).cpu().numpy(), M[:, cols])
    assert np.array_equal(be.block_gather(d[:, 10:], r, None).cpu().numpy(), M[rows][:, 10:])
    blk = rng.standard_normal((5, 70))
    ref = M.copy()
    ref[np.ix_(rows, cols)] += -0.5 * blk
    be.block_scatter_add(d, be.from_numpy(blk), r, c, alpha=-0.5)
    assert np.allclose(d.cpu().numpy(), ref, atol=1e-15)
    ref[np.ix_(rows, cols)] = 2.0 * blk
    be.block_scatter_add(d, be.from_numpy(blk), r, c, alpha=2.0, assign=True)
    assert np.array_equal(d.cpu().numpy(), ref)
