"""CPU tests of the oracle itself: it must reproduce (i) the dense stencil /
matrix algebra it restates and (ii) every known-answer vector the reference's
own tests hold for this path."""
import math

import numpy as np
import pytest

import pylops_mpi_oracle as o


# ---- partition bookkeeping: tests/test_distributedarray.py:29-48, 87-126 ----------
@pytest.mark.parametrize("shape,axis", [((500, 501), 1), ((501, 500), 0), ((200, 201, 101), 1),
                                        ((200, 201, 101), 2), ((7,), 0)])
@pytest.mark.parametrize("P", [1, 2, 3, 4, 9])
def test_local_split_sums_and_remainder(shape, axis, P):
    ext = [o.local_split(shape, P, r, o.SCATTER, axis)[axis] for r in range(P)]
    assert sum(ext) == shape[axis]
    assert max(ext) - min(ext) <= 1
    assert ext == sorted(ext, reverse=True)          # remainder to the low ranks
    for r in range(P):
        assert o.local_split(shape, P, r, o.BROADCAST, axis) == shape


def test_to_dist_asarray_roundtrip():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((11, 21))
    for P in (1, 2, 3, 4, 8):
        for axis in (0, 1):
            loc = o.to_dist(x, P, axis=axis)
            assert np.array_equal(o.asarray(loc, axis=axis), x)


# ---- dot / norm: tests/test_distributedarray.py:62-84, 201-222, 270-361 --------------
@pytest.mark.parametrize("shape", [(600, 600), (1200,)])
@pytest.mark.parametrize("P", [1, 2, 4])
def test_dot_norm_vs_numpy(shape, P):
    np.random.seed(42)
    a = np.random.normal(100, 100, shape)
    b = np.random.normal(300, 300, shape)
    al, bl = o.to_dist(a, P), o.to_dist(b, P)
    np.testing.assert_allclose(o.dot(al, bl)[0], np.dot(a.flatten(), b.flatten()), rtol=1e-14)
    for ord_ in (1, 2, np.inf, -np.inf, 0, 3):
        np.testing.assert_allclose(o.norm(al, ord_)[0], np.linalg.norm(a.flatten(), ord_), rtol=1e-13)
    # BROADCAST operands are re-scattered: every element counted once
    np.testing.assert_allclose(o.dot([a] * P, [b] * P, partition=o.BROADCAST)[0],
                               np.dot(a.flatten(), b.flatten()), rtol=1e-14)


def test_masked_dot_counts_subgroups():
    # test_distributedarray.py:270-311: mask groups reduce independently
    P, n = 4, 120
    mask = [0, 0, 1, 1]
    x = np.arange(n, dtype=float)
    xl = o.to_dist(x, P)
    d = o.dot(xl, xl, mask=mask)
    assert d[0] == d[1] == sum(np.dot(xl[r], xl[r]) for r in (0, 1))
    assert d[2] == d[3] == sum(np.dot(xl[r], xl[r]) for r in (2, 3))


# ---- MPIFirstDerivative: plot_derivative.py:36-43 KAT + dense stencil -------------------
def test_first_derivative_kat_config1():
    x = np.zeros((11, 21))
    x[5, 10] = 1.0
    y = np.concatenate(o.first_derivative(o.to_dist(x.ravel(), 2), (11, 21))).reshape(11, 21)
    expect = np.zeros((11, 21))
    expect[4, 10], expect[6, 10] = 0.5, -0.5
    assert np.array_equal(y, expect)


@pytest.mark.parametrize("dims", [(600,), (100, 151), (101, 51, 10), (79, 11, 5), (11, 21)])
@pytest.mark.parametrize("kind,order", [("forward", 3), ("backward", 3), ("centered", 3), ("centered", 5)])
@pytest.mark.parametrize("edge", [False, True])
def test_first_derivative_equals_dense_stencil(dims, kind, order, edge):
    rng = np.random.default_rng(42)
    N, n = dims[0], int(np.prod(dims))
    for P in (1, 2, 3, 4):
        for h in (1.0, 0.4):
            x = rng.normal(0, 10, n)
            D = o.first_derivative_dense(N, h, kind, edge, order)
            X = x.reshape(N, -1)
            try:
                y = np.concatenate(o.first_derivative(o.to_dist(x, P), dims, h, kind, edge, order, False))
                ya = np.concatenate(o.first_derivative(o.to_dist(x, P), dims, h, kind, edge, order, True))
            except (ValueError, IndexError):
                continue   # the reference itself cannot run this split (halo > neighbour extent)
            np.testing.assert_allclose(y, (D @ X).ravel(), rtol=1e-12, atol=1e-12)
            np.testing.assert_allclose(ya, (D.T @ X).ravel(), rtol=1e-12, atol=1e-12)


def test_reference_limit_config1_p8():
    # (11,21) flat split over 8 ranks cannot be re-partitioned with neighbours only
    x = np.zeros(231)
    with pytest.raises(ValueError):
        o.first_derivative(o.to_dist(x, 8), (11, 21))


# ---- BlockDiag / VStack KATs: test_blockdiag.py:24-71, test_stack.py:29-79 ------------------
@pytest.mark.parametrize("ny,nx", [(101, 101), (301, 101)])
@pytest.mark.parametrize("P", [1, 2, 4])
@pytest.mark.parametrize("cx", [False, True])
def test_blockdiag_vstack_kats(ny, nx, P, cx):
    dt = np.complex128 if cx else np.float64
    blocks = [[((r + 1) * np.ones((ny, nx))).astype(dt)] for r in range(P)]
    x = o.to_dist(np.ones(P * nx, dtype=dt), P)
    y = o.blockdiag(blocks, x)
    for r in range(P):
        np.testing.assert_allclose(y[r], (r + 1) * nx * np.ones(ny), rtol=1e-14)
    xa = o.blockdiag(blocks, o.to_dist(np.ones(P * ny, dtype=dt), P), adjoint=True)
    for r in range(P):
        np.testing.assert_allclose(xa[r], (r + 1) * ny * np.ones(nx), rtol=1e-14)
    yv = o.vstack_matvec(blocks, np.ones(nx, dtype=dt))
    for r in range(P):
        np.testing.assert_allclose(yv[r], (r + 1) * nx * np.ones(ny), rtol=1e-14)
    xv = o.vstack_rmatvec(blocks, o.to_dist(np.ones(P * ny, dtype=dt), P))
    np.testing.assert_allclose(xv, sum(r + 1 for r in range(P)) * ny * np.ones(nx), rtol=1e-14)


# ---- MatrixMult KATs: test_matrixmult.py:37-60, 82-166, 199-271 ---------------------------
@pytest.mark.parametrize("N,K,M,dt", [(64, 64, 64, np.float64), (37, 37, 37, np.float64),
                                      (50, 30, 40, np.float64), (22, 20, 16, np.complex128),
                                      (3, 4, 5, np.float32), (1, 2, 1, np.float64), (2, 1, 3, np.float32)])
@pytest.mark.parametrize("P0", [1, 4, 9])
def test_matrixmult_kats(N, K, M, dt, P0):
    ad = min(N, M, math.isqrt(P0))
    P = ad * ad
    A = np.arange(N * K, dtype=dt).reshape(N, K)
    X = np.arange(K * M, dtype=dt).reshape(K, M)
    if np.issubdtype(dt, np.complexfloating):
        A = A + 0.5j * A
        X = X + 0.7j * X
    rtol = np.finfo(dt).resolution
    Y_ref = A.astype(np.complex128 if np.iscomplexobj(A) else np.float64) @ X
    At = o.summa_tiles(A, P)
    Xt = [t.flatten() for t in o.summa_tiles(X, P)]
    Y = o.summa_matvec(At, Xt, N, K, M, dtype=dt)
    np.testing.assert_allclose(o.block_gather(Y, (N, M)), Y_ref, rtol=rtol * 10)
    Xa = o.summa_matvec(At, Y, N, K, M, dtype=dt, adjoint=True)
    np.testing.assert_allclose(o.block_gather(Xa, (K, M)), A.conj().T @ (A @ X), rtol=rtol * 100)
    # block variant
    blk, bc = int(math.ceil(N / ad)), int(math.ceil(M / ad))
    Arows = [A[(r % ad) * blk:min(N, (r % ad + 1) * blk)] for r in range(P)]
    Xc = [X[:, (r // ad) * bc:min(M, (r // ad + 1) * bc)].flatten() for r in range(P)]
    Yb = o.blockmm_matvec(Arows, Xc, N, K, M, dtype=dt)
    for r in range(P):
        cs = (r // ad) * bc
        np.testing.assert_allclose(Yb[r].reshape(N, -1), Y_ref[:, cs:min(M, cs + bc)], rtol=rtol * 10)


# ---- Fredholm1 KAT: test_fredholm.py:36-95, 114-167 -------------------------------------
@pytest.mark.parametrize("nz", [5, 1])
@pytest.mark.parametrize("cx", [False, True])
@pytest.mark.parametrize("P", [1, 2, 3])
def test_fredholm1_kat(nz, cx, P):
    nsl, nx, ny = 21, 4, 6
    G = np.arange(nsl * nx * ny, dtype=np.float64).reshape(nsl, nx, ny)
    if cx:
        G = G - 1j * G
    x = np.ones((nsl, ny, nz)) + (1j * np.ones((nsl, ny, nz)) if cx else 0)
    ext = [o.local_split((nsl,), P, r)[0] for r in range(P)]
    off = np.cumsum([0] + ext)
    G_loc = [G[off[r]:off[r + 1]] for r in range(P)]
    y = o.fredholm1(G_loc, x.ravel(), nz)
    np.testing.assert_allclose(y, np.matmul(G, x).ravel(), rtol=1e-14)
    xa = o.fredholm1(G_loc, y, nz, adjoint=True)
    np.testing.assert_allclose(xa, np.matmul(G.conj().transpose(0, 2, 1), np.matmul(G, x)).ravel(), rtol=1e-13)


# ---- CGLS: test_solver.py:44-100,150-196 (block = A^H A + 1e-5 I, A = ones) ------------------
@pytest.mark.parametrize("ny,nx", [(11, 11), (31, 11)])
@pytest.mark.parametrize("P", [1, 2, 4])
def test_cgls_matches_dense_solution(ny, nx, P):
    rng = np.random.default_rng(42)
    blocks = []
    for r in range(P):
        A = np.ones((ny, nx)) * (r + 1)
        blocks.append([A.conj().T @ A + 1e-5 * np.eye(nx)])
    xt = rng.normal(1, 10, P * nx)
    mv = lambda v: o.SimArray(o.blockdiag(blocks, v.locs))            # noqa: E731
    rmv = lambda v: o.SimArray(o.blockdiag(blocks, v.locs, adjoint=True))  # noqa: E731
    y = mv(o.SimArray(o.to_dist(xt, P)))
    x0 = o.SimArray(o.to_dist(np.zeros(P * nx), P))
    xinv, istop, iit, r1, r2, cost = o.cgls(mv, rmv, y, x0, niter=nx, tol=1e-5)
    # serial restatement on the dense block-diagonal matrix must give the same iterates
    import scipy.linalg as sl
    Dm = sl.block_diag(*[b[0] for b in blocks])
    mvs = lambda v: o.SimArray([Dm @ v.locs[0]])      # noqa: E731
    rmvs = lambda v: o.SimArray([Dm.T @ v.locs[0]])   # noqa: E731
    xs, _, its, _, _, costs = o.cgls(mvs, rmvs, o.SimArray([Dm @ xt]), o.SimArray([np.zeros(P * nx)]),
                                     niter=nx, tol=1e-5)
    assert iit == its
    np.testing.assert_allclose(xinv.asarray(), xs.asarray(), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(cost, costs, rtol=1e-6, atol=1e-9)
    assert cost[-1] < cost[0]


def test_dottest_config1():
    P, dims = 2, (11, 21)
    u = o.SimArray([np.random.default_rng(42 + r).normal(r, 10, s) for r, s in
                    enumerate([116, 115])])
    v = o.SimArray([np.random.default_rng(52 + r).normal(r, 10, s) for r, s in
                    enumerate([126, 105])])
    mv = lambda a: o.SimArray(o.first_derivative(a.locs, dims))                   # noqa: E731
    rmv = lambda a: o.SimArray(o.first_derivative(a.locs, dims, adjoint=True))    # noqa: E731
    ok, xx, yy = o.dottest(mv, rmv, u, v)
    assert ok


def test_cgls_blockdiag_cost_is_rounding_noise_below_1e_6():
    """Conditioning argument behind the tolerance of the multi-rank CGLS parity check (tests/multi_worker.py,
    tests/parity_checks.py; reference test: tests/test_solver.py:150-196 at P ranks).  The operator is block
    diagonal with blocks (r+1)^2 ones^T ones + 1e-5 I: P distinct large eigenvalues and a 1e-5 cluster, so CGLS
    drives the residual to ~1e-8 of its start in P iterations and everything after that is rounding noise.
    A ONE-ULP relative perturbation of the operator output changes those late cost entries by more than 10 % (and
    can move the stopping iteration by one) at P = 8, while every entry above 1e-6 * cost[0] is reproduced to
    1e-9: the achievable parity is 1e-6 of the problem scale, which is what the multi-rank checks assert."""
    P, ny, nx = 8, 31, 11
    blocks = []
    for r in range(P):
        A = np.ones((ny, nx)) * (r + 1)
        blocks.append([A.T @ A + 1e-5 * np.eye(nx)])
    xt = np.random.default_rng(42).normal(1, 10, P * nx)

    def run(eps):
        rng = np.random.default_rng(0)

        def wrap(locs):
            return o.SimArray([b * (1 + eps * rng.standard_normal(b.shape)) for b in locs] if eps else locs)
        mv = lambda v: wrap(o.blockdiag(blocks, v.locs))                    # noqa: E731
        rmv = lambda v: wrap(o.blockdiag(blocks, v.locs, adjoint=True))     # noqa: E731
        y = o.SimArray(o.blockdiag(blocks, o.to_dist(xt, P)))
        return o.cgls(mv, rmv, y, o.SimArray(o.to_dist(np.zeros(P * nx), P)), niter=nx, tol=1e-5)

    x0, _, it0, _, _, c0 = run(0.0)
    x1, _, it1, _, _, c1 = run(1e-16)
    k = min(len(c0), len(c1))
    c0, c1 = np.asarray(c0[:k]), np.asarray(c1[:k])
    big = c0 > 1e-6 * c0[0]
    assert abs(it0 - it1) <= 1
    np.testing.assert_allclose(c1[big], c0[big], rtol=1e-9)                      # well-conditioned part: reproducible
    assert np.max(np.abs(c1[~big] - c0[~big]) / c0[~big]) > 0.1                  # noise part: > 10 % from ONE ulp
    np.testing.assert_allclose(c1, c0, rtol=1e-6, atol=1e-6 * c0[0])             # the bound the parity checks use
    np.testing.assert_allclose(x1.asarray(), x0.asarray(), rtol=1e-6, atol=1e-6 * np.abs(xt).max())


@pytest.mark.parametrize("nt", [31, 32, 64])
def test_mdc_frequency_domain_claim_isometry_of_the_inverse_transform_stage(nt):
    """Claim behind MPIMDC(data_domain="frequency") (pylops_mpi_b200/waveeqprocessing/MDC.py): the last two stages of
    the reference chain, F1^H I1^H (MDC.py:55-69: zero-pad the band, inverse real FFT with the sqrt(2) weighting),
    are an ISOMETRY on band-limited spectra whose DC bin is real (and whose band excludes Nyquist) -- so CGLS on
    || I1 F1 d - Fredholm1 I F m || has the same normal equations, hence the same iterates, as the time-domain MDD.
    Checked with the oracle's restatement of the transform (pinned through the reference chain fixtures)."""
    rng = np.random.default_rng(nt)
    nfft = int(np.ceil((nt + 1) / 2))
    nfmax = nfft - 2
    ntr = 7
    z = rng.standard_normal((nfmax, ntr)) + 1j * rng.standard_normal((nfmax, ntr))
    z[0] = z[0].real                                   # physical spectrum: real DC
    zp = np.zeros((nfft, ntr), dtype=complex)
    zp[:nfmax] = z
    t = o._fft_real_adj(zp, nt, False)
    assert abs(np.linalg.norm(t) - np.linalg.norm(z)) <= 1e-12 * np.linalg.norm(z)
    # and I1 F1 is its left inverse on that subspace
    back = o._fft_real(t, nt, False)[:nfmax]
    np.testing.assert_allclose(back, z, rtol=1e-12, atol=1e-12)
    # with an imaginary DC component the isometry fails (the time-domain chain projects it away): documented caveat
    z2 = z.copy()
    z2[0] = z2[0] + 1j
    zp[:nfmax] = z2
    assert np.linalg.norm(o._fft_real_adj(zp, nt, False)) < np.linalg.norm(z2)
