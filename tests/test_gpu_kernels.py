"""GPU parity tests, kernel level: every libb200lops entry point is called through
the C ABI (ctypes) and compared with the CPU oracle / NumPy on the same seeded
inputs.  Multi-rank semantics of the stencil kernel are exercised on ONE GPU by
invoking the per-rank kernel for each simulated rank with explicit halo rows."""
import ctypes as C

import numpy as np
import pytest
import torch

import pylops_mpi_oracle as o

pytestmark = pytest.mark.gpu

KINDS = {"forward": 0, "backward": 1, "centered": 2}


@pytest.fixture(scope="module")
def L():
    import pylops_mpi_b200._lib as L
    return L


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


# --------------------------------------------------------------------------
# first derivative: per-rank kernel vs oracle per-rank output
# --------------------------------------------------------------------------
def run_fd_rank(L, xg, dims, P, r, kind, order, edge, h, adjoint, code):
    """apply the kernel as rank r of P on the row-block partition of xg (N x ncols)"""
    N = dims[0]
    rows = [o.local_split((N,), P, q)[0] for q in range(P)]
    off = np.cumsum([0] + rows)
    lo_need, hi_need = C.c_int(), C.c_int()
    L.check(L.lib.b2_first_derivative_halo(KINDS[kind], order, int(adjoint), C.byref(lo_need), C.byref(hi_need)))
    r0, r1 = off[r], off[r + 1]
    n_lo = min(lo_need.value, r0)
    n_hi = min(hi_need.value, N - r1)
    x = dev(xg[r0:r1])
    lo = dev(xg[r0 - n_lo:r0]) if n_lo else None
    hi = dev(xg[r1:r1 + n_hi]) if n_hi else None
    y = torch.empty_like(x)
    ncols = xg.shape[1] * (2 if np.iscomplexobj(xg) else 1)
    L.check(L.lib.b2_first_derivative(L.ctx(), x.data_ptr(), y.data_ptr(),
                                      lo.data_ptr() if lo is not None else None, n_lo,
                                      hi.data_ptr() if hi is not None else None, n_hi,
                                      r1 - r0, ncols, r0, N, KINDS[kind], order, int(edge), float(h),
                                      int(adjoint), code, L.stream()), "fd")
    return y.cpu().numpy()


@pytest.mark.parametrize("dims", [(11, 21), (600,), (100, 151), (101, 51, 10), (79, 11, 5), (64, 256), (9, 32)])
@pytest.mark.parametrize("kind,order", [("forward", 3), ("backward", 3), ("centered", 3), ("centered", 5)])
@pytest.mark.parametrize("edge", [False, True])
def test_first_derivative_per_rank_f64(L, dims, kind, order, edge):
    rng = np.random.default_rng(42)
    N, n = dims[0], int(np.prod(dims))
    for P in (1, 2, 3, 4):
        for h in (1.0, 0.4):
            for adjoint in (False, True):
                x = rng.normal(0, 10, n)
                xg = x.reshape(N, -1)
                try:
                    ref = o.first_derivative(o.to_dist(x, P), dims, h, kind, edge, order, adjoint)
                except (ValueError, IndexError):
                    D = o.first_derivative_dense(N, h, kind, edge, order)
                    full = ((D.T if adjoint else D) @ xg)
                    rows = np.cumsum([0] + [o.local_split((N,), P, q)[0] for q in range(P)])
                    ref = [full[rows[q]:rows[q + 1]].ravel() for q in range(P)]
                for r in range(P):
                    got = run_fd_rank(L, xg, dims, P, r, kind, order, edge, h, adjoint, L.F64)
                    np.testing.assert_allclose(got.ravel(), ref[r], rtol=1e-12, atol=1e-12,
                                               err_msg=f"{dims} P={P} r={r} {kind}{order} edge={edge} adj={adjoint}")


def test_first_derivative_kat_bit_exact(L):
    # plot_derivative.py:36-43 / README.md:73-94
    x = np.zeros((11, 21))
    x[5, 10] = 1.0
    y = np.concatenate([run_fd_rank(L, x, (11, 21), 2, r, "centered", 3, False, 1.0, False, L.F64) for r in range(2)])
    expect = np.zeros((11, 21))
    expect[4, 10], expect[6, 10] = 0.5, -0.5
    assert np.array_equal(y, expect)
    ref = np.concatenate(o.first_derivative(o.to_dist(x.ravel(), 2), (11, 21))).reshape(11, 21)
    assert np.array_equal(y, ref)


@pytest.mark.parametrize("kind,order", [("forward", 3), ("centered", 3), ("centered", 5)])
def test_first_derivative_f32_and_complex(L, kind, order):
    rng = np.random.default_rng(7)
    dims = (257, 96)
    x32 = rng.standard_normal(dims).astype(np.float32)
    xc = (rng.standard_normal(dims) + 1j * rng.standard_normal(dims))
    for adjoint in (False, True):
        D = o.first_derivative_dense(dims[0], 0.5, kind, True, order)
        Dm = D.T if adjoint else D
        got = np.concatenate([run_fd_rank(L, x32, dims, 3, r, kind, order, True, 0.5, adjoint, L.F32) for r in range(3)])
        np.testing.assert_allclose(got, Dm @ x32.astype(np.float64), rtol=2e-5, atol=2e-5)
        gotc = np.concatenate([run_fd_rank(L, xc, dims, 2, r, kind, order, True, 0.5, adjoint, L.F64) for r in range(2)])
        np.testing.assert_allclose(gotc, Dm @ xc, rtol=1e-12, atol=1e-12)


def test_first_derivative_missing_halo_is_an_error(L):
    x = torch.zeros((4, 32), dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    rc = L.lib.b2_first_derivative(L.ctx(), x.data_ptr(), y.data_ptr(), None, 0, None, 0, 4, 32, 4, 12,
                                   2, 3, 0, 1.0, 0, L.F64, L.stream())
    assert rc == 2003
    with pytest.raises(L.B200Error):
        L.check(rc, "fd")


def test_first_derivative_large_properties(L):
    """full-size style checks: exact adjointness <Dx,y> = <x,D^T y> and agreement with a
    torch float32 restatement of the stencil on a > L2 array"""
    torch.manual_seed(0)
    N, ncols = 16384, 4096          # 256 MiB float32
    x = torch.randn(N, ncols, device="cuda", dtype=torch.float32)
    v = torch.randn(N, ncols, device="cuda", dtype=torch.float32)
    y = torch.empty_like(x)
    z = torch.empty_like(x)
    for kind, order in ((2, 3), (2, 5), (0, 3)):
        L.check(L.lib.b2_first_derivative(L.ctx(), x.data_ptr(), y.data_ptr(), None, 0, None, 0, N, ncols, 0, N,
                                          kind, order, 0, 1.0, 0, L.F32, L.stream()))
        L.check(L.lib.b2_first_derivative(L.ctx(), v.data_ptr(), z.data_ptr(), None, 0, None, 0, N, ncols, 0, N,
                                          kind, order, 0, 1.0, 1, L.F32, L.stream()))
        lhs = torch.dot(y.double().view(-1), v.double().view(-1)).item()
        rhs = torch.dot(x.double().view(-1), z.double().view(-1)).item()
        assert abs(lhs - rhs) <= 1e-6 * max(abs(lhs), abs(rhs), 1.0)
        ref = torch.zeros_like(x)
        if (kind, order) == (2, 3):
            ref[1:-1] = 0.5 * (x[2:] - x[:-2])
        elif (kind, order) == (2, 5):
            ref[2:-2] = x[:-4] / 12.0 - 2 * x[1:-3] / 3.0 + 2 * x[3:-1] / 3.0 - x[4:] / 12.0
        else:
            ref[:-1] = x[1:] - x[:-1]
        assert torch.allclose(y, ref, rtol=1e-5, atol=1e-5)


def test_first_derivative_host_pipeline(L):
    rng = np.random.default_rng(3)
    N, ncols = 3000, 1024
    x = rng.standard_normal((N, ncols)).astype(np.float32)
    xh = torch.as_tensor(x).pin_memory()
    yh = torch.empty_like(xh).pin_memory()
    for kind, order, adj in ((2, 3, 0), (2, 5, 1), (1, 3, 0)):
        L.check(L.lib.b2_first_derivative_host(L.ctx(), xh.data_ptr(), yh.data_ptr(), N, ncols, 0, N, kind, order, 1,
                                               2.0, adj, L.F32), "fd_host")
        name = {0: "forward", 1: "backward", 2: "centered"}[kind]
        D = o.first_derivative_dense(N, 2.0, name, True, order)
        Dm = D.T if adj else D
        # dense N x N on a column sample keeps the CPU check cheap
        cols = np.arange(0, ncols, 97)
        np.testing.assert_allclose(yh.numpy()[:, cols], Dm @ x[:, cols].astype(np.float64), rtol=2e-5, atol=2e-5)


# --------------------------------------------------------------------------
# element-wise + reductions
# --------------------------------------------------------------------------
DT = {"f32": (np.float32, 0), "f64": (np.float64, 1), "c64": (np.complex64, 2), "c128": (np.complex128, 3)}


def rnd(rng, n, npdt):
    if np.issubdtype(npdt, np.complexfloating):
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(npdt)
    return rng.standard_normal(n).astype(npdt)


@pytest.mark.parametrize("dt", list(DT))
@pytest.mark.parametrize("n", [0, 1, 3, 257, 4099, 1 << 20])
def test_lincomb_mul_fill(L, dt, n):
    npdt, code = DT[dt]
    rng = np.random.default_rng(n + 1)
    x, y = rnd(rng, n, npdt), rnd(rng, n, npdt)
    xd, yd = dev(x), dev(y)
    out = torch.empty_like(xd)
    cx = np.issubdtype(npdt, np.complexfloating)
    tol = dict(rtol=2e-6, atol=2e-6) if dt in ("f32", "c64") else dict(rtol=1e-14, atol=1e-14)
    for a, b, conj in [(1.0, -1.0, 0), (2.5, 0.5, 0), (-1.0, None, 0)] + ([(1.0, None, 1), (0.5 - 2j, 1 + 1j, 0), (1j, 2.0, 1)] if cx else []):
        if n == 0:
            continue
        L.check(L.lib.b2_lincomb(L.ctx(), out.data_ptr(), L.cpair(a), xd.data_ptr(),
                                 L.cpair(b) if b is not None else None, yd.data_ptr() if b is not None else None,
                                 n, code, conj, L.stream()))
        xx = x.conj() if conj else x
        ref = a * xx + (b * y if b is not None else 0)
        np.testing.assert_allclose(out.cpu().numpy(), ref.astype(npdt), **tol)
    if n:
        L.check(L.lib.b2_mul(L.ctx(), out.data_ptr(), xd.data_ptr(), yd.data_ptr(), n, code, 0, L.stream()))
        np.testing.assert_allclose(out.cpu().numpy(), x * y, **tol)
        L.check(L.lib.b2_fill(L.ctx(), out.data_ptr(), L.cpair(3.0 - (1j if cx else 0)), n, code, L.stream()))
        assert np.all(out.cpu().numpy() == npdt(3.0 - (1j if cx else 0)))
        # unaligned views take the scalar path
        if n > 8:
            o2 = torch.empty(n + 1, dtype=xd.dtype, device="cuda")[1:]
            x2 = torch.cat([xd[:1], xd])[1:]
            L.check(L.lib.b2_lincomb(L.ctx(), o2.data_ptr(), L.cpair(2.0), x2.data_ptr(), L.cpair(1.0), yd.data_ptr(),
                                     n, code, 0, L.stream()))
            np.testing.assert_allclose(o2.cpu().numpy(), (2.0 * x + y).astype(npdt), **tol)


@pytest.mark.parametrize("dt", list(DT))
@pytest.mark.parametrize("n", [0, 1, 5, 1023, 65537, 3_000_001])
def test_dot_and_norm_partials(L, dt, n):
    npdt, code = DT[dt]
    rng = np.random.default_rng(n + 11)
    x, y = rnd(rng, n, npdt), rnd(rng, n, npdt)
    if n > 4:
        x[3] = 0
    xd, yd = dev(x), dev(y)
    out = torch.empty(2, dtype=torch.float64, device="cuda")
    x64 = x.astype(np.complex128 if np.iscomplexobj(x) else np.float64)
    y64 = y.astype(x64.dtype)
    scale = np.linalg.norm(x64) * np.linalg.norm(y64) + 1e-300
    for conj in (0, 1):
        L.check(L.lib.b2_dot(L.ctx(), xd.data_ptr() if n else None, yd.data_ptr() if n else None, n, code, conj,
                             out.data_ptr(), L.stream()))
        got = complex(*out.cpu().numpy())
        ref = np.vdot(x64, y64) if conj else np.dot(x64, y64)
        assert abs(got - ref) <= 1e-13 * scale + 1e-300
    o1 = torch.empty(1, dtype=torch.float64, device="cuda")
    a = np.abs(x64)
    refs = {0: np.count_nonzero(x), 1: a.sum(), 2: (a ** 2).sum(), 3: a.max() if n else 0.0,
            4: a.min() if n else np.inf, 5: (a ** 3).sum()}
    for kind, ref in refs.items():
        L.check(L.lib.b2_norm_partial(L.ctx(), xd.data_ptr() if n else None, n, code, kind, 3.0, o1.data_ptr(), L.stream()))
        got = o1.item()
        assert got == ref or abs(got - ref) <= 1e-12 * abs(ref), (kind, got, ref)


def test_dot_multi(L):
    rng = np.random.default_rng(5)
    n = 100_003
    for dt in ("f32", "f64", "c128"):
        npdt, code = DT[dt]
        arrs = [rnd(rng, n, npdt) for _ in range(3)]
        ds = [dev(a) for a in arrs]
        out = torch.zeros(6, dtype=torch.float64, device="cuda")
        ptrs = (C.c_void_p * 3)(*[d.data_ptr() for d in ds])
        L.check(L.lib.b2_dot_multi(L.ctx(), 3, ptrs, ptrs, n, code, 1, out.data_ptr(), L.stream()))
        res = out.cpu().numpy()
        for i, a in enumerate(arrs):
            ref = np.vdot(a.astype(np.complex128), a.astype(np.complex128)).real
            got = res[2 * i] if dt == "c128" else res[i]
            assert abs(got - ref) <= 1e-12 * ref


# --------------------------------------------------------------------------
# gemv / gemm / batched gemm
# --------------------------------------------------------------------------
@pytest.mark.parametrize("dt", list(DT))
@pytest.mark.parametrize("m,n", [(1, 1), (3, 5), (101, 101), (301, 101), (64, 1000), (1000, 64), (513, 1027)])
def test_gemv_all_ops(L, dt, m, n):
    npdt, code = DT[dt]
    rng = np.random.default_rng(m * 1000 + n)
    A = rnd(rng, m * n, npdt).reshape(m, n)
    Ad = dev(A)
    A64 = A.astype(np.complex128 if np.iscomplexobj(A) else np.float64)
    tol = 2e-5 if dt in ("f32", "c64") else 1e-12
    for op, fn in ((0, lambda v: A64 @ v), (1, lambda v: A64.T @ v), (2, lambda v: A64.conj().T @ v)):
        nin, nout = (n, m) if op == 0 else (m, n)
        x = rnd(rng, nin, npdt)
        xd = dev(x)
        yd = torch.empty(nout, dtype=xd.dtype, device="cuda")
        L.check(L.lib.b2_gemv(L.ctx(), Ad.data_ptr(), n, m, n, xd.data_ptr(), yd.data_ptr(), op, code, code, L.stream()))
        ref = fn(x.astype(A64.dtype))
        scale = np.abs(A64).sum(axis=1 if op == 0 else 0).max() * np.abs(x).max() + 1e-30
        np.testing.assert_allclose(yd.cpu().numpy(), ref, rtol=tol, atol=tol * scale)


@pytest.mark.parametrize("dt", list(DT))
@pytest.mark.parametrize("m", [1, 3, 4, 9, 130])
def test_gemv_long_rows_split_kernel(L, dt, m):
    """rows of >= 32 KB take the row-splitting kernel (8 warps sweep 4 rows together): ragged row counts, a row
    pitch larger than n (view of a wider matrix, scalar remainder columns) and every element type vs float64"""
    npdt, code = DT[dt]
    esz = np.dtype(npdt).itemsize
    vecw = 16 // esz
    n = 32768 // esz + 3 * 256 * vecw + 5 * vecw + (1 if vecw > 1 else 0)   # >= 32 KB, ragged in every loop of the kernel
    lda = -(-n // vecw) * vecw + 2 * vecw                                    # pitch: multiple of 16 bytes, > n
    rng = np.random.default_rng(m * 31 + esz)
    A = rnd(rng, m * lda, npdt).reshape(m, lda)
    x = rnd(rng, n, npdt)
    Ad, xd = dev(A), dev(x)
    yd = torch.empty(m, dtype=xd.dtype, device="cuda")
    L.check(L.lib.b2_gemv(L.ctx(), Ad.data_ptr(), lda, m, n, xd.data_ptr(), yd.data_ptr(), 0, code, code, L.stream()))
    A64 = A[:, :n].astype(np.complex128 if np.iscomplexobj(A) else np.float64)
    ref = A64 @ x.astype(A64.dtype)
    tol = 2e-5 if dt in ("f32", "c64") else 1e-12
    scale = np.abs(A64).sum(axis=1).max() * np.abs(x).max() + 1e-30
    np.testing.assert_allclose(yd.cpu().numpy(), ref, rtol=tol, atol=tol * scale)


def test_gemv_bf16_long_rows(L):
    torch.manual_seed(2)
    for m, n in ((5, 32768), (64, 16384 + 8 * 300)):
        A = (torch.randn(m, n, device="cuda") / 180).to(torch.bfloat16)
        x = torch.randn(n, device="cuda")
        y = torch.empty(m, device="cuda")
        L.check(L.lib.b2_gemv(L.ctx(), A.data_ptr(), n, m, n, x.data_ptr(), y.data_ptr(), 0, L.BF16, L.F32, L.stream()))
        ref = A.double() @ x.double()
        assert torch.allclose(y.double(), ref, rtol=1e-4, atol=1e-4)


def test_gemv_bf16(L):
    torch.manual_seed(1)
    m, n = 1024, 2048
    A = (torch.randn(m, n, device="cuda") / 45).to(torch.bfloat16)
    for op in (0, 1):
        x = torch.randn(n if op == 0 else m, device="cuda")
        y = torch.empty(m if op == 0 else n, device="cuda")
        L.check(L.lib.b2_gemv(L.ctx(), A.data_ptr(), n, m, n, x.data_ptr(), y.data_ptr(), op, L.BF16, L.F32, L.stream()))
        A64 = A.double()
        ref = (A64 @ x.double()) if op == 0 else (A64.T @ x.double())
        assert torch.allclose(y.double(), ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("dt", list(DT))
@pytest.mark.parametrize("m,n,k", [(64, 64, 64), (37, 37, 37), (50, 40, 30), (3, 5, 4), (1, 1, 2), (2, 3, 1), (130, 70, 33)])
def test_gemm_simt(L, dt, m, n, k):
    npdt, code = DT[dt]
    rng = np.random.default_rng(m + 7 * n + 13 * k)
    tol = 3e-5 if dt in ("f32", "c64") else 1e-12
    B = rnd(rng, k * n, npdt).reshape(k, n)
    for op in (0, 1, 2):
        A = rnd(rng, m * k, npdt).reshape((m, k) if op == 0 else (k, m))
        A64 = A.astype(np.complex128 if np.iscomplexobj(A) else np.float64)
        opA = A64 if op == 0 else (A64.T if op == 1 else A64.conj().T)
        Ad, Bd = dev(A), dev(B)
        Cd = torch.ones((m, n), dtype=Ad.dtype, device="cuda")
        L.check(L.lib.b2_gemm(L.ctx(), Ad.data_ptr(), A.shape[1], Bd.data_ptr(), n, Cd.data_ptr(), n, m, n, k, op, 0,
                              code, L.stream()))
        ref = opA @ B.astype(A64.dtype)
        scale = np.abs(ref).max() + 1
        np.testing.assert_allclose(Cd.cpu().numpy(), ref, rtol=tol, atol=tol * scale)
        L.check(L.lib.b2_gemm(L.ctx(), Ad.data_ptr(), A.shape[1], Bd.data_ptr(), n, Cd.data_ptr(), n, m, n, k, op, 1,
                              code, L.stream()))
        np.testing.assert_allclose(Cd.cpu().numpy(), 2 * ref, rtol=tol, atol=2 * tol * scale)


@pytest.mark.parametrize("nz", [5, 1])
@pytest.mark.parametrize("dt", ["f32", "c64", "f64", "c128"])
def test_batched_gemm_fredholm_kat(L, nz, dt):
    # test_fredholm.py:36-95: G = arange(21*4*6) (- 1j * same), x = ones (+ 1j)
    npdt, code = DT[dt]
    cx = np.issubdtype(npdt, np.complexfloating)
    nsl, nx, ny = 21, 4, 6
    G = np.arange(nsl * nx * ny, dtype=np.float64).reshape(nsl, nx, ny)
    G = (G - 1j * G) if cx else G
    x = np.ones((nsl, ny, nz)) + (1j if cx else 0)
    Gd, xd = dev(G.astype(npdt)), dev(x.astype(npdt))
    yd = torch.empty((nsl, nx, nz), dtype=Gd.dtype, device="cuda")
    L.check(L.lib.b2_batched_gemm(L.ctx(), Gd.data_ptr(), xd.data_ptr(), yd.data_ptr(), nsl, nx, ny, nz, 0, code, L.stream()))
    ref = np.matmul(G, x)
    tol = 1e-5 if dt in ("f32", "c64") else 1e-13
    np.testing.assert_allclose(yd.cpu().numpy(), ref, rtol=tol)
    xa = torch.empty((nsl, ny, nz), dtype=Gd.dtype, device="cuda")
    L.check(L.lib.b2_batched_gemm(L.ctx(), Gd.data_ptr(), yd.data_ptr(), xa.data_ptr(), nsl, nx, ny, nz, 1, code, L.stream()))
    refa = np.matmul(G.conj().transpose(0, 2, 1), yd.cpu().numpy().astype(G.dtype))
    np.testing.assert_allclose(xa.cpu().numpy(), refa, rtol=tol * 10)


# --------------------------------------------------------------------------
# bf16 tile product on tcgen05 tensor cores
# --------------------------------------------------------------------------
@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (128, 256, 256), (256, 512, 128), (1024, 1024, 1024),
                                   (200, 264, 72), (8, 8, 8), (136, 40, 1000), (8, 16, 24), (384, 256, 64)])
@pytest.mark.parametrize("op", [0, 1])
def test_gemm_bf16_tcgen05(L, m, n, k, op):
    torch.manual_seed(m * 7 + n * 3 + k + op)
    A = (torch.randn((m, k) if op == 0 else (k, m), device="cuda") / 8).to(torch.bfloat16)
    B = (torch.randn(k, n, device="cuda") / 8).to(torch.bfloat16)
    Cm = torch.full((m, n), 7.0, device="cuda")
    lda = A.shape[1]
    L.check(L.lib.b2_gemm_bf16(L.ctx(), A.data_ptr(), lda, B.data_ptr(), n, Cm.data_ptr(), n, m, n, k, op, 0,
                               L.stream()), "b2_gemm_bf16")
    torch.cuda.synchronize()
    A64 = A.double() if op == 0 else A.double().T
    ref = A64 @ B.double()
    # fp32 accumulation of exact bf16 products: error <= ~k * eps32 * sum|a||b|
    bound = (A64.abs() @ B.double().abs()) * (k * 6e-8) + 1e-6
    err = (Cm.double() - ref).abs()
    assert bool((err <= bound).all()), f"max err {err.max().item()} at {m},{n},{k},{op}"
    # accumulate into C
    L.check(L.lib.b2_gemm_bf16(L.ctx(), A.data_ptr(), lda, B.data_ptr(), n, Cm.data_ptr(), n, m, n, k, op, 1,
                               L.stream()), "b2_gemm_bf16")
    err = (Cm.double() - 2 * ref).abs()
    assert bool((err <= 2 * bound + 1e-5).all())


def test_gemm_bf16_alignment_error_is_loud(L):
    A = torch.zeros(8, 12, device="cuda", dtype=torch.bfloat16)
    B = torch.zeros(12, 8, device="cuda", dtype=torch.bfloat16)
    Cm = torch.zeros(8, 8, device="cuda")
    rc = L.lib.b2_gemm_bf16(L.ctx(), A.data_ptr(), 12, B.data_ptr(), 8, Cm.data_ptr(), 8, 8, 8, 12, 0, 0, L.stream())
    assert rc == 2006


@pytest.mark.parametrize("variant", ["0", "1"])
def test_gemm_bf16_kernel_variants(variant):
    """both tensor-core kernels on every shape: B2_GEMM_2CTA=1 (cta_group::2 pair, the default) and
    B2_GEMM_2CTA=0 (1-CTA); the choice is latched at first use -> own process"""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "gemm2cta_worker.py")], capture_output=True, text=True,
                       timeout=240, env=dict(os.environ, B2_GEMM_2CTA=variant))
    assert r.returncode == 0 and "GEMM2CTA_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_c_abi_client(tmp_path):
    """plain-C99 program (tests/abi/abi_smoke.c) drives the host-buffer plugin entry point through the C ABI"""
    import subprocess
    from test_host_logic import _build_c_client
    exe = _build_c_client(tmp_path)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "max |err|" in res.stdout
