"""GPU tests at BASELINE.json's full sizes, through size-independent properties (the oracle is too slow there):
linearity of the JVP, adjoint identity <J v, w> = <v, J' w>, finite-difference consistency, z-slice invariance of the 3D
stencil, GMRES residual check against an independent operator application, coloured sparse Jacobian == JVP on every
seed, LU solve residual, and end-to-end Newton convergence to the reference's acceptance bound."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dp(nls, ctx, f):
    return nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))


def test_jvp_properties_3d_n100(nls, ctx):
    N = 100
    dp = _dp(nls, ctx, nls.Brusselator3D(N))
    n = dp.n
    rng = np.random.default_rng(0)
    u = dp.u0(1)
    v, w = ctx.to_device(rng.standard_normal(n)), ctx.to_device(rng.standard_normal(n))
    Jv, Jw = dp.jvp(u, v), dp.jvp(u, w)
    # linearity: J(2v - 3w) == 2 Jv - 3 Jw
    z = v.copy().scal(2.0).axpy(-3.0, w)
    Jz = dp.jvp(u, z)
    ref = Jv.copy().scal(2.0).axpy(-3.0, Jw)
    scale = ref.norm(np.inf)
    assert Jz.axpy(-1.0, ref).norm(np.inf) <= 1e-12 * scale
    # adjoint identity
    JTw = dp.vjp(u, w)
    a, b = Jv.dot(w), v.dot(JTw)
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b))
    # fused residual+JVP equals the separate kernels bit for bit
    f1, J1 = dp.residual_jvp(u, v)
    assert np.array_equal(f1.to_host(), dp.residual(u).to_host()) and np.array_equal(J1.to_host(), Jv.to_host())
    # finite-difference JVP is consistent with the exact tangent
    fd = dp.jvp(u, v, fd=True)
    assert fd.axpy(-1.0, Jv).norm(np.inf) <= 1e-4 * Jv.norm(np.inf)


def test_z_invariance_3d_n100(nls, ctx):
    # z-independent state => every k-slice of the 3D residual equals the 2D residual bit for bit
    N = 100
    d3, d2 = _dp(nls, ctx, nls.Brusselator3D(N)), _dp(nls, ctx, nls.Brusselator2D(N))
    f3 = d3.residual(d3.u0(0)).to_host()
    f2 = d2.residual(d2.u0(0)).to_host()
    for k in (0, 37, 99):
        for s in range(2):
            assert np.array_equal(f3[s * N ** 3 + k * N * N: s * N ** 3 + (k + 1) * N * N], f2[s * N * N:(s + 1) * N * N])


def test_gmres_true_residual_3d_n64(nls, ctx):
    N = 64
    dp = _dp(nls, ctx, nls.Brusselator3D(N))
    u = dp.u0(1)
    b = dp.residual(u)
    for orth, eng in (("cgs2", "multikernel"), ("cgs2", "resident"), ("mgs", "multikernel"), ("mgs", "resident")):
        gm = nls.GmresSolver(ctx, dp.n, nls.KrylovJL_GMRES(orth=orth, engine=eng, itmax=400), atol=0.0, rtol=1e-6)
        x, st = gm.solve(nls.JacobianOperator(dp, u), b)
        r = dp.jvp(u, x).axpy(-1.0, b)
        assert st.status == nls.abi.LS_SOLVED
        assert r.norm(2) <= 1.5e-6 * st.rnorm0, (orth, eng, r.norm(2) / st.rnorm0)
        assert abs(r.norm(2) - st.rnorm) <= 0.05 * st.tol + 1e-3 * st.rnorm  # Givens estimate tracks the true residual


def test_sparse_jacobian_3d_n64(nls, ctx):
    N = 64
    dp = _dp(nls, ctx, nls.Brusselator3D(N))
    sj = nls.SparseJacobian(dp)
    assert sj.nnz == 8 * dp.n and 16 <= sj.ncolors <= 24  # greedy natural-order colouring of the 3D 8-nnz pattern
    colptr, rowval, colors = sj.colptr, sj.rowval, sj.colors
    # valid distance-1 column colouring of the column intersection graph (checked on a strided sample of rows)
    import scipy.sparse as sp
    A = sp.csc_matrix((np.ones(sj.nnz), rowval - 1, colptr - 1), shape=(dp.n, dp.n)).tocsr()
    for r in range(0, dp.n, 9973):
        cols = A.indices[A.indptr[r]:A.indptr[r + 1]]
        assert len(set(colors[cols])) == len(cols)
    u = dp.u0(1)
    nz = sj.fill(u)
    rng = np.random.default_rng(1)
    x = ctx.to_device(rng.standard_normal(dp.n))
    y1, y2 = sj.mul(nz, x), dp.jvp(u, x)
    assert y1.axpy(-1.0, y2).norm(np.inf) <= 1e-11 * y2.norm(np.inf)
    t1, t2 = sj.mul(nz, x, transpose=True), dp.vjp(u, x)
    assert t1.axpy(-1.0, t2).norm(np.inf) <= 1e-11 * t2.norm(np.inf)


def test_dense_lu_n8192(nls, ctx):
    # BASELINE config 2 shape at N=64 (n = 8192): LU of the Brusselator Jacobian solves J x = f to working accuracy
    N = 64
    dp = _dp(nls, ctx, nls.Brusselator2D(N))
    n = dp.n
    u = dp.u0()
    J = dp.dense_jacobian(u)
    b = dp.residual(u)
    ipiv = ctx.empty(n, np.int64)
    info = C.c_int32(-1)
    L = nls.abi.lib()
    nls.abi.check(ctx.handle, L.b200_getrf(ctx.handle, n, J.ptr, n, ipiv.ptr, C.byref(info)))
    assert info.value == 0
    p = ipiv.to_host()
    assert p.min() >= 1 and p.max() <= n and np.all(p >= np.arange(1, n + 1))
    x = b.copy()
    nls.abi.check(ctx.handle, L.b200_getrs(ctx.handle, n, 1, J.ptr, n, ipiv.ptr, x.ptr, n))
    r = dp.jvp(u, x).axpy(-1.0, b)
    assert r.norm(np.inf) <= 1e-10 * b.norm(np.inf)


def test_newton_configs_full_size(nls, ctx):
    # config 3 at N=64 and config 4 at N=48 converge to the reference's acceptance bound (||resid||_inf < 1e-8 at abstol 1e-8)
    f = nls.Brusselator3D(64)
    u0 = _dp(nls, ctx, f).u0(1)
    sol = nls.solve(nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx), nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES()), abstol=1e-8)
    assert sol.retcode == nls.ReturnCode.Success and sol.resid_inf < 1e-8 and sol.stats.nsteps <= 5
    f = nls.Brusselator3D(48)
    u0 = _dp(nls, ctx, f).u0(1)
    fs = nls.NonlinearFunction(f, sparsity=nls.TracerSparsityDetector())
    sol2 = nls.solve(nls.NonlinearProblem(fs, u0, (3.4, 1.0, 10.0), ctx=ctx), nls.TrustRegion(linsolve=nls.KrylovJL_GMRES()), abstol=1e-8)
    assert sol2.retcode == nls.ReturnCode.Success and sol2.resid_inf < 1e-8 and sol2.stats.njacs == sol2.stats.nsteps
