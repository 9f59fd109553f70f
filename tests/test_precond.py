"""GMRES preconditioning through the `precs` hook (`KrylovJL_GMRES(precs = (A, p) -> (Pl, Pr))`,
test/Core/core_tests__item21.jl; Krylov.jl gmres!(; M = Pl, N = Pr)).  The parity target is the pinned oracle GMRES run on
the explicitly preconditioned dense operator (M^-1 J, M^-1 b) / (J N^-1, b)."""
import numpy as np
import pytest


def _block_jacobi_inverse(J):
    """Dense inverse of the 2x2 species blocks on the diagonal of J (cell c couples rows c and NC + c)."""
    n = J.shape[0]
    NC = n // 2
    Minv = np.zeros_like(J)
    for c in range(NC):
        idx = [c, NC + c]
        Minv[np.ix_(idx, idx)] = np.linalg.inv(J[np.ix_(idx, idx)])
    return Minv


def _case(po, kind, N):
    P = po.OracleProblem.bruss2d(N) if kind == "2d" else po.OracleProblem.bruss3d(N)
    u = P.u0(1) * (1.0 + 0.05 * np.sin(np.arange(P.n)))
    return P, u, P.residual(u), P.dense_jac(u)


def test_block_jacobi_restatement_matches_dense_jacobian_blocks(po):
    P, u, b, J = _case(po, "2d", 8)
    NC = P.n // 2
    a = 10.0 * (8 - 1) ** 2
    uc, vc = u[:NC], u[NC:]
    assert np.allclose(np.diag(J)[:NC], -4 * a + 2 * uc * vc - 4.4, rtol=1e-13)      # d00 = -4a + 2uv - (A + 1)
    assert np.allclose(np.diag(J)[NC:], -4 * a - uc * uc, rtol=1e-13)               # d11
    assert np.allclose(J[np.arange(NC), NC + np.arange(NC)], uc * uc, rtol=1e-13)   # d01
    assert np.allclose(J[NC + np.arange(NC), np.arange(NC)], 3.4 - 2 * uc * vc, rtol=1e-13)  # d10
    Minv = _block_jacobi_inverse(J)
    x = np.linalg.solve(Minv @ J, Minv @ b)
    assert np.allclose(J @ x, b, rtol=1e-9, atol=1e-9 * np.abs(b).max())


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,side", [("2d", 10, "left"), ("2d", 10, "right"), ("3d", 6, "left"), ("3d", 6, "right")])
def test_gpu_preconditioned_gmres_vs_oracle(nls, ctx, po, kind, N, side):
    P, u, b, J = _case(po, kind, N)
    Minv = _block_jacobi_inverse(J)
    f = nls.Brusselator2D(N) if kind == "2d" else nls.Brusselator3D(N)
    dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
    du, db = ctx.to_device(u), ctx.to_device(b)
    # the preconditioner itself
    pre = nls.BlockJacobi(side)
    gm = nls.GmresSolver(ctx, P.n, nls.KrylovJL_GMRES(orth="cgs2"), atol=1e-11, rtol=1e-11, keep_hessenberg=100000)
    opts = po.default_gmres_opts(atol=1e-11, rtol=1e-11, orth=po.ORTH_CGS2)
    if side == "left":
        xo, so, ho = po.gmres(Minv @ b, dense=Minv @ J, opts=opts, want_hessenberg=100000)
        x, st = gm.solve(nls.JacobianOperator(dp, du), db, Pl=pre.linop(dp, du))
    else:
        yo, so, ho = po.gmres(b, dense=J @ Minv, opts=opts, want_hessenberg=100000)
        xo = Minv @ yo
        x, st = gm.solve(nls.JacobianOperator(dp, du), db, Pr=pre.linop(dp, du))
    assert st.status == nls.abi.LS_SOLVED == so.status and abs(st.iters - so.iters) <= 1
    assert abs(st.rnorm0 - so.rnorm0) <= 1e-10 * so.rnorm0        # the stopping test sees the preconditioned residual
    assert np.abs(x.to_host() - xo).max() <= 1e-7 * np.abs(xo).max()
    k = min(st.iters, so.iters, 20)
    cnt = k * (k + 3) // 2
    hg = gm.hessenberg(st.iters)
    assert np.abs(hg[:cnt] - ho[:cnt]).max() <= 1e-8 * np.abs(ho[:cnt]).max()
    assert np.abs(J @ x.to_host() - b).max() <= 1e-7 * np.abs(b).max()
    # a later unpreconditioned solve on the same cache is unaffected
    x2, st2 = gm.solve(nls.JacobianOperator(dp, du), db)
    xo2, so2 = po.gmres(b, prob=P, u=u, opts=opts)
    assert abs(st2.iters - so2.iters) <= 1 and np.abs(x2.to_host() - xo2).max() <= 1e-7 * np.abs(xo2).max()
    # the resident engine refuses preconditioners explicitly
    gm3 = nls.GmresSolver(ctx, P.n, nls.KrylovJL_GMRES(engine="resident"), atol=1e-9, rtol=1e-9)
    with pytest.raises(nls.abi.B200Error) as e:
        gm3.solve(nls.JacobianOperator(dp, du), db, Pl=pre.linop(dp, du))
    assert e.value.code == nls.abi.ERR_UNSUPPORTED


@pytest.mark.gpu
@pytest.mark.parametrize("side", ["left", "right"])
def test_gpu_newton_with_precs(nls, ctx, po, side):
    N = 16
    P = po.OracleProblem.bruss2d(N)
    u0 = P.u0()
    prob = nls.NonlinearProblem(nls.Brusselator2D(N), u0, (3.4, 1.0, 10.0), ctx=ctx)
    ref = nls.solve(prob, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES()), abstol=1e-8)
    # Left preconditioning makes Krylov.jl test the PRECONDITIONED residual (here ~1/(4a) = 1e-4 times the true one), so the
    # inherited absolute tolerance (abstol = 1e-8, solve.jl:203) would stop the linear solves far too early — the same
    # happens in the reference; a left-preconditioned run therefore sets the Krylov tolerances itself.
    kw = dict(atol=1e-13, rtol=1e-9) if side == "left" else {}
    sol = nls.solve(prob, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(precs=nls.BlockJacobi(side), **kw)), abstol=1e-8)
    assert sol.retcode == ref.retcode == nls.ReturnCode.Success and sol.stats.nsteps <= ref.stats.nsteps + 1
    assert np.abs(sol.resid).max() < 1e-8 and np.abs(sol.u - ref.u).max() <= 1e-6 * np.abs(ref.u).max()
    uo, fo, ro, _ = P.newton(u0, po.default_newton_opts(abstol=1e-8, linsolve=po.LINSOLVE_GMRES, gmres_orth=po.ORTH_CGS2))
    assert np.abs(sol.u - uo).max() <= 1e-6 * np.abs(uo).max()
    # sparse concrete J + TrustRegion takes the same hook
    fs = nls.NonlinearFunction(nls.Brusselator2D(N), sparsity=nls.TracerSparsityDetector())
    sol2 = nls.solve(nls.NonlinearProblem(fs, u0, (3.4, 1.0, 10.0), ctx=ctx), nls.TrustRegion(linsolve=nls.KrylovJL_GMRES(precs=nls.BlockJacobi(side), **kw)), abstol=1e-8)
    assert sol2.retcode == nls.ReturnCode.Success and np.abs(sol2.u - uo).max() <= 1e-6 * np.abs(uo).max()
