"""GPU parity tests of the resident (cooperative, TMA-staged, one-basis-read-per-pass) Arnoldi engine against the
CPU oracle and against the multi-kernel engine."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(nls, ctx, po, kind, N):
    if kind == "2d":
        f, P = nls.Brusselator2D(N), po.OracleProblem.bruss2d(N)
    else:
        f, P = nls.Brusselator3D(N), po.OracleProblem.bruss3d(N)
    dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
    return f, P, dp


@pytest.mark.parametrize("kind,N,orth", [("2d", 12, "cgs2"), ("2d", 12, "mgs"), ("3d", 8, "cgs2"), ("3d", 10, "mgs"), ("2d", 40, "cgs2")])
def test_resident_gmres_vs_oracle(nls, ctx, po, kind, N, orth):
    f, P, dp = _setup(nls, ctx, po, kind, N)
    u = P.u0(1)
    b = P.residual(u)
    # the resident engine orthogonalises vector by vector (MGS); "cgs2" requests run it twice.  The oracle's CGS2 and MGS
    # runs are the matching references (same Krylov space, same stopping rule).
    ocode = po.ORTH_MGS if orth == "mgs" else po.ORTH_CGS2
    xo, so, ho = po.gmres(b, prob=P, u=u, opts=po.default_gmres_opts(atol=1e-10, rtol=1e-10, orth=ocode), want_hessenberg=400000)
    gm = nls.GmresSolver(ctx, P.n, nls.KrylovJL_GMRES(orth=orth, engine="resident", check_every=3), atol=1e-10, rtol=1e-10, keep_hessenberg=400000)
    x, st = gm.solve(nls.JacobianOperator(dp, ctx.to_device(u)), ctx.to_device(b))
    assert st.status == nls.abi.LS_SOLVED == so.status
    assert abs(st.iters - so.iters) <= 1
    assert np.abs(x.to_host() - xo).max() <= 1e-7 * np.abs(xo).max()
    k = min(st.iters, so.iters, 30)
    cnt = k * (k + 3) // 2
    hg = gm.hessenberg(st.iters)
    assert np.abs(hg[:cnt] - ho[:cnt]).max() <= 1e-8 * np.abs(ho[:cnt]).max()
    # and against the multi-kernel engine on the same inputs
    gm2 = nls.GmresSolver(ctx, P.n, nls.KrylovJL_GMRES(orth=orth, engine="multikernel"), atol=1e-10, rtol=1e-10)
    x2, st2 = gm2.solve(nls.JacobianOperator(dp, ctx.to_device(u)), ctx.to_device(b))
    assert abs(st2.iters - st.iters) <= 1 and np.abs(x.to_host() - x2.to_host()).max() <= 1e-7 * np.abs(xo).max()
    # deterministic: bit-identical on a second run
    x3, st3 = gm.solve(nls.JacobianOperator(dp, ctx.to_device(u)), ctx.to_device(b))
    assert st3.iters == st.iters and np.array_equal(x3.to_host(), x.to_host())


def test_resident_engine_eligibility(nls, ctx, po):
    # odd cell count (N = 9 -> 729 cells): TMA bulk copies need 16-byte aligned species planes -> explicit request refuses,
    # automatic selection silently uses the multi-kernel engine
    f, P, dp = _setup(nls, ctx, po, "3d", 9)
    u = P.u0(1)
    b = P.residual(u)
    gm = nls.GmresSolver(ctx, P.n, nls.KrylovJL_GMRES(engine="resident"), atol=1e-9, rtol=1e-9)
    with pytest.raises(nls.abi.B200Error) as e:
        gm.solve(nls.JacobianOperator(dp, ctx.to_device(u)), ctx.to_device(b))
    assert e.value.code == nls.abi.ERR_UNSUPPORTED
    gm2 = nls.GmresSolver(ctx, P.n, nls.KrylovJL_GMRES(engine="auto"), atol=1e-9, rtol=1e-9)
    x, st = gm2.solve(nls.JacobianOperator(dp, ctx.to_device(u)), ctx.to_device(b))
    assert st.status == nls.abi.LS_SOLVED


def test_resident_newton_vs_oracle(nls, ctx, po):
    N = 16
    f, P, dp = _setup(nls, ctx, po, "3d", N)
    u0 = P.u0(1)
    sol = nls.solve(nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx), nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(engine="resident")), abstol=1e-8)
    uo, fo, ro, tro = P.newton(u0, po.default_newton_opts(abstol=1e-8, gmres_orth=po.ORTH_CGS2))
    assert sol.retcode == ro.retcode == po.RC_SUCCESS and sol.stats.nsteps == ro.nsteps
    assert np.abs(sol.u - uo).max() <= 1e-6 * np.abs(uo).max() and np.abs(sol.resid).max() < 1e-8
    for tg, t in zip(sol.trace, tro):
        assert abs(tg.lin_iters - t.lin_iters) <= 2


@pytest.mark.parametrize("kind,N,orth", [("2d", 12, "mgs"), ("3d", 8, "cgs2"), ("2d", 40, "cgs2")])
def test_resident_gmres_on_assembled_sparse_jacobian(nls, ctx, po, kind, N, orth):
    """The resident engine with the matvec taken from the CSR view of the coloured sparse Jacobian (BASELINE config 4's
    operator): same Krylov iteration as the matrix-free operator and as the multi-kernel engine on the same matrix."""
    f, P, dp = _setup(nls, ctx, po, kind, N)
    u = P.u0(1)
    b = P.residual(u)
    sj = nls.SparseJacobian(dp)
    du = ctx.to_device(u)
    nz = sj.fill(du)
    ocode = po.ORTH_MGS if orth == "mgs" else po.ORTH_CGS2
    xo, so = po.gmres(b, prob=P, u=u, opts=po.default_gmres_opts(atol=1e-10, rtol=1e-10, orth=ocode))
    gm = nls.GmresSolver(ctx, P.n, nls.KrylovJL_GMRES(orth=orth, engine="resident"), atol=1e-10, rtol=1e-10)
    x, st = gm.solve(("sparse_jac", sj, nz), ctx.to_device(b))
    assert st.status == nls.abi.LS_SOLVED == so.status and abs(st.iters - so.iters) <= 1
    assert np.abs(x.to_host() - xo).max() <= 1e-7 * np.abs(xo).max()
    gm2 = nls.GmresSolver(ctx, P.n, nls.KrylovJL_GMRES(orth=orth, engine="multikernel"), atol=1e-10, rtol=1e-10)
    x2, st2 = gm2.solve(("sparse_jac", sj, nz), ctx.to_device(b))
    assert abs(st2.iters - st.iters) <= 1 and np.abs(x.to_host() - x2.to_host()).max() <= 1e-7 * np.abs(xo).max()
    x3, st3 = gm.solve(("sparse_jac", sj, nz), ctx.to_device(b))
    assert st3.iters == st.iters and np.array_equal(x3.to_host(), x.to_host())
