"""GPU tests of the round-2 additions: the multigrid preconditioner (`precs`), the sparse direct route
(`NewtonRaphson()` on a sparse prototype -> KLU/UMFPACK in the reference; here RCM + banded LU), `maxtime`, and the
host-callback polling fix."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dp(nls, ctx, f):
    return nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))


def _h(x):
    return x.to_host() if hasattr(x, "to_host") else np.asarray(x)


def _apply(nls, ctx, op, x, n):
    y = ctx.zeros(n)
    nls.abi.check(ctx.handle, nls.abi.lib().b200_linop_apply(op, x.ptr, y.ptr))
    return y


# ------------------------------------------------------------------------------------------------ multigrid
@pytest.mark.parametrize("dim,N", [(2, 12), (2, 10), (2, 7), (2, 30), (3, 6), (3, 10), (3, 12)])
def test_multigrid_vcycle_vs_numpy_restatement(nls, ctx, po, dim, N):
    from oracle import mg_numpy as mgn
    P = po.OracleProblem.bruss2d(N) if dim == 2 else po.OracleProblem.bruss3d(N)
    u = P.u0(1) * (1.0 + 0.05 * np.sin(np.arange(P.n)))
    mg = mgn.Multigrid(N, dim, u)
    dp = _dp(nls, ctx, nls.Brusselator2D(N) if dim == 2 else nls.Brusselator3D(N))
    du = ctx.to_device(u)
    op = nls.Multigrid("right").linop(dp, du)
    rng = np.random.default_rng(N)
    for _ in range(2):
        x = rng.standard_normal(P.n)
        y = _apply(nls, ctx, op, ctx.to_device(x), P.n).to_host()
        ref = mg.vcycle(x)
        assert np.abs(y - ref).max() <= 1e-11 * np.abs(ref).max(), (dim, N, mg.sizes())
    nls.abi.lib().b200_linop_destroy(op)


@pytest.mark.parametrize("side", ["left", "right"])
def test_multigrid_preconditioned_gmres_vs_oracle(nls, ctx, po, side):
    from oracle import mg_numpy as mgn
    N = 12
    P = po.OracleProblem.bruss2d(N)
    u = P.u0(0) * (1.0 + 0.05 * np.sin(np.arange(P.n)))
    b, J = P.residual(u), P.dense_jac(u)
    Minv = mgn.Multigrid(N, 2, u).dense_inverse()
    dp = _dp(nls, ctx, nls.Brusselator2D(N))
    du, db = ctx.to_device(u), ctx.to_device(b)
    pre = nls.Multigrid(side)
    gm = nls.GmresSolver(ctx, P.n, nls.KrylovJL_GMRES(orth="cgs2"), atol=1e-11, rtol=1e-11, keep_hessenberg=10000)
    opts = po.default_gmres_opts(atol=1e-11, rtol=1e-11, orth=po.ORTH_CGS2)
    if side == "left":
        xo, so, ho = po.gmres(Minv @ b, dense=Minv @ J, opts=opts, want_hessenberg=10000)
        x, st = gm.solve(nls.JacobianOperator(dp, du), db, Pl=pre.linop(dp, du))
    else:
        yo, so, ho = po.gmres(b, dense=J @ Minv, opts=opts, want_hessenberg=10000)
        xo = Minv @ yo
        x, st = gm.solve(nls.JacobianOperator(dp, du), db, Pr=pre.linop(dp, du))
    assert st.status == nls.abi.LS_SOLVED == so.status and abs(st.iters - so.iters) <= 1
    assert st.iters <= 16                                     # unpreconditioned: 54
    assert np.abs(x.to_host() - xo).max() <= 1e-7 * np.abs(xo).max()
    k = min(st.iters, so.iters, 8)
    cnt = k * (k + 3) // 2
    assert np.abs(gm.hessenberg(st.iters)[:cnt] - ho[:cnt]).max() <= 1e-8 * np.abs(ho[:cnt]).max()
    assert np.abs(J @ x.to_host() - b).max() <= 1e-7 * np.abs(b).max()


def test_newton_with_multigrid_cuts_the_krylov_iterations(nls, ctx):
    N = 40
    f = nls.Brusselator3D(N)
    u0 = _dp(nls, ctx, f).u0(1)
    prob = nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx)
    ref = nls.solve(prob, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(orth="mgs")), abstol=1e-8)
    sol = nls.solve(prob, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(orth="mgs", precs=nls.Multigrid("right"))), abstol=1e-8)
    assert sol.retcode == ref.retcode == nls.ReturnCode.Success and sol.stats.nsteps == ref.stats.nsteps
    assert sol.resid_inf < 1e-8 and np.abs(_h(sol.u) - _h(ref.u)).max() <= 1e-6 * np.abs(_h(ref.u)).max()
    li, lr = [t.lin_iters for t in sol.trace], [t.lin_iters for t in ref.trace]
    assert max(li) <= 40 and min(lr) >= 5 * max(li), (li, lr)
    # with Eisenstat-Walker forcing on top
    sol2 = nls.solve(prob, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(orth="mgs", precs=nls.Multigrid("right")), forcing=nls.EisenstatWalkerForcing2()), abstol=1e-8)
    assert sol2.retcode == nls.ReturnCode.Success and sol2.resid_inf < 1e-8 and sol2.stats.njvp <= sol.stats.njvp


# ------------------------------------------------------------------------------------------------ sparse direct route
@pytest.mark.parametrize("dim,N", [(2, 16), (2, 33), (3, 8)])
def test_sparse_band_lu_vs_scipy(nls, ctx, dim, N):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    dp = _dp(nls, ctx, nls.Brusselator2D(N) if dim == 2 else nls.Brusselator3D(N))
    sj = nls.SparseJacobian(dp)
    u = dp.u0(1 if dim == 3 else 0)
    nz = sj.fill(u)
    A = sp.csc_matrix((nz.to_host(), sj.rowval - 1, sj.colptr - 1), shape=(dp.n, dp.n))
    lu = nls.SparseBandLU(ctx, dp.n, sj.colptr, sj.rowval, 1)
    kl, ku = lu.bandwidth()
    per_line = N if dim == 2 else N * N
    assert 0 < kl <= 8 * per_line and 0 < ku <= 8 * per_line      # RCM keeps the periodic stencil inside a band of a few grid lines / planes
    assert lu.factor(nz) == 0
    rng = np.random.default_rng(0)
    ref_lu = spla.splu(A)
    for _ in range(2):
        b = rng.standard_normal(dp.n)
        x = lu.solve(ctx.to_device(b)).to_host()
        xr = ref_lu.solve(b)
        assert np.abs(x - xr).max() <= 1e-9 * np.abs(xr).max()
        assert np.abs(A @ x - b).max() <= 1e-9 * np.abs(A).max() * np.abs(x).max()
    # structurally fine but numerically singular: a zeroed column is reported like LAPACK's info > 0
    bad = nz.to_host().copy()
    bad[sj.colptr[5] - 1:sj.colptr[6] - 1] = 0.0
    assert lu.factor(ctx.to_device(bad)) > 0


def test_newton_on_sparse_prototype_takes_the_direct_route(nls, ctx, po):
    """sparsity_tests__item1.jl:54-93: NewtonRaphson() with a sparse jac_prototype / detector, ||resid||_inf < 1e-8 at abstol 1e-8."""
    N = 32
    P = po.OracleProblem.bruss2d(N)
    u0 = P.u0()
    fs = nls.NonlinearFunction(nls.Brusselator2D(N), sparsity=nls.TracerSparsityDetector())
    sol = nls.solve(nls.NonlinearProblem(fs, u0, (3.4, 1.0, 10.0), ctx=ctx), nls.NewtonRaphson(), abstol=1e-8)
    dense = nls.solve(nls.NonlinearProblem(nls.Brusselator2D(N), u0, (3.4, 1.0, 10.0), ctx=ctx), nls.NewtonRaphson(), abstol=1e-8)
    assert sol.retcode == nls.ReturnCode.Success and np.abs(sol.resid).max() < 1e-8
    s = sol.stats
    # one coloured Jacobian + one factorisation + one solve per step; a sparse prototype gives no extra Jacobian at init
    assert (s.nsteps, s.nf, s.njacs, s.nfactors, s.nsolve) == (3, 3, 3, 3, 3) and s.njvp == 0
    assert np.abs(_h(sol.u) - _h(dense.u)).max() <= 1e-9 * np.abs(_h(dense.u)).max()
    fn = [t.fnorm_inf for t in sol.trace]
    assert abs(fn[0] - 18.00640584541683) <= 1e-9 * 18.0 and abs(fn[1] - 5.434051731e-4) <= 1e-9      # SURVEY.md §A.4 probe values
    sol_k = nls.solve(nls.NonlinearProblem(fs, u0, (3.4, 1.0, 10.0), ctx=ctx), nls.NewtonRaphson(linsolve=nls.KLUFactorization()), abstol=1e-8)
    assert np.array_equal(_h(sol_k.u), _h(sol.u))
    tr = nls.solve(nls.NonlinearProblem(fs, u0, (3.4, 1.0, 10.0), ctx=ctx), nls.TrustRegion(), abstol=1e-8)
    assert tr.retcode == nls.ReturnCode.Success and np.abs(_h(tr.u) - _h(dense.u)).max() <= 1e-8 * np.abs(_h(dense.u)).max()


def test_sparse_direct_route_refuses_a_band_that_cannot_fit(nls, ctx):
    f = nls.NonlinearFunction(nls.Brusselator3D(64), sparsity=nls.TracerSparsityDetector())
    dp = _dp(nls, ctx, nls.Brusselator3D(64))
    with pytest.raises(nls.abi.B200Error) as e:
        nls.init(nls.NonlinearProblem(f, dp.u0(1), (3.4, 1.0, 10.0), ctx=ctx), nls.NewtonRaphson(), abstol=1e-8)
    assert e.value.code == nls.abi.ERR_NOMEM and "KrylovJL_GMRES" in str(e.value)


# ------------------------------------------------------------------------------------------------ maxtime, callbacks
def test_maxtime_stops_the_solve(nls, ctx):
    f = nls.Brusselator3D(48)
    u0 = _dp(nls, ctx, f).u0(1)
    sol = nls.solve(nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx), nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES()), abstol=1e-8, maxtime=1e-4)
    assert sol.retcode == nls.ReturnCode.MaxTime and sol.stats.nsteps == 1      # checked after the step that crossed the limit (solve.jl:847-855)
    sol = nls.solve(nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx), nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES()), abstol=1e-8, maxtime=3600.0)
    assert sol.retcode == nls.ReturnCode.Success


def test_callback_operator_is_never_called_after_convergence(nls, ctx, po):
    """the status word is polled every iteration for host-callback operators, so user code never sees the
    operand of an iteration that a finished solve skipped (and is not called at all when ||r0|| <= tol)."""
    N = 8
    P = po.OracleProblem.bruss2d(N)
    u = P.u0(0)
    dp = _dp(nls, ctx, nls.Brusselator2D(N))
    du = ctx.to_device(u)
    calls = []

    def mv(y, x):
        xh = x.to_host()
        assert np.all(np.isfinite(xh)) and abs(np.linalg.norm(xh) - 1.0) < 1e-10     # always a normalised Krylov vector
        calls.append(1)
        y.copy_from_host(P.jvp(u, xh))
    b = P.residual(u)
    gm = nls.GmresSolver(ctx, P.n, nls.KrylovJL_GMRES(orth="cgs2", check_every=8), atol=1e-9, rtol=1e-9)
    x, st = gm.solve(mv, ctx.to_device(b))
    assert st.status == nls.abi.LS_SOLVED and len(calls) == st.iters
    calls.clear()
    x, st = gm.solve(mv, ctx.to_device(np.zeros(P.n)))
    assert st.status == nls.abi.LS_SOLVED and st.iters == 0 and len(calls) == 0


def test_singular_lu_is_rescued_by_pivoted_qr(nls, ctx):
    """linear_solve.jl:48-55: a failed LU falls back to a column-pivoted QR instead of failing the step.  f = u.^2 .- 2 with one
    component started at 0: J = diag(2u) is exactly singular there, the QR's basic solution leaves that component alone and
    the others converge — the solve ends on maxiters, not on InternalLinearSolveFailed."""
    n = 12
    u0 = np.ones(n)
    u0[7] = 0.0
    # AbsNorm mode: no best-iterate rollback (||f||_inf stays 2 at the stuck component, so a Best mode would hand back u0)
    sol = nls.solve(nls.NonlinearProblem(nls.QuadraticFunction(n), u0, 2.0, ctx=ctx), nls.NewtonRaphson(), abstol=1e-10, maxiters=12,
                    termination_condition=nls.AbsNormTerminationMode())
    assert sol.retcode == nls.ReturnCode.MaxIters and sol.stats.nsteps == 12
    u = _h(sol.u)
    keep = np.arange(n) != 7
    assert np.abs(u[keep] - np.sqrt(2.0)).max() < 1e-12 and u[7] == 0.0


def test_ensemble_trajectories_that_outgrow_the_basis_slab_are_not_truncated(nls, ctx, po, monkeypatch):
    """the batched ensemble kernel keeps a fixed number of Krylov columns per CTA.  A trajectory whose linear
    solve needs more is redone by the general Newton driver (same algorithm, growing basis) instead of stopping GMRES early: with
    the slab shrunk to 24 columns every trajectory takes that route and the results still match the per-trajectory oracle."""
    import bench
    N, K = 16, 12
    P = po.OracleProblem.bruss2d(N)
    A, B = bench.ensemble_params(K)
    u0 = np.tile(P.u0(), (K, 1)) * (1.0 + 0.01 * np.arange(K))[:, None]
    alg = nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(orth="mgs"))
    uo, ro, rco, nso, njo, reso = po.ensemble_solve(N, u0, A, B, opts=po.default_newton_opts(abstol=1e-8, gmres_orth=po.ORTH_MGS))
    assert njo.max() / nso.max() > 24          # the linear solves do need more than 24 columns
    full = nls.EnsembleCache(ctx, N, K, 10.0, alg, abstol=1e-8)
    rf = full.solve(ctx.to_device(u0.ravel()), ctx.to_device(A), ctx.to_device(B))
    monkeypatch.setenv("B200_ENS_BASIS_COLUMNS", "24")
    small = nls.EnsembleCache(ctx, N, K, 10.0, alg, abstol=1e-8)
    rs = small.solve(ctx.to_device(u0.ravel()), ctx.to_device(A), ctx.to_device(B))
    monkeypatch.delenv("B200_ENS_BASIS_COLUMNS")
    assert rs.nsuccess == rf.nsuccess == K == reso.nsuccess
    assert np.array_equal(small.rc.to_host(), rco) and np.array_equal(small.ns.to_host(), nso)
    us = small.u_out.to_host().reshape(K, -1)
    assert np.abs(us - uo).max() <= 1e-6 * np.abs(uo).max()
    assert np.all(np.abs(small.nj.to_host() - njo) <= 2 * nso)
    assert small.resid.to_host().max() < 1e-8


def test_limited_memory_broyden_at_scale_vs_numpy(nls, ctx):
    """LimitedMemoryBroyden on 2*10^5 unknowns (the tall-skinny products with the n x threshold factors; threshold 4, so the circular
    buffer wraps) against the NumPy restatement: same number of steps, residual history and root."""
    from oracle import newton_numpy as nn
    n = 200_000
    u0 = np.linspace(0.6, 3.0, n)
    # abstol 1e-7: the solve ends (9 steps) before any component of du reaches the reset tolerance eps^(3/4), whose <= test is
    # decided by the last bit
    ref = nn.solve_broyden(nn.Quadratic(n, 2.0), u0, init_jacobian="low_rank", threshold=4, max_resets=3, termination=nn.Termination(abstol=1e-7), maxiters=100)
    sol = nls.solve(nls.NonlinearProblem(nls.QuadraticFunction(n), u0, 2.0, ctx=ctx), nls.LimitedMemoryBroyden(threshold=4), abstol=1e-7, maxiters=100)
    assert sol.retcode == ref["retcode"] == nls.ReturnCode.Success
    assert sol.stats.nsteps == ref["nsteps"] and sol.stats.nf == ref["nf"]
    fn = np.array([t.fnorm_inf for t in sol.trace])
    big = np.array(ref["fnorm_inf"]) > 1e-6 * max(ref["fnorm_inf"])
    assert np.allclose(fn[big], np.array(ref["fnorm_inf"])[big], rtol=1e-7)
    u = sol.u.to_host() if hasattr(sol.u, "to_host") else np.asarray(sol.u)
    assert ref["nsteps"] == 9 and ref["nresets"] == 0
    assert np.abs(u - np.sqrt(2.0)).max() < 1e-7 and np.abs(u - ref["u"]).max() < 1e-10


def test_klement_diagonal_at_scale_vs_numpy(nls, ctx):
    """Klement() (default diagonal structure: an n-vector Jacobian, elementwise kernels) on 5*10^5 unknowns against the NumPy restatement.
    Two families of components (starts 1.0 and 1.8) so that no component converges much earlier than the others: the rule divides a
    numerator that cancels to rounding level by J^2 du^2 once a component is nearly converged, and from there its iterates depend on
    the last bit of the residual (a spread of starting values makes even the NumPy run jump, see make_control_golden.py)."""
    from oracle import newton_numpy as nn
    n = 500_000
    u0 = np.where(np.arange(n) % 2 == 0, 1.0, 1.8)
    ref = nn.solve_klement(nn.Quadratic(n, 2.0), u0, termination=nn.Termination(abstol=1e-6), maxiters=100)
    sol = nls.solve(nls.NonlinearProblem(nls.QuadraticFunction(n), u0, 2.0, ctx=ctx), nls.Klement(), abstol=1e-6, maxiters=100)
    assert sol.retcode == ref["retcode"] == nls.ReturnCode.Success
    s = sol.stats
    assert (s.nsteps, s.nf, s.nsolve, s.nfactors, s.njacs) == (ref["nsteps"], ref["nf"], ref["nsolve"], ref["nfactors"], ref["njacs"]) and ref["nsteps"] == 6
    fn = np.array([t.fnorm_inf for t in sol.trace])
    assert np.allclose(fn[:5], np.array(ref["fnorm_inf"])[:5], rtol=1e-8)
    u = sol.u.to_host() if hasattr(sol.u, "to_host") else np.asarray(sol.u)
    assert np.abs(u - np.sqrt(2.0)).max() < 1e-6 and sol.resid_inf <= 1e-6
