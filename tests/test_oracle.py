"""CPU tests: the oracle (oracle/oracle.c) against the golden fixtures, the reference's own acceptance bounds and
independent NumPy/SciPy computations.  These pin the checker before it is trusted by the GPU parity tests."""
import numpy as np
import pytest
import scipy.sparse as sp


def test_residual_u0_golden(po, golden):
    for N in (8, 32):
        P = po.OracleProblem.bruss2d(N)
        u0 = P.u0()
        assert np.array_equal(u0, golden["u0_%d" % N]) or np.allclose(u0, golden["u0_%d" % N], rtol=1e-15, atol=0)
        f0 = P.residual(golden["u0_%d" % N])
        assert np.allclose(f0, golden["f0_%d" % N], rtol=1e-13, atol=1e-9)
    assert abs(np.abs(po.OracleProblem.bruss2d(32).residual(golden["u0_32"])).max() - 1440.481446976741) < 1e-9


def test_jvp_vjp_golden(po, golden):
    # SciMLJacobianOperators core_tests__item3.jl:40-58 checks JVP/VJP against the analytic Jacobian at atol 1e-5
    for N in (8, 32):
        P = po.OracleProblem.bruss2d(N)
        u0, v = golden["u0_%d" % N], golden["v_%d" % N]
        sc = np.abs(golden["Jv_%d" % N]).max()
        assert np.abs(P.jvp(u0, v) - golden["Jv_%d" % N]).max() < 1e-12 * sc
        assert np.abs(P.vjp(u0, v) - golden["JTv_%d" % N]).max() < 1e-12 * sc
        fd = P.jvp_fd(u0, v)
        assert np.abs(fd - golden["Jv_%d" % N]).max() < 1e-5 * sc


def test_pattern_and_colors_golden(po, golden):
    for N in (8, 32):
        P = po.OracleProblem.bruss2d(N)
        colptr, rowval = P.pattern(1)
        assert np.array_equal(colptr, golden["colptr_%d" % N])
        assert np.array_equal(rowval, golden["rowval_%d" % N])
        colors, nc = po.coloring_column(P.n, colptr, rowval, 1)
        assert np.array_equal(colors, golden["colors_%d" % N])
        nz = P.sparse_jac(golden["u0_%d" % N], colptr, rowval, colors, nc)
        assert np.allclose(nz, golden["nzval_%d" % N], rtol=1e-13, atol=1e-9)
    assert golden["colors_32"].max() == 12 and list(golden["colors_32"][:6]) == [1, 2, 3, 1, 2, 3]


def test_dense_jac_matches_sparse(po, golden):
    P = po.OracleProblem.bruss2d(8)
    J = P.dense_jac(golden["u0_8"])
    Jg = sp.csc_matrix((golden["nzval_8"], golden["rowval_8"] - 1, golden["colptr_8"] - 1), shape=(P.n, P.n)).toarray()
    assert np.allclose(J, Jg, rtol=1e-13, atol=1e-9)


@pytest.mark.parametrize("linsolve", ["dense", "gmres", "sparse"])
def test_newton_brusselator_reference_bound(po, golden, linsolve):
    # sparsity_tests__item1.jl:54-93: solve(prob, NewtonRaphson(); abstol = 1e-8) -> norm(sol.resid, Inf) < 1e-8
    P = po.OracleProblem.bruss2d(32)
    ls = {"dense": po.LINSOLVE_DENSE_LU, "gmres": po.LINSOLVE_GMRES, "sparse": po.LINSOLVE_SPARSE_GMRES}[linsolve]
    u, fu, res, tr = P.newton(golden["u0_32"], po.default_newton_opts(abstol=1e-8, linsolve=ls))
    assert res.retcode == po.RC_SUCCESS
    assert np.abs(fu).max() < 1e-8 and res.resid_inf < 1e-8
    assert np.abs(u - golden["root_32"]).max() / np.abs(golden["root_32"]).max() < 1e-9
    if linsolve == "dense":
        hist = [t.fnorm_inf for t in tr]
        assert np.allclose(hist, golden["newton_hist_32"][1:], rtol=1e-5)
        assert (res.nsteps, res.nf, res.njacs, res.nfactors, res.nsolve) == (3, 3, 4, 3, 3)


def test_quadratic_sqrt2(po, golden):
    # rootfind_tests__item1.jl / common_rootfind_testing.jl: u.*u .- 2, u0 = ones -> sqrt(2), err < 1e-9
    P = po.OracleProblem.quadratic(1000, 2.0)
    for ls in (po.LINSOLVE_GMRES, po.LINSOLVE_DENSE_LU, po.LINSOLVE_SPARSE_GMRES):
        u, fu, res, tr = P.newton(np.ones(1000), po.default_newton_opts(linsolve=ls))
        assert res.retcode == po.RC_SUCCESS
        assert np.abs(u - np.sqrt(2)).max() < 1e-9 and np.abs(fu).max() < 1e-9
        assert res.nsteps == 5
    it = golden["quadratic_iterates"]
    assert abs(it[1] - 1.5) < 1e-15 and abs(it[2] - 17 / 12) < 1e-15


@pytest.mark.parametrize("tr", [False, True])
def test_tridiag_jfnk(po, golden, tr):
    # rootfind_tests__item20.jl:32-54: custom jvp + KrylovJL_GMRES, abstol = 1e-13 -> maximum(abs, resid) < 1e-6
    p = golden["tridiag_p"]
    P = po.OracleProblem.tridiag_quad(p)
    o = po.default_newton_opts(abstol=1e-13, linsolve=po.LINSOLVE_GMRES, globalization=po.GLOB_TRUST_REGION if tr else po.GLOB_NONE)
    u, fu, res, _ = P.newton(p, o)
    assert np.abs(fu).max() < 1e-6
    assert np.abs(u - golden["tridiag_root"]).max() < 1e-9


def test_gmres_against_direct(po, golden):
    P = po.OracleProblem.bruss2d(8)
    u0 = golden["u0_8"]
    J = P.dense_jac(u0)
    b = P.residual(u0)
    xref = np.linalg.solve(J, b)
    for orth in (po.ORTH_MGS, po.ORTH_CGS2):
        x, st = po.gmres(b, prob=P, u=u0, opts=po.default_gmres_opts(atol=0.0, rtol=1e-12, orth=orth))
        assert st.status == po.LS_SOLVED
        assert np.abs(x - xref).max() / np.abs(xref).max() < 1e-8
        assert abs(np.linalg.norm(b - J @ x) - st.rnorm) < 1e-8 * st.rnorm0 + 1e-12
    # restarted GMRES converges to the same solution; dense / CSC operators agree with the matrix-free one
    x2, st2 = po.gmres(b, dense=J, opts=po.default_gmres_opts(atol=0.0, rtol=1e-12, restart=40, itmax=4000))
    assert st2.status == po.LS_SOLVED and st2.restarts > 0
    assert np.abs(x2 - xref).max() / np.abs(xref).max() < 1e-7
    cp, rv = P.pattern(1)
    col, nc = po.coloring_column(P.n, cp, rv)
    nz = P.sparse_jac(u0, cp, rv, col, nc)
    x3, st3 = po.gmres(b, csc=(cp, rv, nz, 1), opts=po.default_gmres_opts(atol=0.0, rtol=1e-12))
    assert np.abs(x3 - xref).max() / np.abs(xref).max() < 1e-8
    assert np.allclose(po.spmv(P.n, cp, rv, nz, golden["v_8"]), J @ golden["v_8"], rtol=1e-12, atol=1e-8)
    assert np.allclose(po.spmv(P.n, cp, rv, nz, golden["v_8"], transpose=True), J.T @ golden["v_8"], rtol=1e-12, atol=1e-8)


def test_lu_against_lapack(po):
    rng = np.random.default_rng(1)
    for n in (5, 64, 200):
        A = rng.standard_normal((n, n))
        b = rng.standard_normal(n)
        LU, ipiv, info = po.getrf(A)
        assert info == 0
        x = po.getrs(LU, ipiv, b)
        assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-9, atol=1e-9)
        import scipy.linalg as sl
        lu2, piv2 = sl.lu_factor(A)
        assert np.array_equal(ipiv - 1, piv2)
        assert np.allclose(LU, lu2, rtol=1e-10, atol=1e-12)
    LU, ipiv, info = po.getrf(np.zeros((3, 3)))
    assert info == 1  # LAPACK: first exactly-zero pivot, 1-based


def test_3d_reduces_to_2d(po):
    # SURVEY.md §A.2 invariant: z-independent data => every k-slice of the 3D residual / JVP equals the 2D one exactly
    N = 10
    P2, P3 = po.OracleProblem.bruss2d(N), po.OracleProblem.bruss3d(N)
    u2 = P2.u0()
    u3 = P3.u0(0)
    f2, f3 = P2.residual(u2), P3.residual(u3)
    rng = np.random.default_rng(3)
    v2 = rng.standard_normal(P2.n)
    v3 = np.concatenate([np.tile(v2[:N * N], N), np.tile(v2[N * N:], N)])
    J2, J3 = P2.jvp(u2, v2), P3.jvp(u3, v3)
    for k in range(N):
        for s in range(2):
            sl3 = slice(s * N ** 3 + k * N * N, s * N ** 3 + (k + 1) * N * N)
            sl2 = slice(s * N * N, (s + 1) * N * N)
            assert np.array_equal(f3[sl3], f2[sl2])
            assert np.array_equal(J3[sl3], J2[sl2])
    cp, rv = P3.pattern(1)
    assert len(rv) == 8 * P3.n
    col, nc = po.coloring_column(P3.n, cp, rv)
    assert nc == 17
    # a colouring is valid iff no two columns sharing a row share a colour
    A = sp.csc_matrix((np.ones(len(rv)), rv - 1, cp - 1), shape=(P3.n, P3.n)).tocsr()
    for r in range(0, P3.n, 37):
        cols = A.indices[A.indptr[r]:A.indptr[r + 1]]
        assert len(set(col[cols])) == len(cols)


def test_trust_region_and_forcing(po, golden):
    P = po.OracleProblem.bruss2d(32)
    u0 = golden["u0_32"]
    for kw in (dict(globalization=po.GLOB_TRUST_REGION), dict(forcing=po.FORCING_EW2), dict(jvp_mode=po.JVP_FINITE_DIFF)):
        u, fu, res, tr = P.newton(u0, po.default_newton_opts(abstol=1e-8, linsolve=po.LINSOLVE_GMRES, gmres_orth=po.ORTH_CGS2, **kw))
        assert res.retcode == po.RC_SUCCESS and np.abs(fu).max() < 1e-8
        assert np.abs(u - golden["root_32"]).max() / np.abs(golden["root_32"]).max() < 1e-6
    # a tiny initial radius forces Cauchy / dogleg steps and radius growth
    u, fu, res, tr = P.newton(u0, po.default_newton_opts(abstol=1e-8, linsolve=po.LINSOLVE_GMRES, gmres_orth=po.ORTH_CGS2,
                                                         globalization=po.GLOB_TRUST_REGION, tr_initial_trust_radius=1.0))
    assert res.retcode == po.RC_SUCCESS and np.abs(fu).max() < 1e-8 and res.nsteps > 3
    assert tr[0].trust_radius == 2.0


def test_termination_modes(po):
    P = po.OracleProblem.quadratic(10, 2.0)
    u, fu, res, _ = P.newton(np.ones(10), po.default_newton_opts(maxiters=2, linsolve=po.LINSOLVE_GMRES))
    assert res.retcode == po.RC_MAXITERS and res.nsteps == 2
    u, fu, res, _ = P.newton(np.ones(10), po.default_newton_opts(termination=po.TERM_ABS_NORM, abstol=1e-6, linsolve=po.LINSOLVE_GMRES))
    assert res.retcode == po.RC_SUCCESS and res.resid_inf <= 1e-6
    # zero-residual start still takes one step (termination is checked after the update)
    u, fu, res, _ = P.newton(np.full(10, np.sqrt(2.0)), po.default_newton_opts(abstol=1e-9, linsolve=po.LINSOLVE_GMRES))
    assert res.nsteps == 1 and res.retcode == po.RC_SUCCESS
    # negative p has no real root: Newton from u0 = 1 hits a zero derivative -> non-finite -> Unstable or stall, never Success
    Pn = po.OracleProblem.quadratic(4, -1.0)
    u, fu, res, _ = Pn.newton(np.ones(4), po.default_newton_opts(maxiters=50, linsolve=po.LINSOLVE_DENSE_LU))
    assert res.retcode != po.RC_SUCCESS


def test_ensemble(po):
    # core_tests__item6.jl:14-20: every trajectory succeeds; here each has its own (A, B)
    N, K = 8, 12
    P = po.OracleProblem.bruss2d(N)
    u0 = np.tile(P.u0(), (K, 1))
    A = 3.4 + 0.1 * (np.arange(K) % 64) / 64
    B = 1.0 + 0.05 * (np.arange(K) // 4) / 128
    u, resid, rc, ns, nj, res = po.ensemble_solve(N, u0, A, B, opts=po.default_newton_opts(abstol=1e-8, gmres_orth=po.ORTH_CGS2))
    assert res.nsuccess == K and np.all(rc == po.RC_SUCCESS) and resid.max() < 1e-8
    Pm = po.OracleProblem.bruss2d(N, A=A[5], B=B[5])
    um, _, rm, _ = Pm.newton(P.u0(), po.default_newton_opts(abstol=1e-8, gmres_orth=po.ORTH_CGS2))
    assert np.array_equal(um, u[5]) and rm.nsteps == ns[5]
