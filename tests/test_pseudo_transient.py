"""PseudoTransient = DampedNewtonDescent + SwitchedEvolutionRelaxation (SURVEY.md §8 f3;
lib/NonlinearSolveFirstOrder/src/pseudo_transient.jl:37-56, 157-170; lib/NonlinearSolveBase/src/descent/damped_newton.jl:290-340):
(J + I/alpha) du = -f with alpha_{n+1} = alpha_n ||f_{n-1}||_2 / ||f_n||_2."""
import numpy as np
import pytest


def _pt(po, **kw):
    return po.default_newton_opts(descent=po.DESCENT_PSEUDO_TRANSIENT, **kw)


def test_oracle_pseudo_transient_reference_anchor(po):
    # rootfind_tests__item5.jl:29-33: PseudoTransient(alpha_initial = 10) on u.^2 .- 2, u0 = ones -> err < 1e-9 (abstol 1e-9)
    P = po.OracleProblem.quadratic(1000, 2.0)
    u0 = np.ones(1000)
    for ls in (po.LINSOLVE_DENSE_LU, po.LINSOLVE_GMRES):
        u, f, r, tr = P.newton(u0, _pt(po, abstol=1e-9, linsolve=ls, pt_alpha_initial=10.0))
        assert r.retcode == po.RC_SUCCESS and np.abs(u * u - 2.0).max() < 1e-9
        # switched evolution relaxation: alpha_1 = alpha_0 (ratio 1 at the first solve), then alpha grows as the residual falls
        al = [t.trust_radius for t in tr]
        fn = [np.abs(u0 * u0 - 2.0).max()] + [t.fnorm_inf for t in tr]   # uniform problem: ratio of inf-norms == ratio of 2-norms
        assert al[0] == 10.0
        for k in range(1, len(al)):
            assert abs(al[k] / al[k - 1] - fn[k - 1] / fn[k]) <= 1e-9 * (fn[k - 1] / fn[k])
        # first step solves (2u + 1/alpha) x = f exactly
        assert abs(tr[0].step_norm2 - np.sqrt(1000.0) * abs((1.0 - 2.0) / (2.0 + 0.1))) < 1e-9


def test_oracle_pseudo_transient_limits(po):
    P = po.OracleProblem.bruss2d(12)
    u0 = P.u0()
    un, fn, rn, _ = P.newton(u0, po.default_newton_opts(abstol=1e-8, linsolve=po.LINSOLVE_DENSE_LU))
    # alpha -> infinity recovers Newton
    u, f, r, tr = P.newton(u0, _pt(po, abstol=1e-8, linsolve=po.LINSOLVE_DENSE_LU, pt_alpha_initial=1e14))
    assert r.retcode == po.RC_SUCCESS and r.nsteps == rn.nsteps and np.abs(u - un).max() <= 1e-9 * np.abs(un).max()
    # a moderate pseudo time step still converges to the same root, in more steps, with the same NLStats rules
    for ls in (po.LINSOLVE_DENSE_LU, po.LINSOLVE_GMRES, po.LINSOLVE_SPARSE_GMRES):
        u, f, r, tr = P.newton(u0, _pt(po, abstol=1e-8, linsolve=ls, gmres_orth=po.ORTH_CGS2, pt_alpha_initial=0.1))
        assert r.retcode == po.RC_SUCCESS and r.nsteps > rn.nsteps and r.nf == r.nsteps == r.nsolve
        assert np.abs(f).max() < 1e-8 and np.abs(u - un).max() <= 1e-6 * np.abs(un).max()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["dense", "gmres", "sparse", "gmres_linesearch"])
def test_gpu_pseudo_transient_vs_oracle(nls, ctx, po, variant):
    N = 12
    P = po.OracleProblem.bruss2d(N)
    u0 = P.u0()
    f = nls.Brusselator2D(N)
    okw = dict(abstol=1e-8, pt_alpha_initial=0.1)
    if variant == "dense":
        alg, okw["linsolve"] = nls.PseudoTransient(alpha_initial=0.1), po.LINSOLVE_DENSE_LU
    elif variant == "gmres":
        alg = nls.PseudoTransient(alpha_initial=0.1, linsolve=nls.KrylovJL_GMRES())
        okw.update(linsolve=po.LINSOLVE_GMRES, gmres_orth=po.ORTH_CGS2)
    elif variant == "sparse":
        f = nls.NonlinearFunction(f, sparsity=nls.TracerSparsityDetector())
        alg = nls.PseudoTransient(alpha_initial=0.1, linsolve=nls.KrylovJL_GMRES())
        okw.update(linsolve=po.LINSOLVE_SPARSE_GMRES, gmres_orth=po.ORTH_CGS2)
    else:
        alg = nls.PseudoTransient(alpha_initial=0.1, linsolve=nls.KrylovJL_GMRES(), linesearch=nls.BackTracking())
        okw.update(linsolve=po.LINSOLVE_GMRES, gmres_orth=po.ORTH_CGS2, globalization=po.GLOB_LINESEARCH)
    sol = nls.solve(nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx), alg, abstol=1e-8)
    uo, fo, ro, tro = P.newton(u0, _pt(po, **okw))
    assert sol.retcode == ro.retcode == po.RC_SUCCESS
    assert (sol.stats.nsteps, sol.stats.nf, sol.stats.nsolve, sol.stats.njacs) == (ro.nsteps, ro.nf, ro.nsolve, ro.njacs)
    assert np.abs(sol.u - uo).max() <= 1e-6 * np.abs(uo).max() and np.abs(sol.resid).max() < 1e-8
    if variant != "gmres_linesearch":   # the trace slot carries alpha (line-search runs put the step length there)
        # compare the damping actually applied, 1/alpha: near convergence alpha ~ 1e9 is a ratio of rounding-level residual
        # norms and only its reciprocal (~1e-9 against Jacobian entries of 1e3) is meaningful
        for tg, t in zip(sol.trace, tro):
            assert abs(1.0 / tg.trust_radius - 1.0 / t.trust_radius) <= 1e-6 / t.trust_radius + 1e-9
    # reference anchor on the device: u.^2 .- 2 with alpha_initial = 10 (rootfind_tests__item5.jl:29-33)
    for alg2 in (nls.PseudoTransient(alpha_initial=10.0), nls.PseudoTransient(alpha_initial=10.0, linsolve=nls.KrylovJL_GMRES())):
        s2 = nls.solve(nls.NonlinearProblem(nls.QuadraticFunction(1000), np.ones(1000), 2.0, ctx=ctx), alg2, abstol=1e-9)
        assert nls.successful_retcode(s2.retcode) and np.abs(s2.u * s2.u - 2.0).max() < 1e-9
