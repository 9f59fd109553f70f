"""BackTracking line search (NewtonRaphson(linesearch = BackTracking()), solve.jl:249-273, 392-408).
LineSearch.jl is an external, unpinned dependency of the reference: the oracle restates the LineSearches.jl BackTracking
algorithm (cubic interpolation) and is the parity target; the reference's own acceptance bound (residual below abstol) is
asserted as well."""
import numpy as np
import pytest


def _case(po, N, scale):
    P = po.OracleProblem.bruss2d(N)
    return P, P.u0() * scale + 0.1


def test_oracle_backtracking_triggers_and_converges(po):
    P, u0 = _case(po, 12, 10.0)
    plain = P.newton(u0, po.default_newton_opts(abstol=1e-8, linsolve=po.LINSOLVE_DENSE_LU, maxiters=60))
    ls = P.newton(u0, po.default_newton_opts(abstol=1e-8, linsolve=po.LINSOLVE_DENSE_LU, globalization=po.GLOB_LINESEARCH, maxiters=60))
    assert plain[2].retcode == ls[2].retcode == po.RC_SUCCESS
    alphas = [t.trust_radius for t in ls[3]]
    assert min(alphas) < 1.0 and alphas[-1] == 1.0          # backtracked early, full steps near the root
    assert ls[2].nsteps < plain[2].nsteps                    # and it pays off on this start
    assert ls[2].nf > ls[2].nsteps                           # every trial evaluation is counted (NLStats.nf)
    assert np.abs(ls[1]).max() < 1e-8
    # from the reference initial condition the full Newton step is always accepted: identical iterates
    P2, u02 = _case(po, 16, 1.0)
    a = P2.newton(u02, po.default_newton_opts(abstol=1e-8, linsolve=po.LINSOLVE_GMRES, gmres_orth=po.ORTH_CGS2))
    b = P2.newton(u02, po.default_newton_opts(abstol=1e-8, linsolve=po.LINSOLVE_GMRES, gmres_orth=po.ORTH_CGS2, globalization=po.GLOB_LINESEARCH))
    assert np.array_equal(a[0], b[0]) and all(t.trust_radius == 1.0 for t in b[3])


@pytest.mark.gpu
@pytest.mark.parametrize("linsolve", ["dense", "gmres"])
def test_gpu_backtracking_vs_oracle(nls, ctx, po, linsolve):
    N = 12
    P, u0 = _case(po, N, 10.0)
    if linsolve == "dense":
        alg = nls.NewtonRaphson(linesearch=nls.BackTracking())
        okw = dict(linsolve=po.LINSOLVE_DENSE_LU)
    else:
        alg = nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(), linesearch=nls.BackTracking())
        okw = dict(linsolve=po.LINSOLVE_GMRES, gmres_orth=po.ORTH_CGS2)
    sol = nls.solve(nls.NonlinearProblem(nls.Brusselator2D(N), u0, (3.4, 1.0, 10.0), ctx=ctx), alg, abstol=1e-8, maxiters=60)
    uo, fo, ro, tro = P.newton(u0, po.default_newton_opts(abstol=1e-8, globalization=po.GLOB_LINESEARCH, maxiters=60, **okw))
    assert sol.retcode == ro.retcode == po.RC_SUCCESS
    assert sol.stats.nsteps == ro.nsteps and sol.stats.nf == ro.nf
    assert np.abs(sol.u - uo).max() <= 1e-6 * np.abs(uo).max() and np.abs(sol.resid).max() < 1e-8
    for tg, t in zip(sol.trace, tro):   # same accepted step lengths
        assert abs(tg.trust_radius - t.trust_radius) <= 1e-6 * max(t.trust_radius, 1e-3)
