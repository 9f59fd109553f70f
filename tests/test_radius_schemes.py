"""RadiusUpdateSchemes of the trust-region globalisation (lib/NonlinearSolveFirstOrder/src/trust_region.jl:
defaults 330-384, update rules 431-509): Simple, NLsolve, NocedalWright, Hei, Yuan, Fan."""
import numpy as np
import pytest

SCHEMES = ["Simple", "NLsolve", "NocedalWright", "Hei", "Yuan", "Fan"]


def _opts(po, sch, **kw):
    return po.default_newton_opts(globalization=po.GLOB_TRUST_REGION, tr_scheme=sch, **kw)


def test_oracle_radius_schemes_reference_anchor(po):
    # rootfind tests of the reference run every radius update scheme on u.^2 .- 2 from ones and require err < 1e-9
    P = po.OracleProblem.quadratic(1000, 2.0)
    for sch in range(6):
        for ls in (po.LINSOLVE_GMRES, po.LINSOLVE_DENSE_LU):
            u, f, r, tr = P.newton(np.ones(1000), _opts(po, sch, abstol=1e-9, linsolve=ls))
            assert r.retcode == po.RC_SUCCESS and np.abs(u * u - 2.0).max() < 1e-9, (sch, ls)
            assert r.nf == r.nsteps  # one trial evaluation per step (trust_region.jl:401-402)


def test_oracle_radius_scheme_defaults(po):
    P = po.OracleProblem.bruss2d(12)
    u0 = P.u0()
    f0 = P.residual(u0)
    rad = {}
    for sch in range(6):
        u, f, r, tr = P.newton(u0, _opts(po, sch, abstol=1e-8, gmres_orth=po.ORTH_CGS2, maxiters=1))
        rad[sch] = tr[0]
    fu_norm, u_norm = np.linalg.norm(f0), np.linalg.norm(u0)
    mtr = max(fu_norm, u0.max() - u0.min())
    # first accepted step of Simple doubles max_tr / 11 (clamped); NocedalWright keeps it unless the step sits on the boundary
    assert abs(rad[po.TR_SIMPLE].trust_radius - min(2 * mtr / 11, mtr)) < 1e-9 * mtr
    assert abs(rad[po.TR_NOCEDAL_WRIGHT].trust_radius - mtr / 11) < 1e-9 * mtr
    # Fan: after one very successful step p1 = min(0.1 * 12, 1e18) and the radius is p1 ||f(u_1)||_2^0.99  (trust_region.jl:491-499)
    u1 = P.newton(u0, _opts(po, po.TR_FAN, abstol=1e-8, gmres_orth=po.ORTH_CGS2, maxiters=1))[0]
    assert abs(rad[po.TR_FAN].trust_radius - 1.2 * np.linalg.norm(P.residual(u1)) ** 0.99) < 1e-9 * rad[po.TR_FAN].trust_radius
    # NLsolve starts from ||u0||, Hei from 1: the first radius is expand_factor * ||du|| resp. rfunc(rho) * ||du||
    assert rad[po.TR_NLSOLVE].trust_radius > 0 and rad[po.TR_HEI].trust_radius > 0
    for sch in range(6):
        u, f, r, tr = P.newton(u0, _opts(po, sch, abstol=1e-8, gmres_orth=po.ORTH_CGS2, maxiters=200))
        assert r.retcode == po.RC_SUCCESS and np.abs(f).max() < 1e-8, sch


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", SCHEMES)
def test_gpu_radius_schemes_vs_oracle(nls, ctx, po, scheme):
    code = getattr(nls.RadiusUpdateSchemes, scheme)
    N = 12
    P = po.OracleProblem.bruss2d(N)
    u0 = P.u0() * 10.0 + 0.1          # far enough from the root that every scheme rejects and shrinks a few times
    for linsolve, okw in ((nls.KrylovJL_GMRES(), dict(linsolve=po.LINSOLVE_GMRES, gmres_orth=po.ORTH_CGS2)), (None, dict(linsolve=po.LINSOLVE_DENSE_LU))):
        sol = nls.solve(nls.NonlinearProblem(nls.Brusselator2D(N), u0, (3.4, 1.0, 10.0), ctx=ctx),
                        nls.TrustRegion(linsolve=linsolve, radius_update_scheme=code), abstol=1e-8, maxiters=200)
        uo, fo, ro, tro = P.newton(u0, _opts(po, code, abstol=1e-8, maxiters=200, **okw))
        assert sol.retcode == ro.retcode
        assert (sol.stats.nsteps, sol.stats.nf, sol.stats.nsolve) == (ro.nsteps, ro.nf, ro.nsolve)
        assert np.abs(sol.u - uo).max() <= 1e-6 * np.abs(uo).max()
        for tg, t in zip(sol.trace, tro):
            assert tg.accepted == t.accepted
            # Fan / Yuan radii are proportional to ||f(u_trial)|| resp. ||J' f(u_trial)||: rounding-level at convergence -> absolute floor
            assert abs(tg.trust_radius - t.trust_radius) <= 1e-6 * t.trust_radius + 1e-7
    # the reference's anchor on the device
    s2 = nls.solve(nls.NonlinearProblem(nls.QuadraticFunction(1000), np.ones(1000), 2.0, ctx=ctx),
                   nls.TrustRegion(linsolve=nls.KrylovJL_GMRES(), radius_update_scheme=code), abstol=1e-9)
    assert nls.successful_retcode(s2.retcode) and np.abs(s2.u * s2.u - 2.0).max() < 1e-9
