"""Small end-to-end exercise of every engine (for compute-sanitizer runs)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import nonlinearsolve_jl_b200 as nls
ctx = nls.Context(0)
for f in (nls.Brusselator2D(12), nls.Brusselator3D(8)):
    dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
    u0 = dp.u0(1)
    for alg in (nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(engine="multikernel")), nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(engine="resident")),
                nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(orth="mgs")), nls.NewtonRaphson(), nls.TrustRegion(linsolve=nls.KrylovJL_GMRES(), concrete_jac=True),
                nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(engine="resident", orth="mgs"), linesearch=nls.BackTracking()),
                nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(precs=nls.BlockJacobi("right"))), nls.PseudoTransient(alpha_initial=0.1, linsolve=nls.KrylovJL_GMRES()),
                nls.TrustRegion(linsolve=nls.KrylovJL_GMRES(), radius_update_scheme=nls.RadiusUpdateSchemes.Yuan)):
        sol = nls.solve(nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx), alg, abstol=1e-8)
        print(type(f).__name__, alg.name, nls.ReturnCode.name(sol.retcode), sol.stats.nsteps, sol.resid_inf)
K = 5
c = nls.EnsembleCache(ctx, 8, K, 10.0, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES()), abstol=1e-8)
d8 = nls._DeviceProblem(ctx, nls.NonlinearProblem(nls.Brusselator2D(8), None, (3.4, 1.0, 10.0), ctx=ctx))
r = c.solve(ctx.to_device(np.tile(d8.u0().to_host(), K)), ctx.to_device(np.full(K, 3.4)), ctx.to_device(np.full(K, 1.0)))
print("ensemble", r.nsuccess)
