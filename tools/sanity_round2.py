"""compute-sanitizer target for the round-2 kernels: multigrid V-cycle (all transfer ratios), banded LU, the new reductions
behind the termination modes, the r3g resident kernel on small grids.   compute-sanitizer --tool memcheck python tools/sanity_round2.py"""
import sys
sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import nonlinearsolve_jl_b200 as nls  # noqa: E402

ctx = nls.Context(0)
for f in (nls.Brusselator2D(12), nls.Brusselator2D(15), nls.Brusselator3D(10), nls.Brusselator3D(7)):
    even = (f.N ** (3 if isinstance(f, nls.Brusselator3D) else 2)) % 2 == 0   # the resident engine needs an even cell count
    dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
    u0 = dp.u0(1)
    for alg, kw in ((nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(precs=nls.Multigrid("right"))), {}),
                    (nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(precs=nls.Multigrid("left"), atol=1e-13, rtol=1e-9)), {}),
                    (nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(engine="resident" if even else "auto", orth="mgs")), dict(termination_condition=nls.RelNormSafeBestTerminationMode(norm="l2"))),
                    (nls.TrustRegion(radius_update_scheme=nls.RadiusUpdateSchemes.Bastin), dict(termination_condition=nls.RelTerminationMode(), reltol=1e-10, maxiters=40))):
        sol = nls.solve(nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx), alg, abstol=1e-8, **kw)
        print(type(f).__name__, f.N, alg.name, nls.ReturnCode.name(sol.retcode), sol.stats.nsteps, sol.stats.njvp, sol.resid_inf, flush=True)
    fs = nls.NonlinearFunction(f, sparsity=nls.TracerSparsityDetector())
    sol = nls.solve(nls.NonlinearProblem(fs, u0, (3.4, 1.0, 10.0), ctx=ctx), nls.NewtonRaphson(), abstol=1e-8)
    print(type(f).__name__, f.N, "sparse direct", nls.ReturnCode.name(sol.retcode), sol.stats.nsteps, sol.stats.nfactors, sol.resid_inf, flush=True)
