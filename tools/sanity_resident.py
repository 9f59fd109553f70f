"""compute-sanitizer target for the resident Arnoldi kernels (racecheck / memcheck / synccheck):
small grids through whole Newton solves (both Gram-Schmidt variants, matrix-free and assembled-sparse operator) and a burst
of full-size Arnoldi steps at the benchmarked size (3D N=100: 27 row pairs per thread, register stage full, cp.async annex
in use).      compute-sanitizer --tool racecheck python tools/sanity_resident.py [N_burst] [steps]"""
import sys
sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import nonlinearsolve_jl_b200 as nls  # noqa: E402

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 100
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 7
ctx = nls.Context(0)
for f in (nls.Brusselator2D(12), nls.Brusselator3D(8)):
    dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
    u0 = dp.u0(1)
    for alg in (nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(engine="resident", orth="mgs")),
                nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(engine="resident", orth="cgs2")),
                nls.TrustRegion(linsolve=nls.KrylovJL_GMRES(engine="resident", orth="mgs"), concrete_jac=True)):
        sol = nls.solve(nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx), alg, abstol=1e-8)
        print(type(f).__name__, alg.name, nls.ReturnCode.name(sol.retcode), sol.stats.nsteps, sol.stats.njvp, sol.resid_inf, flush=True)
f = nls.Brusselator3D(NB)
dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
u = dp.u0(1)
b = dp.residual(u)
for orth in ("mgs", "cgs2"):
    gm = nls.GmresSolver(ctx, dp.n, nls.KrylovJL_GMRES(orth=orth, engine="resident", itmax=STEPS), atol=0.0, rtol=1e-14)
    x, st = gm.solve(nls.JacobianOperator(dp, u), b)
    print("burst N=%d" % NB, orth, st.status, st.iters, st.rnorm / st.rnorm0, flush=True)
    del gm
