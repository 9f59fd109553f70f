"""GPU exploration: does the reference's default orthogonalisation (MGS, no reorthogonalisation) converge on the headline problem?"""
import json, sys, time
sys.path.insert(0, ".")
import nonlinearsolve_jl_b200 as nls
N = int(sys.argv[1]); itmax = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
ctx = nls.Context(0)
f = nls.Brusselator3D(N)
dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
u0 = dp.u0(1)
for orth in ("mgs", "cgs2"):
    cache = nls.init(nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx), nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(orth=orth, engine="resident", itmax=itmax)), abstol=1e-8, maxiters=8)
    t = time.time(); sol = cache.solve(to_host=False); ctx.sync(); dt = time.time() - t
    print(json.dumps({"N": N, "orth": orth, "wall_s": round(dt, 2), "retcode": nls.ReturnCode.name(sol.retcode), "nsteps": sol.stats.nsteps, "njvp": sol.stats.njvp,
                      "resid_inf": sol.resid_inf, "trace": [(t_.lin_iters, t_.lin_status, float("%.2e" % t_.fnorm_inf), float("%.2e" % t_.lin_rnorm)) for t_ in sol.trace]}))
    del cache
