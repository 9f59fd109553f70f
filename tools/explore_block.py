"""GPU exploration: headline workload with L2-blocked Gram-Schmidt, several block sizes.  python tools/explore_block.py N b1 b2 ..."""
import json
import sys
import time

sys.path.insert(0, ".")
import nonlinearsolve_jl_b200 as nls  # noqa: E402

N = int(sys.argv[1])
blocks = [int(b) for b in sys.argv[2:]] or [0, 2, 4]
ctx = nls.Context(0)
f = nls.Brusselator3D(N)
dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
u0 = dp.u0(1)
prob = nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx)
for b in blocks:
    cache = nls.init(prob, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(orth="cgs2", block=b)), abstol=1e-8, maxiters=20)
    for rep in range(2):
        cache.reinit(u0)
        ctx.profile(True)
        l0 = ctx.kernel_launches()
        t = time.time()
        sol = cache.solve(to_host=False)
        ctx.sync()
        dt = time.time() - t
        rep_ = ctx.profile_report()
        ctx.profile(False, reset=False)
    print(json.dumps({"N": N, "block": b, "wall_s": round(dt, 3), "retcode": nls.ReturnCode.name(sol.retcode), "nsteps": sol.stats.nsteps,
                      "njvp": sol.stats.njvp, "jvps_per_s": round(sol.stats.njvp / dt, 1), "resid_inf": sol.resid_inf,
                      "launches": ctx.kernel_launches() - l0, "lin_iters": [t_.lin_iters for t_ in sol.trace],
                      "profile": {k: {"ms": round(v["ms"], 1), "gbs": round(v["gbs"], 0), "n": v["launches"]} for k, v in rep_.items()}}))
    del cache
