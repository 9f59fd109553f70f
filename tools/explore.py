"""GPU exploration helper (not part of the product): run the headline workload at a few sizes and print the Newton /
GMRES trace, timings and the per-kernel-family profile.  Usage: python tools/explore.py N [orth] [itmax] [abstol] [rtol]"""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import nonlinearsolve_jl_b200 as nls  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
orth = sys.argv[2] if len(sys.argv) > 2 else "cgs2"
itmax = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
abstol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-8
rtol = float(sys.argv[5]) if len(sys.argv) > 5 else None
ctx = nls.Context(0)
f = nls.Brusselator3D(N)
prob = nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx)
dp = nls._DeviceProblem(ctx, prob)
u0 = dp.u0(1)
prob = nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx)
cache = nls.init(prob, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(orth=orth, itmax=itmax, rtol=rtol)), abstol=abstol, maxiters=20)
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
    print(json.dumps({"N": N, "orth": orth, "rep": rep, "wall_s": round(dt, 3), "retcode": nls.ReturnCode.name(sol.retcode),
                      "nsteps": sol.stats.nsteps, "njvp": sol.stats.njvp, "resid_inf": sol.resid_inf, "launches": ctx.kernel_launches() - l0,
                      "bytes_TB": round(sol.bytes_moved / 1e12, 3), "achieved_GBs_wall": round(sol.bytes_moved / dt / 1e9, 1),
                      "trace": [(t_.iter, t_.lin_iters, t_.lin_status, float("%.3e" % t_.fnorm_inf), float("%.3e" % t_.lin_rnorm)) for t_ in sol.trace],
                      "profile": {k: {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items()} for k, v in rep_.items()}}))
