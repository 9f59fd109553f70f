"""Where the multigrid-preconditioned config-3 solve spends its time: wall clock against the per-family CUDA-event totals
(ctx.profile), launches per solve.   python tools/precond_breakdown.py"""
import json
import sys
import time
sys.path.insert(0, ".")
import nonlinearsolve_jl_b200 as nls  # noqa: E402

ctx = nls.Context(0)
f = nls.Brusselator3D(100)
dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
u0 = dp.u0(1)
out = {}
for name, ls in (("mgs", nls.KrylovJL_GMRES(orth="mgs", precs=nls.Multigrid("right"))), ("cgs2", nls.KrylovJL_GMRES(orth="cgs2", precs=nls.Multigrid("right")))):
    alg = nls.NewtonRaphson(linsolve=ls)
    prob = nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx)
    for _ in range(2):
        nls.solve(prob, alg, abstol=1e-8)
    ctx.sync()
    l0 = ctx.kernel_launches()
    t0 = time.perf_counter()
    for _ in range(5):
        sol = nls.solve(prob, alg, abstol=1e-8)
    ctx.sync()
    wall = (time.perf_counter() - t0) / 5
    launches = (ctx.kernel_launches() - l0) / 5
    ctx.profile(True)
    sol = nls.solve(prob, alg, abstol=1e-8)
    ctx.sync()
    rep = ctx.profile_report()
    ctx.profile(False)
    out[name] = {"wall_ms": wall * 1e3, "launches": launches, "njvp": sol.stats.njvp, "nsteps": sol.stats.nsteps,
                 "event_ms_total": sum(v["ms"] for v in rep.values()), "families": {k: {"ms": round(v["ms"], 3), "n": v["launches"]} for k, v in rep.items()}}
print(json.dumps(out, indent=1))
