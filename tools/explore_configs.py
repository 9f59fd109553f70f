"""GPU exploration helper: BASELINE configs 2 (dense LU) and 4 (sparse coloured Jacobian + TrustRegion) at a given size."""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import nonlinearsolve_jl_b200 as nls  # noqa: E402

which = sys.argv[1]
N = int(sys.argv[2])
ctx = nls.Context(0)
if which == "dense":
    f = nls.Brusselator2D(N)
    alg = nls.NewtonRaphson()
else:
    f = nls.NonlinearFunction(nls.Brusselator3D(N), sparsity=nls.TracerSparsityDetector())
    alg = nls.TrustRegion(linsolve=nls.KrylovJL_GMRES(orth=sys.argv[3] if len(sys.argv) > 3 else "cgs2", engine=sys.argv[4] if len(sys.argv) > 4 else "auto"))
base = f.f if isinstance(f, nls.NonlinearFunction) else f
dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(base, None, (3.4, 1.0, 10.0), ctx=ctx))
u0 = dp.u0(1)
t = time.time()
cache = nls.init(nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx), alg, abstol=1e-8, maxiters=30)
ctx.sync()
t_init = time.time() - t
for rep in range(2):
    cache.reinit(u0)
    ctx.profile(True)
    t = time.time()
    sol = cache.solve(to_host=False)
    ctx.sync()
    dt = time.time() - t
    prof = ctx.profile_report()
    ctx.profile(False, reset=False)
    st = sol.stats
    out = {"config": which, "N": N, "n": dp.n, "rep": rep, "init_s": round(t_init, 3), "solve_s": round(dt, 3), "retcode": nls.ReturnCode.name(sol.retcode),
           "nsteps": st.nsteps, "nf": st.nf, "njacs": st.njacs, "nfactors": st.nfactors, "nsolve": st.nsolve, "njvp": st.njvp, "resid_inf": sol.resid_inf,
           "trace": [(t_.iter, t_.lin_iters, t_.accepted, float("%.3e" % t_.fnorm_inf), float("%.3e" % t_.trust_radius)) for t_ in sol.trace]}
    if which == "dense":
        out["lu_tflops_incl_fill_and_solve"] = round(st.nfactors * (2.0 / 3.0) * dp.n ** 3 / dt / 1e12, 3)
    print(json.dumps(out))
