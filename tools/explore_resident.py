"""GPU exploration: resident (cooperative, TMA-staged) Arnoldi engine vs the multi-kernel engine.  python tools/explore_resident.py N [engine ...]"""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import nonlinearsolve_jl_b200 as nls  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

N = int(sys.argv[1])
engines = sys.argv[2:] or ["multikernel", "resident"]
ctx = nls.Context(0)
# small parity check first (2D N=12 and 3D N=8) against the oracle's MGS2-equivalent (CGS2) run
for kind, n_ in (("2d", 12), ("3d", 8)):
    f = nls.Brusselator2D(n_) if kind == "2d" else nls.Brusselator3D(n_)
    P = po.OracleProblem.bruss2d(n_) if kind == "2d" else po.OracleProblem.bruss3d(n_)
    dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
    u = P.u0(1)
    b = P.residual(u)
    xo, so = po.gmres(b, prob=P, u=u, opts=po.default_gmres_opts(atol=1e-10, rtol=1e-10, orth=po.ORTH_CGS2))
    gm = nls.GmresSolver(ctx, P.n, nls.KrylovJL_GMRES(orth="cgs2", engine="resident", check_every=4), atol=1e-10, rtol=1e-10)
    x, st = gm.solve(nls.JacobianOperator(dp, ctx.to_device(u)), ctx.to_device(b))
    err = np.abs(x.to_host() - xo).max() / np.abs(xo).max()
    print(json.dumps({"parity": kind, "N": n_, "status": st.status, "iters": st.iters, "oracle_iters": so.iters, "rel_err": err}))
    assert st.status == 1 and err < 1e-6, (st.status, err)

f = nls.Brusselator3D(N)
dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
u0 = dp.u0(1)
prob = nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx)
for eng in engines:
    cache = nls.init(prob, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(orth="cgs2", engine=eng)), abstol=1e-8, maxiters=20)
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
    print(json.dumps({"N": N, "engine": eng, "wall_s": round(dt, 3), "retcode": nls.ReturnCode.name(sol.retcode), "nsteps": sol.stats.nsteps,
                      "njvp": sol.stats.njvp, "jvps_per_s": round(sol.stats.njvp / dt, 1), "resid_inf": sol.resid_inf,
                      "launches": ctx.kernel_launches() - l0, "lin_iters": [t_.lin_iters for t_ in sol.trace],
                      "bytes_TB": round(sol.bytes_moved / 1e12, 2), "whole_gbs": round(sol.bytes_moved / dt / 1e9, 0),
                      "profile": {k: {"ms": round(v["ms"], 1), "gbs": round(v["gbs"], 0), "n": v["launches"]} for k, v in rep_.items()}}))
    del cache
