"""GPU: phase timers of the resident Arnoldi kernel on the headline problem (single GMRES solve, itmax iterations)."""
import ctypes as C
import json
import sys
import time
sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import nonlinearsolve_jl_b200 as nls  # noqa: E402
N = int(sys.argv[1]); itmax = int(sys.argv[2]) if len(sys.argv) > 2 else 300
ORTH = sys.argv[3] if len(sys.argv) > 3 else "cgs2"
ENGINES = sys.argv[4].split(",") if len(sys.argv) > 4 else ("resident", "multikernel")
ctx = nls.Context(0)
f = nls.Brusselator3D(N)
dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
u = dp.u0(1); b = dp.residual(u)
for eng in ENGINES:
    gm = nls.GmresSolver(ctx, dp.n, nls.KrylovJL_GMRES(orth=ORTH, engine=eng, itmax=itmax), atol=0.0, rtol=1e-14)
    out = np.zeros(4 * 148 + 3)
    for rep in range(2):
        nls.abi.check(ctx.handle, nls.abi.lib().b200_gmres_debug(gm._h, 1, None, 0))
        ctx.sync(); t = time.time()
        x, st = gm.solve(nls.JacobianOperator(dp, u), b)
        ctx.sync(); dt = time.time() - t
        nls.abi.check(ctx.handle, nls.abi.lib().b200_gmres_debug(gm._h, 1, out.ctypes.data_as(C.c_void_p), 4 * 148 + 3))
    steps = max(out[-3], 1)
    print("poll rounds|pure update (CTA 0):", out[-2] / steps, "publish cycles:", out[-1] / steps)
    ph = out[:-3].reshape(148, 4) / steps
    print(json.dumps({"engine": eng, "iters": st.iters, "wall_s": round(dt, 3), "us_per_vector_pass": round(dt * 1e6 / (itmax * (itmax + 1)), 3), "steps": steps,
                      "mean": [round(x) for x in ph.mean(0)], "min": [round(x) for x in ph.min(0)], "max": [round(x) for x in ph.max(0)],
                      "argmax_wait": int(ph[:, 0].argmax()), "argmin_gather": int(ph[:, 2].argmin()), "wait_sorted_top": [round(x) for x in np.sort(ph[:, 0])[-6:]]}))
    del gm
