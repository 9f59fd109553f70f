"""What the linear tolerance of BASELINE config 3 means for each Gram-Schmidt variant on the GPU: first Newton step of the 3D
N = 100 Brusselator, GMRES to atol 1e-8 + rtol 3e-13 ||r0|| (the tolerances NewtonRaphson hands to the linear solver), MGS (one
pass) and the reorthogonalised variant, resident engine.  Prints iterations, the solver's residual estimate and the TRUE residual
||b - J x|| recomputed with the JVP kernel.   python tools/mgs_attainable.py [N]"""
import json
import sys
sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import nonlinearsolve_jl_b200 as nls  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = nls.Context(0)
f = nls.Brusselator3D(N)
dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
u = dp.u0(1)
b = dp.residual(u)
out = {}
for orth in ("mgs", "cgs2"):
    gm = nls.GmresSolver(ctx, dp.n, nls.KrylovJL_GMRES(orth=orth, engine="resident"), atol=1e-8, rtol=3e-13)
    x, st = gm.solve(nls.JacobianOperator(dp, u), b)
    r = dp.jvp(u, x)
    r.scal(-1.0).axpy(1.0, b)
    out[orth] = {"iters": st.iters, "status": st.status, "rnorm0": st.rnorm0, "tol": 1e-8 + 3e-13 * st.rnorm0, "estimate": st.rnorm, "true_residual": r.norm(2),
                 "true_over_r0": r.norm(2) / st.rnorm0}
print(json.dumps(out, indent=1))
