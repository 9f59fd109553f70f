"""GPU: wall time of a burst of resident Arnoldi steps on the headline problem (single GMRES solve, `itmax` iterations).
    python tools/resident_phases.py N [itmax] [mgs|cgs2] [engine,engine...]"""
import json
import sys
import time
sys.path.insert(0, ".")
import nonlinearsolve_jl_b200 as nls  # noqa: E402
N = int(sys.argv[1]); itmax = int(sys.argv[2]) if len(sys.argv) > 2 else 300
ORTH = sys.argv[3] if len(sys.argv) > 3 else "mgs"
ENGINES = sys.argv[4].split(",") if len(sys.argv) > 4 else ("resident", "multikernel")
ctx = nls.Context(0)
f = nls.Brusselator3D(N)
dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
u = dp.u0(1); b = dp.residual(u)
for eng in ENGINES:
    gm = nls.GmresSolver(ctx, dp.n, nls.KrylovJL_GMRES(orth=ORTH, engine=eng, itmax=itmax), atol=0.0, rtol=1e-14)
    for rep in range(3):
        ctx.sync(); t = time.time()
        x, st = gm.solve(nls.JacobianOperator(dp, u), b)
        ctx.sync(); dt = time.time() - t
    passes = 1 if ORTH == "mgs" else 2
    print(json.dumps({"engine": eng, "orth": ORTH, "iters": st.iters, "wall_s": round(dt, 4), "us_per_basis_vector": round(dt * 1e6 / (passes * itmax * (itmax + 1) / 2), 3)}))
    del gm
