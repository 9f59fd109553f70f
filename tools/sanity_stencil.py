"""compute-sanitizer target for the stand-alone stencil kernels (TMA ring, producer / consumer warps): N = 24 (two chunks, both
wraps in one copy), 26 (ragged last chunk), 100 (the benchmarked shape), residual with the fused norm, JVP, VJP.
    compute-sanitizer --tool racecheck python tools/sanity_stencil.py"""
import sys
sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import nonlinearsolve_jl_b200 as nls  # noqa: E402

ctx = nls.Context(0)
rng = np.random.default_rng(0)
for N in (24, 26, 100):
    f = nls.Brusselator3D(N)
    dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
    u, v = dp.u0(1), ctx.to_device(rng.standard_normal(dp.n))
    r, jv, jtv = dp.residual(u), dp.jvp(u, v), dp.vjp(u, v)
    f2, j2 = dp.residual_jvp(u, v)
    print("ring", N, float(np.abs(r.to_host() - f2.to_host()).max()), float(np.abs(jv.to_host() - j2.to_host()).max()), jtv.norm(2), flush=True)
