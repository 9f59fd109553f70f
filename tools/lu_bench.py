"""GPU: time b200_getrf on a Brusselator Jacobian (config 2 shape) or a random matrix; verify with a solve.  python tools/lu_bench.py n [n ...]"""
import ctypes as C
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import nonlinearsolve_jl_b200 as nls  # noqa: E402

ctx = nls.Context(0)
L = nls.abi.lib()
for arg in sys.argv[1:]:
    N = int(arg)  # 2D Brusselator grid size; n = 2 N^2
    f = nls.Brusselator2D(N)
    n = f.n()
    dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
    u = dp.u0()
    J = ctx.empty(n * n)
    ipiv = ctx.empty(n, np.int64)
    b = dp.residual(u)
    info = C.c_int32()
    for rep in range(2):
        nls.abi.check(ctx.handle, L.b200_dense_jac_fill(dp.handle, u.ptr, J.ptr, n))
        ctx.sync()
        ctx.profile(True)
        t = time.time()
        nls.abi.check(ctx.handle, L.b200_getrf(ctx.handle, n, J.ptr, n, ipiv.ptr, C.byref(info)))
        ctx.sync()
        dt = time.time() - t
        prof = ctx.profile_report()
        ctx.profile(False, reset=False)
    x = b.copy()
    t = time.time()
    nls.abi.check(ctx.handle, L.b200_getrs(ctx.handle, n, 1, J.ptr, n, ipiv.ptr, x.ptr, n))
    ctx.sync()
    dts = time.time() - t
    r = dp.jvp(u, x)  # J x
    rel = np.abs(r.to_host() - b.to_host()).max() / np.abs(b.to_host()).max()
    fl = (2.0 / 3.0) * n ** 3
    print(json.dumps({"N": N, "n": n, "getrf_s": round(dt, 4), "tflops": round(fl / dt / 1e12, 2), "getrs_s": round(dts, 4), "info": info.value,
                      "solve_rel_resid": rel,
                      "profile": {k: {"ms": round(v["ms"], 1), "n": v["launches"], "tflops": round(v["bytes"] / (v["ms"] * 1e-3) / 1e12, 2) if k == "lu_gemm" else None}
                                  for k, v in prof.items() if k.startswith("lu")}}))
    del J
