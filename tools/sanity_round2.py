"""compute-sanitizer target for the round-2 kernels: multigrid V-cycle (all transfer ratios), banded LU, the new reductions
behind the termination modes, the r3g resident kernel on small grids.   compute-sanitizer --tool memcheck python tools/sanity_round2.py"""
import sys
sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import nonlinearsolve_jl_b200 as nls  # noqa: E402

ctx = nls.Context(0)
for f in (nls.Brusselator2D(12), nls.Brusselator2D(15), nls.Brusselator3D(10), nls.Brusselator3D(7)):
    even = (f.N ** (3 if isinstance(f, nls.Brusselator3D) else 2)) % 2 == 0   # the resident engine needs an even cell count
    dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
    u0 = dp.u0(1)
    for alg, kw in ((nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(precs=nls.Multigrid("right"))), {}),
                    (nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(precs=nls.Multigrid("left"), atol=1e-13, rtol=1e-9)), {}),
                    (nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(engine="resident" if even else "auto", orth="mgs")), dict(termination_condition=nls.RelNormSafeBestTerminationMode(norm="l2"))),
                    (nls.TrustRegion(radius_update_scheme=nls.RadiusUpdateSchemes.Bastin), dict(termination_condition=nls.RelTerminationMode(), reltol=1e-10, maxiters=40))):
        sol = nls.solve(nls.NonlinearProblem(f, u0, (3.4, 1.0, 10.0), ctx=ctx), alg, abstol=1e-8, **kw)
        print(type(f).__name__, f.N, alg.name, nls.ReturnCode.name(sol.retcode), sol.stats.nsteps, sol.stats.njvp, sol.resid_inf, flush=True)
    fs = nls.NonlinearFunction(f, sparsity=nls.TracerSparsityDetector())
    sol = nls.solve(nls.NonlinearProblem(fs, u0, (3.4, 1.0, 10.0), ctx=ctx), nls.NewtonRaphson(), abstol=1e-8)
    print(type(f).__name__, f.N, "sparse direct", nls.ReturnCode.name(sol.retcode), sol.stats.nsteps, sol.stats.nfactors, sol.resid_inf, flush=True)

# ---- round-2 late additions: TMA-ring stencil kernels (N = 24: two chunks, both wraps in one copy; N = 26: ragged last chunk),
#      the CSC operator through its row view, ensemble trajectories redone by the general driver
import os  # noqa: E402
rng = np.random.default_rng(0)
for N in (24, 26):
    f = nls.Brusselator3D(N)
    dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
    u, v = dp.u0(1), ctx.to_device(rng.standard_normal(dp.n))
    r, jv, jtv = dp.residual(u), dp.jvp(u, v), dp.vjp(u, v)
    f2, j2 = dp.residual_jvp(u, v)
    print("ring", N, float(np.abs(r.to_host() - f2.to_host()).max()), float(np.abs(jv.to_host() - j2.to_host()).max()), jtv.norm(2), flush=True)
    sol = nls.solve(nls.NonlinearProblem(f, u, (3.4, 1.0, 10.0), ctx=ctx), nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(orth="mgs")), abstol=1e-8)
    print("ring newton", N, nls.ReturnCode.name(sol.retcode), sol.stats.nsteps, sol.stats.njvp, sol.resid_inf, flush=True)
f = nls.Brusselator2D(10)
dp = nls._DeviceProblem(ctx, nls.NonlinearProblem(f, None, (3.4, 1.0, 10.0), ctx=ctx))
sj = nls.SparseJacobian(dp)
nz = sj.fill(dp.u0())
gm = nls.GmresSolver(ctx, dp.n, nls.KrylovJL_GMRES(), atol=0.0, rtol=1e-10)
x, st = gm.solve(("csc", ctx.to_device(sj.colptr, np.int64), ctx.to_device(sj.rowval, np.int64), nz, 1), dp.residual(dp.u0()))
print("csc rows", st.status, st.iters, flush=True)
os.environ["B200_ENS_BASIS_COLUMNS"] = "24"
K, N = 6, 16
P0 = nls.Brusselator2D(N)
d0 = nls._DeviceProblem(ctx, nls.NonlinearProblem(P0, None, (3.4, 1.0, 10.0), ctx=ctx))
u0 = np.tile(d0.u0().to_host(), (K, 1))
cache = nls.EnsembleCache(ctx, N, K, 10.0, nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES(orth="mgs")), abstol=1e-8)
res = cache.solve(ctx.to_device(u0.ravel()), ctx.to_device(np.full(K, 3.4)), ctx.to_device(np.full(K, 1.0)))
print("ens deferred", res.nsuccess, cache.nj.to_host().tolist(), flush=True)
