"""Generates tests/golden/bruss3d_n100_newton.npz: the FULL BASELINE config-3 solve (3D Brusselator N=100, 2*10^6 unknowns,
NewtonRaphson + matrix-free GMRES, Krylov.jl-default MGS, abstol 1e-8, perturbed u0) run once on the CPU oracle port
(oracle/oracle.c, OpenMP) — about 10-20 minutes on 8 cores, 13 GB of Krylov basis.  The GPU test
tests/test_gpu_n100_parity.py compares the CUDA path against this artefact (root relative error <= 1e-6, per-step linear
iteration counts +-2, NLStats equal).  The root is committed as a strided subsample plus norms (the full vector is 16 MB).

    python tests/golden/make_n100_golden.py [N]         (N defaults to 100)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
STRIDE = 97
po.build()
po.set_threads(len(os.sched_getaffinity(0)))
P = po.OracleProblem.bruss3d(N)
u0 = P.u0(1)
t0 = time.time()
u, fu, res, tr = P.newton(u0, po.default_newton_opts(abstol=1e-8, linsolve=po.LINSOLVE_GMRES, gmres_orth=po.ORTH_MGS))
dt = time.time() - t0
out = dict(N=N, stride=STRIDE, retcode=res.retcode, nsteps=res.nsteps, nf=res.nf, njacs=res.njacs, nfactors=res.nfactors, nsolve=res.nsolve,
           njvp=res.njvp, resid_inf=res.resid_inf,
           lin_iters=np.array([t.lin_iters for t in tr], dtype=np.int64), lin_status=np.array([t.lin_status for t in tr], dtype=np.int64),
           fnorm_inf=np.array([t.fnorm_inf for t in tr]), step_norm2=np.array([t.step_norm2 for t in tr]),
           lin_rnorm=np.array([t.lin_rnorm for t in tr]),
           u_sample=u[::STRIDE].copy(), u_norm2=np.linalg.norm(u), u_norminf=np.abs(u).max(), u_sum=u.sum(),
           fu_sample=fu[::STRIDE].copy(), seconds=dt, threads=po.get_threads())
path = os.path.join(ROOT, "tests", "golden", "bruss3d_n%d_newton.npz" % N)
np.savez_compressed(path, **out)
print(json.dumps({k: (v.tolist() if isinstance(v, np.ndarray) and v.size < 16 else (None if isinstance(v, np.ndarray) else v)) for k, v in out.items()}))
