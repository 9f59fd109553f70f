"""Generates tests/golden/bruss3d_n100_newton.npz: the FULL BASELINE config-3 solve (3D Brusselator N=100, 2*10^6 unknowns,
NewtonRaphson + matrix-free GMRES, abstol 1e-8, perturbed u0) run once on the CPU oracle port (oracle/oracle.c, OpenMP) — about
half an hour on 8 cores, 13 GB of Krylov basis.

Orthogonalisation: the artefact is generated with the REORTHOGONALISED Gram-Schmidt (orth = cgs2).  With Krylov.jl's default
(plain MGS) the CPU port does not finish this configuration: the linear tolerance the reference inherits
(atol 1e-8 + rtol eps^(4/5) ||r0||, i.e. 3e-13 relative) lies below what MGS-GMRES can attain on this operator in double
precision — the residual estimate of the first Newton step stalls at 5.16e-6 (6.6e-12 ||r0||) from iteration ~800 on and the
unrestarted solver keeps adding basis vectors until memory runs out (observed twice: 54 GB after 3400 iterations; set
ORC_PROGRESS=1 to watch).  DESIGN.md §2 discusses what that means for parity of the MGS runs.  The GPU test
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
ORTH = sys.argv[2] if len(sys.argv) > 2 else "cgs2"
STRIDE = 97
po.build()
po.set_threads(len(os.sched_getaffinity(0)))
P = po.OracleProblem.bruss3d(N)
u0 = P.u0(1)
t0 = time.time()
u, fu, res, tr = P.newton(u0, po.default_newton_opts(abstol=1e-8, linsolve=po.LINSOLVE_GMRES, gmres_orth={"mgs": po.ORTH_MGS, "cgs2": po.ORTH_CGS2}[ORTH]))
dt = time.time() - t0
out = dict(N=N, stride=STRIDE, orth=ORTH, retcode=res.retcode, nsteps=res.nsteps, nf=res.nf, njacs=res.njacs, nfactors=res.nfactors, nsolve=res.nsolve,
           njvp=res.njvp, resid_inf=res.resid_inf,
           lin_iters=np.array([t.lin_iters for t in tr], dtype=np.int64), lin_status=np.array([t.lin_status for t in tr], dtype=np.int64),
           fnorm_inf=np.array([t.fnorm_inf for t in tr]), step_norm2=np.array([t.step_norm2 for t in tr]),
           lin_rnorm=np.array([t.lin_rnorm for t in tr]),
           u_sample=u[::STRIDE].copy(), u_norm2=np.linalg.norm(u), u_norminf=np.abs(u).max(), u_sum=u.sum(),
           fu_sample=fu[::STRIDE].copy(), seconds=dt, threads=po.get_threads())
path = os.path.join(ROOT, "tests", "golden", "bruss3d_n%d_newton.npz" % N)
np.savez_compressed(path, **out)
print(json.dumps({k: (v.tolist() if isinstance(v, np.ndarray) and v.size < 16 else (None if isinstance(v, np.ndarray) else v)) for k, v in out.items()}))
