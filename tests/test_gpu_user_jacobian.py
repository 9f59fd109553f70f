"""b3 plug-in point on the GPU: `NonlinearFunction(f!; jac = jac!, jac_prototype = J0)` with user closures that run their own
device code (torch) — dense jac!, sparse jac! writing nzval of a user CSC pattern, and a user pattern without jac! (coloured
JVP sweeps through the callback), all through the whole-solve driver (jacobian.jl:119-125, 241-247)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_user_jac_and_jac_prototype(nls, ctx, po, golden):
    torch = pytest.importorskip("torch")
    pv = golden["tridiag_p"]
    p = torch.tensor(pv, device="cuda")
    n = p.numel()

    def T(x):
        y = 2.0 * x
        y[1:] -= x[:-1]
        y[:-1] -= x[1:]
        return y

    def F(du, u, _p):
        du_t, u_t = torch.as_tensor(du, device="cuda"), torch.as_tensor(u, device="cuda")
        du_t.copy_(u_t + 0.1 * u_t * T(u_t) - p)
        torch.cuda.synchronize()

    def JVP(Jv, v, u, _p):
        Jv_t, v_t, u_t = (torch.as_tensor(a, device="cuda") for a in (Jv, v, u))
        Jv_t.copy_(v_t + 0.1 * (u_t * T(v_t) + v_t * T(u_t)))
        torch.cuda.synchronize()

    # J = I + 0.1 diag(T u) + 0.1 diag(u) T :  J[i,i] = 1 + 0.1 (T u)_i + 0.2 u_i ,  J[i,i±1] = -0.1 u_i
    def JAC_DENSE(J, u, _p):
        J_t, u_t = torch.as_tensor(J, device="cuda").view(n, n), torch.as_tensor(u, device="cuda")   # column-major: J_t[c, r]
        J_t.zero_()
        idx = torch.arange(n, device="cuda")
        J_t[idx, idx] = 1.0 + 0.1 * T(u_t) + 0.2 * u_t
        J_t[idx[1:], idx[:-1]] = -0.1 * u_t[:-1]      # entry (row i, col i+1) lives at J_t[i+1, i]
        J_t[idx[:-1], idx[1:]] = -0.1 * u_t[1:]       # entry (row i, col i-1)
        torch.cuda.synchronize()

    # CSC pattern of the tridiagonal matrix, 1-based like a Julia SparseMatrixCSC
    rows, cols = [], []
    for c in range(n):
        for r in (c - 1, c, c + 1):
            if 0 <= r < n:
                rows.append(r)
                cols.append(c)
    rows, cols = np.array(rows), np.array(cols)
    colptr = np.concatenate([[0], np.cumsum(np.bincount(cols, minlength=n))]).astype(np.int64) + 1
    rowval = rows.astype(np.int64) + 1
    rows_t, cols_t = torch.tensor(rows, device="cuda"), torch.tensor(cols, device="cuda")

    def JAC_NZ(nz, u, _p):
        nz_t, u_t = torch.as_tensor(nz, device="cuda"), torch.as_tensor(u, device="cuda")
        diag = 1.0 + 0.1 * T(u_t) + 0.2 * u_t
        nz_t.copy_(torch.where(rows_t == cols_t, diag[rows_t], -0.1 * u_t[rows_t]))
        torch.cuda.synchronize()

    root = golden["tridiag_root"]
    calls = {"dense": 0, "nz": 0}

    def counted(fn, key):
        def g(*a):
            calls[key] += 1
            return fn(*a)
        return g

    # (a) dense jac! + LU
    f1 = nls.NonlinearFunction(F, jvp=JVP, n=n, jac=counted(JAC_DENSE, "dense"))
    s1 = nls.solve(nls.NonlinearProblem(f1, pv, None, ctx=ctx), nls.NewtonRaphson(), abstol=1e-11)
    assert nls.successful_retcode(s1.retcode) and np.abs(s1.u - root).max() < 1e-8
    assert calls["dense"] == s1.stats.njacs - 1 and s1.stats.nfactors == s1.stats.nsteps   # jac! is what fills J (init call excluded)
    # (b) sparse jac! writing nzval of the user pattern, GMRES on the assembled matrix
    f2 = nls.NonlinearFunction(F, jvp=JVP, n=n, jac=counted(JAC_NZ, "nz"), jac_prototype=(colptr, rowval, 1))
    s2 = nls.solve(nls.NonlinearProblem(f2, pv, None, ctx=ctx), nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES()), abstol=1e-11)
    assert nls.successful_retcode(s2.retcode) and np.abs(s2.u - root).max() < 1e-8 and calls["nz"] == s2.stats.njacs
    # (c) user pattern only: colouring (3 colours for a tridiagonal matrix) + compressed JVP sweeps through the callback
    f3 = nls.NonlinearFunction(F, jvp=JVP, n=n, jac_prototype=(colptr, rowval, 1))
    s3 = nls.solve(nls.NonlinearProblem(f3, pv, None, ctx=ctx), nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES()), abstol=1e-11)
    assert nls.successful_retcode(s3.retcode) and np.abs(s3.u - root).max() < 1e-8 and s3.stats.njacs == s3.stats.nsteps
    # the three routes agree with each other and with the matrix-free solve
    s0 = nls.solve(nls.NonlinearProblem(nls.NonlinearFunction(F, jvp=JVP, n=n), pv, None, ctx=ctx), nls.NewtonRaphson(linsolve=nls.KrylovJL_GMRES()), abstol=1e-11)
    for s in (s1, s2, s3):
        assert np.abs(s.u - s0.u).max() < 1e-8
