"""Generates the golden fixtures in tests/golden/ with an INDEPENDENT NumPy/SciPy restatement of the reference problem
(not the C oracle, not the CUDA library):

    python tests/golden/make_golden.py

The Julia reference cannot be executed in this environment (no julia binary), so these vectors pin the oracle and the
CUDA path against a second, vectorised implementation of `brusselator_2d_loop` / `init_brusselator_2d`
(lib/NonlinearSolveFirstOrder/test/sparsity_tests__item1.jl:7-50) and against SciPy's sparse LU for the Newton roots.
The reference's own acceptance bounds for these cases (||resid||_inf < 1e-8 at abstol = 1e-8) are asserted here too.
"""
import os

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))


def grid(N):
    return np.arange(N) / (N - 1)  # range(0, stop = 1, length = N)


def forcing(N):
    x = grid(N)[:, None]
    y = grid(N)[None, :]
    return np.where((x - 0.3) ** 2 + (y - 0.6) ** 2 <= 0.1 ** 2, 5.0, 0.0)  # [i, j]


def u0_2d(N):
    x = grid(N)[:, None] * np.ones((1, N))
    y = np.ones((N, 1)) * grid(N)[None, :]
    u = np.empty((N, N, 2))
    u[:, :, 0] = 22.0 * (y * (1 - y)) ** 1.5
    u[:, :, 1] = 27.0 * (x * (1 - x)) ** 1.5
    return u


def lap2(w):
    return np.roll(w, 1, 0) + np.roll(w, -1, 0) + np.roll(w, -1, 1) + np.roll(w, 1, 1) - 4 * w


def residual_2d(u, A=3.4, B=1.0, alpha=10.0):
    N = u.shape[0]
    a = alpha / (1.0 / (N - 1)) ** 2
    U, V = u[:, :, 0], u[:, :, 1]
    du = np.empty_like(u)
    du[:, :, 0] = a * lap2(U) + B + U ** 2 * V - (A + 1) * U + forcing(N)
    du[:, :, 1] = a * lap2(V) + A * U - U ** 2 * V
    return du


def jacobian_2d(u, A=3.4, B=1.0, alpha=10.0):
    """Sparse Jacobian assembled from Kronecker products (column-major vec, species planar)."""
    N = u.shape[0]
    a = alpha / (1.0 / (N - 1)) ** 2
    S = sp.diags([np.ones(N - 1), np.ones(N - 1), [1.0], [1.0]], [1, -1, N - 1, -(N - 1)], shape=(N, N))  # periodic shift sum
    I = sp.identity(N)
    L = sp.kron(I, S) + sp.kron(S, I) - 4 * sp.identity(N * N)  # vec index = i + N j
    U = u[:, :, 0].ravel(order="F")
    V = u[:, :, 1].ravel(order="F")
    J00 = a * L + sp.diags(2 * U * V - (A + 1))
    J01 = sp.diags(U ** 2)
    J10 = sp.diags(A - 2 * U * V)
    J11 = a * L - sp.diags(U ** 2)
    return sp.bmat([[J00, J01], [J10, J11]], format="csc")


def vec(u):
    return u.ravel(order="F")


def newton_splu(N, abstol=1e-8, maxiters=50):
    u = u0_2d(N)
    hist = [np.abs(residual_2d(u)).max()]
    for _ in range(maxiters):
        J = jacobian_2d(u)
        f = vec(residual_2d(u))
        du = spla.splu(J).solve(f)
        u = (vec(u) - du).reshape(u.shape, order="F")
        hist.append(np.abs(residual_2d(u)).max())
        if hist[-1] <= abstol:
            break
    return u, np.array(hist)


def greedy_column_coloring(J):
    """Natural-order greedy distance-1 colouring of the column intersection graph (all column degrees tie for the
    periodic stencil, so LargestFirst == natural order).  Pure-Python, independent of oracle.c."""
    Jc = sp.csc_matrix(J)
    Jr = sp.csr_matrix(J)
    n = Jc.shape[1]
    colors = np.zeros(n, dtype=np.int64)
    for v in range(n):
        forb = set()
        for r in Jc.indices[Jc.indptr[v]:Jc.indptr[v + 1]]:
            for x in Jr.indices[Jr.indptr[r]:Jr.indptr[r + 1]]:
                if colors[x]:
                    forb.add(colors[x])
        c = 1
        while c in forb:
            c += 1
        colors[v] = c
    return colors


def main():
    rng = np.random.default_rng(20260924)
    out = {}
    for N in (8, 32):
        u0 = u0_2d(N)
        f0 = residual_2d(u0)
        v = rng.standard_normal(2 * N * N)
        J = jacobian_2d(u0)
        out["u0_%d" % N] = vec(u0)
        out["f0_%d" % N] = vec(f0)
        out["v_%d" % N] = v
        out["Jv_%d" % N] = J @ v
        out["JTv_%d" % N] = J.T @ v
        # structural pattern (what a global sparsity detector reports): evaluate at u = v = 1 where no entry vanishes
        Js = sp.csc_matrix(jacobian_2d(np.ones((N, N, 2))))
        Js.sort_indices()
        assert Js.nnz == 6 * 2 * N * N and np.all(Js.data != 0)
        out["colptr_%d" % N] = Js.indptr.astype(np.int64) + 1  # 1-based like Julia's SparseMatrixCSC
        out["rowval_%d" % N] = Js.indices.astype(np.int64) + 1
        Jd = np.asarray(J.todense())
        cols = np.repeat(np.arange(2 * N * N), np.diff(Js.indptr))
        out["nzval_%d" % N] = Jd[Js.indices, cols]
        out["colors_%d" % N] = greedy_column_coloring(Js)
        root, hist = newton_splu(N)
        assert hist[-1] < 1e-8, hist  # the reference's acceptance bound (sparsity_tests__item1.jl:54-93)
        out["root_%d" % N] = vec(root)
        out["newton_hist_%d" % N] = hist
        print("N=%d  ||f0||inf=%.16g  newton ||f||inf: %s  ncolors=%d" % (N, np.abs(f0).max(), hist, out["colors_%d" % N].max()))
    # u.^2 .- 2 -> sqrt(2) (rootfind_tests__item1.jl), Newton iterates from u0 = 1
    it = [1.0]
    for _ in range(5):
        it.append(it[-1] - (it[-1] ** 2 - 2.0) / (2 * it[-1]))
    out["quadratic_iterates"] = np.array(it)
    # tridiagonal-quadratic JFNK problem (rootfind_tests__item20.jl): root by dense Newton
    n = 100
    p = rng.random(n)
    T = sp.diags([-np.ones(n - 1), 2 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1]).toarray()
    u = p.copy()
    for _ in range(30):
        F = u + 0.1 * u * (T @ u) - p
        Jd = np.eye(n) + 0.1 * (np.diag(u) @ T + np.diag(T @ u))
        u = u - np.linalg.solve(Jd, F)
    assert np.abs(u + 0.1 * u * (T @ u) - p).max() < 1e-13
    out["tridiag_p"] = p
    out["tridiag_root"] = u
    np.savez_compressed(os.path.join(HERE, "brusselator_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "brusselator_golden.npz"))


if __name__ == "__main__":
    main()
