"""Property tests that pin the oracle's building blocks against independent NumPy / SciPy implementations on random
inputs (the oracle is the parity target of the GPU tests, so it is itself cross-checked as widely as is cheap)."""
import numpy as np
import pytest
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla


@pytest.mark.parametrize("seed", range(4))
def test_gmres_random_dense_vs_numpy(po, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(5, 60))
    A = rng.standard_normal((n, n)) + n * np.eye(n) * rng.uniform(0.2, 1.0)
    b = rng.standard_normal(n)
    xr = np.linalg.solve(A, b)
    for orth in (po.ORTH_MGS, po.ORTH_CGS2):
        x, st = po.gmres(b, dense=A, opts=po.default_gmres_opts(atol=0.0, rtol=1e-12, orth=orth))
        assert st.status == 1 and st.iters <= n
        assert np.abs(x - xr).max() <= 1e-8 * np.abs(xr).max()
        assert abs(st.rnorm0 - np.linalg.norm(b)) <= 1e-12 * np.linalg.norm(b)
    # the Arnoldi relation: the raw Hessenberg reproduces A V_k = V_{k+1} H (checked through the residual norm recursion)
    x, st, h = po.gmres(b, dense=A, opts=po.default_gmres_opts(atol=0.0, rtol=1e-12, orth=po.ORTH_MGS), want_hessenberg=n * (n + 3) // 2)
    assert abs(st.rnorm - np.linalg.norm(b - A @ x)) <= 1e-8 * np.linalg.norm(b)
    # restarted GMRES reaches the same solution
    x2, st2 = po.gmres(b, dense=A, opts=po.default_gmres_opts(atol=0.0, rtol=1e-12, restart=max(2, n // 3), itmax=50 * n))
    assert st2.status == 1 and np.abs(x2 - xr).max() <= 1e-7 * np.abs(xr).max()


@pytest.mark.parametrize("seed", range(4))
def test_lu_random_vs_lapack(po, seed):
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(3, 90))
    A = rng.standard_normal((n, n))
    LU, ipiv, info = po.getrf(A)
    lu_ref, piv_ref = sla.lu_factor(A)
    assert info == 0 and np.array_equal(ipiv - 1, piv_ref)          # LAPACK partial-pivoting sequence, bit-exact indices
    assert np.abs(LU - lu_ref).max() <= 1e-11 * np.abs(lu_ref).max()
    B = rng.standard_normal((n, 3))
    X = po.getrs(LU, ipiv, B)
    assert np.abs(A @ X - B).max() <= 1e-9 * max(1.0, np.abs(X).max()) * np.abs(A).max() * n


@pytest.mark.parametrize("seed", range(4))
def test_coloring_random_patterns_is_a_valid_column_colouring(po, seed):
    rng = np.random.default_rng(200 + seed)
    n = int(rng.integers(8, 120))
    M = sp.random(n, n, density=rng.uniform(0.02, 0.15), random_state=int(seed), format="csc") + sp.eye(n, format="csc")
    M = sp.csc_matrix(M)
    M.sort_indices()
    colptr, rowval = (M.indptr + 1).astype(np.int64), (M.indices + 1).astype(np.int64)
    colors, nc = po.coloring_column(n, colptr, rowval)
    assert colors.min() == 1 and colors.max() == nc
    # structurally orthogonal columns: two columns that share a row never share a colour
    P = (M != 0).astype(np.int64)
    G = (P.T @ P).tocoo()
    for i, j in zip(G.row, G.col):
        if i != j:
            assert colors[i] != colors[j]
    # greedy bound: no more colours than the largest column-intersection degree + 1
    deg = np.diff((P.T @ P).tocsc().indptr) - 1
    assert nc <= deg.max() + 1
    # compressed evaluation recovers the matrix: sum of same-coloured columns, read back through the pattern
    vals = M.toarray()
    for c in range(1, nc + 1):
        comp = vals[:, colors == c].sum(axis=1)
        for j in np.nonzero(colors == c)[0]:
            rows = M.indices[M.indptr[j]:M.indptr[j + 1]]
            assert np.allclose(comp[rows], vals[rows, j])


def test_spmv_and_sparse_jacobian_vs_scipy(po):
    P = po.OracleProblem.bruss3d(5)
    rng = np.random.default_rng(7)
    u = P.u0(1) + 0.1 * rng.standard_normal(P.n)
    cp, rv = P.pattern(1)
    col, nc = po.coloring_column(P.n, cp, rv)
    nz = P.sparse_jac(u, cp, rv, col, nc)
    J = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(P.n, P.n))
    v = rng.standard_normal(P.n)
    assert np.abs(J @ v - P.jvp(u, v)).max() <= 1e-10 * np.abs(P.jvp(u, v)).max()
    assert np.abs(J.T @ v - P.vjp(u, v)).max() <= 1e-10 * np.abs(P.vjp(u, v)).max()
    x = spla.spsolve(J, P.residual(u))
    xo, st = po.gmres(P.residual(u), prob=P, u=u, opts=po.default_gmres_opts(atol=0.0, rtol=1e-12, orth=po.ORTH_CGS2))
    assert st.status == 1 and np.abs(x - xo).max() <= 1e-7 * np.abs(x).max()
