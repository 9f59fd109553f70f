"""ctypes loader for the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  See oracle/oracle.h for what the oracle is pinned against and what is "parity unpinned".
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

# ---- constants mirrored from include/b200newton.h
PROB_BRUSS2D, PROB_BRUSS3D, PROB_QUADRATIC, PROB_TRIDIAG_QUAD = 1, 2, 3, 4
ORTH_MGS, ORTH_CGS, ORTH_CGS2 = 0, 1, 2
LINSOLVE_GMRES, LINSOLVE_DENSE_LU, LINSOLVE_SPARSE_GMRES, LINSOLVE_SPARSE_LU = 0, 1, 2, 3
JVP_EXACT, JVP_FINITE_DIFF = 0, 1
GLOB_NONE, GLOB_TRUST_REGION, GLOB_LINESEARCH = 0, 1, 2
DESCENT_NEWTON, DESCENT_PSEUDO_TRANSIENT, DESCENT_LEVENBERG_MARQUARDT = 0, 1, 2
TR_SIMPLE, TR_NLSOLVE, TR_NOCEDAL_WRIGHT, TR_HEI, TR_YUAN, TR_FAN, TR_BASTIN = range(7)
FORCING_NONE, FORCING_EW2 = 0, 1
TERM_ABS_NORM_SAFE_BEST, TERM_ABS_NORM, TERM_ABS_NORM_SAFE, TERM_NORM, TERM_REL, TERM_REL_NORM, TERM_ABS, TERM_REL_NORM_SAFE, TERM_REL_NORM_SAFE_BEST = range(9)
NORM_INF, NORM_L2 = 0, 1
RC_SUCCESS, RC_MAXITERS, RC_STALLED = 1, 2, 4
LS_SOLVED, LS_MAXITERS, LS_BREAKDOWN, LS_NONFINITE = 1, 2, 3, 4
OP_JVP, OP_JVP_FD, OP_CSC, OP_DENSE = 0, 1, 2, 3
ORDER_NATURAL, ORDER_LARGEST_FIRST = 0, 1


class GmresOpts(C.Structure):
    _fields_ = [("memory", C.c_int32), ("restart", C.c_int32), ("itmax", C.c_int32), ("orth", C.c_int32),
                ("warm_start", C.c_int32), ("engine", C.c_int32), ("check_every", C.c_int32), ("block", C.c_int32),
                ("atol", C.c_double), ("rtol", C.c_double)]


class GmresStats(C.Structure):
    _fields_ = [("status", C.c_int32), ("iters", C.c_int32), ("nmatvec", C.c_int32), ("restarts", C.c_int32),
                ("rnorm0", C.c_double), ("rnorm", C.c_double), ("tol", C.c_double), ("bytes", C.c_double)]


class NewtonOpts(C.Structure):
    _fields_ = [("abstol", C.c_double), ("reltol", C.c_double), ("maxiters", C.c_int32), ("linsolve", C.c_int32),
                ("jvp_mode", C.c_int32), ("globalization", C.c_int32), ("forcing", C.c_int32), ("termination", C.c_int32),
                ("store_trace", C.c_int32), ("fused_step", C.c_int32), ("gmres", GmresOpts),
                ("ew_eta0", C.c_double), ("ew_eta_max", C.c_double), ("ew_gamma", C.c_double), ("ew_alpha", C.c_double),
                ("ew_safeguard_threshold", C.c_double), ("ew_safeguard", C.c_int32), ("max_shrink_times", C.c_int32),
                ("tr_step_threshold", C.c_double), ("tr_shrink_threshold", C.c_double), ("tr_expand_threshold", C.c_double),
                ("tr_shrink_factor", C.c_double), ("tr_expand_factor", C.c_double), ("tr_max_trust_radius", C.c_double),
                ("tr_initial_trust_radius", C.c_double), ("ls_c1", C.c_double), ("ls_rho_hi", C.c_double), ("ls_rho_lo", C.c_double),
                ("ls_maxiters", C.c_int32), ("precond", C.c_int32), ("descent", C.c_int32), ("tr_scheme", C.c_int32),
                ("pt_alpha_initial", C.c_double), ("maxtime", C.c_double), ("term_norm", C.c_int32), ("term_max_stalled_steps", C.c_int32),
                ("lm_damping_initial", C.c_double), ("lm_damping_increase", C.c_double), ("lm_damping_decrease", C.c_double), ("lm_finite_diff_step", C.c_double),
                ("lm_alpha_geodesic", C.c_double), ("lm_b_uphill", C.c_double), ("lm_min_damping_D", C.c_double), ("lm_disable_geodesic", C.c_int32),
                ("reserved0", C.c_int32),
                ("qn_init_jacobian", C.c_int32), ("qn_update_rule", C.c_int32), ("qn_max_resets", C.c_int32), ("qn_threshold", C.c_int32),
                ("qn_reset_tolerance", C.c_double), ("qn_alpha", C.c_double)]


class NewtonResult(C.Structure):
    _fields_ = [("retcode", C.c_int32), ("nsteps", C.c_int32), ("nf", C.c_int32), ("njacs", C.c_int32),
                ("nfactors", C.c_int32), ("nsolve", C.c_int32), ("njvp", C.c_int32), ("ntrace", C.c_int32),
                ("resid_inf", C.c_double), ("bytes", C.c_double)]


class TraceRec(C.Structure):
    _fields_ = [("iter", C.c_int32), ("lin_iters", C.c_int32), ("lin_status", C.c_int32), ("accepted", C.c_int32),
                ("fnorm_inf", C.c_double), ("step_norm2", C.c_double), ("lin_rnorm", C.c_double), ("trust_radius", C.c_double)]


class EnsResult(C.Structure):
    _fields_ = [("nprob", C.c_int32), ("nsuccess", C.c_int32), ("max_nsteps", C.c_int32), ("reserved", C.c_int32),
                ("total_nsteps", C.c_int64), ("total_njvp", C.c_int64), ("worst_resid_inf", C.c_double)]


class Problem(C.Structure):
    _fields_ = [("kind", C.c_int32), ("N", C.c_int32), ("n", C.c_int64), ("A", C.c_double), ("B", C.c_double),
                ("alpha", C.c_double), ("p", C.c_double), ("pvec", C.c_void_p)]


class LinOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("index_base", C.c_int32), ("n", C.c_int64), ("prob", C.POINTER(Problem)),
                ("u", C.c_void_p), ("colptr", C.c_void_p), ("rowval", C.c_void_p), ("nzval", C.c_void_p),
                ("A", C.c_void_p), ("ld", C.c_int64), ("shift", C.c_double)]


def build(force=False):
    """Compile oracle/liboracle.so with the committed Makefile (gcc + OpenMP)."""
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle.h", "Makefile")] + [
        os.path.join(_HERE, "..", "include", "b200newton.h")]
    if not force and os.path.exists(_LIB_PATH) and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src):
        return _LIB_PATH
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, env=env, stdout=subprocess.PIPE,
                   stderr=subprocess.STDOUT)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_get_threads.restype = C.c_int32
        _lib.orc_pattern_nnz.restype = C.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def set_threads(n):
    lib().orc_set_threads(C.c_int32(n))


def get_threads():
    return lib().orc_get_threads()


def default_gmres_opts(**kw):
    o = GmresOpts(memory=20, restart=0, itmax=0, orth=ORTH_MGS, warm_start=0, engine=0, check_every=0, block=0,
                  atol=0.0, rtol=0.0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def default_newton_opts(**kw):
    o = NewtonOpts(abstol=0.0, reltol=0.0, maxiters=1000, linsolve=LINSOLVE_GMRES, jvp_mode=JVP_EXACT, globalization=GLOB_NONE,
                   forcing=FORCING_NONE, termination=TERM_ABS_NORM_SAFE_BEST, store_trace=1, fused_step=1,
                   gmres=default_gmres_opts(), ew_eta0=0.5, ew_eta_max=0.9, ew_gamma=0.9, ew_alpha=2.0,
                   ew_safeguard_threshold=0.1, ew_safeguard=1, max_shrink_times=32)
    for k, v in kw.items():
        if k.startswith("gmres_"):
            setattr(o.gmres, k[len("gmres_"):], v)
        else:
            setattr(o, k, v)
    return o


class OracleProblem:
    """Host-side problem description handed to the oracle (mirrors b200_problem kinds)."""

    def __init__(self, kind, N=0, n=0, A=3.4, B=1.0, alpha=10.0, p=2.0, pvec=None):
        self._pvec = None if pvec is None else _f64(pvec)
        self.c = Problem()
        lib().orc_problem_init(C.byref(self.c), C.c_int32(kind), C.c_int32(N), C.c_int64(n), C.c_double(A), C.c_double(B),
                               C.c_double(alpha), C.c_double(p), _p(self._pvec) if self._pvec is not None else None)
        self.kind, self.N, self.n = kind, N, int(self.c.n)

    @classmethod
    def bruss2d(cls, N, A=3.4, B=1.0, alpha=10.0):
        return cls(PROB_BRUSS2D, N=N, A=A, B=B, alpha=alpha)

    @classmethod
    def bruss3d(cls, N, A=3.4, B=1.0, alpha=10.0):
        return cls(PROB_BRUSS3D, N=N, A=A, B=B, alpha=alpha)

    @classmethod
    def quadratic(cls, n, p=2.0):
        return cls(PROB_QUADRATIC, n=n, p=p)

    @classmethod
    def tridiag_quad(cls, pvec):
        return cls(PROB_TRIDIAG_QUAD, n=len(pvec), pvec=pvec)

    # --- evaluation
    def u0(self, mode=0):
        u = np.empty(self.n)
        lib().orc_u0(C.byref(self.c), C.c_int32(mode), _p(u))
        return u

    def residual(self, u):
        u = _f64(u)
        du = np.empty(self.n)
        lib().orc_residual(C.byref(self.c), _p(u), _p(du))
        return du

    def jvp(self, u, v):
        u, v = _f64(u), _f64(v)
        out = np.empty(self.n)
        lib().orc_jvp(C.byref(self.c), _p(u), _p(v), _p(out))
        return out

    def jvp_fd(self, u, v):
        u, v = _f64(u), _f64(v)
        out = np.empty(self.n)
        lib().orc_jvp_fd(C.byref(self.c), _p(u), _p(v), _p(out))
        return out

    def vjp(self, u, w):
        u, w = _f64(u), _f64(w)
        out = np.empty(self.n)
        lib().orc_vjp(C.byref(self.c), _p(u), _p(w), _p(out))
        return out

    def dense_jac(self, u):
        u = _f64(u)
        J = np.empty((self.n, self.n), order="F")
        lib().orc_dense_jac(C.byref(self.c), _p(u), _p(J), C.c_int64(self.n))
        return J

    def pattern(self, index_base=1):
        nnz = lib().orc_pattern_nnz(C.byref(self.c))
        colptr = np.empty(self.n + 1, dtype=np.int64)
        rowval = np.empty(nnz, dtype=np.int64)
        lib().orc_pattern(C.byref(self.c), C.c_int32(index_base), _p(colptr), _p(rowval))
        return colptr, rowval

    def sparse_jac(self, u, colptr, rowval, colors, ncolors, index_base=1):
        u = _f64(u)
        nz = np.empty(len(rowval))
        lib().orc_sparse_jac_fill(C.byref(self.c), _p(u), _p(colptr), _p(rowval), C.c_int32(index_base), _p(colors),
                                  C.c_int64(ncolors), _p(nz))
        return nz

    def newton(self, u0, opts=None, trace_cap=2048):
        opts = opts or default_newton_opts()
        u0 = _f64(u0)
        u = np.empty(self.n)
        fu = np.empty(self.n)
        res = NewtonResult()
        tr = (TraceRec * trace_cap)()
        lib().orc_newton_solve(C.byref(self.c), _p(u0), C.byref(opts), _p(u), _p(fu), C.byref(res), tr, C.c_int32(trace_cap))
        return u, fu, res, [tr[i] for i in range(res.ntrace)]


def arnoldi_sample(prob, u, k0, count, orth=ORTH_CGS2):
    """Seconds for `count` Arnoldi iterations at basis size k0 (bench.py's CPU baseline sample)."""
    u = _f64(u)
    sec = C.c_double()
    rc = lib().orc_arnoldi_sample(C.byref(prob.c), _p(u), C.c_int32(k0), C.c_int32(count), C.c_int32(orth), C.byref(sec))
    if rc != 0:
        raise MemoryError("orc_arnoldi_sample: allocation failed")
    return sec.value


def coloring_column(n, colptr, rowval, index_base=1, order=ORDER_LARGEST_FIRST):
    colors = np.empty(n, dtype=np.int64)
    nc = C.c_int64(0)
    lib().orc_coloring_column(C.c_int64(n), _p(colptr), _p(rowval), C.c_int32(index_base), C.c_int32(order), _p(colors),
                              C.byref(nc))
    return colors, int(nc.value)


def spmv(n, colptr, rowval, nzval, x, index_base=1, transpose=False):
    x = _f64(x)
    y = np.empty(n)
    fn = lib().orc_spmv_t if transpose else lib().orc_spmv
    fn(C.c_int64(n), _p(colptr), _p(rowval), _p(_f64(nzval)), C.c_int32(index_base), _p(x), _p(y))
    return y


def gmres(b, x0=None, opts=None, prob=None, u=None, fd=False, csc=None, dense=None, want_hessenberg=0):
    """orc_gmres on one of: problem JVP at u | CSC (colptr,rowval,nzval,base) | dense column-major A."""
    b = _f64(b)
    n = len(b)
    opts = opts or default_gmres_opts(atol=1e-8, rtol=1e-8)
    op = LinOp()
    op.n = n
    keep = []
    if prob is not None:
        u = _f64(u)
        keep.append(u)
        op.kind = OP_JVP_FD if fd else OP_JVP
        op.prob = C.pointer(prob.c)
        op.u = _p(u)
    elif csc is not None:
        colptr, rowval, nzval, base = csc
        nzval = _f64(nzval)
        keep += [colptr, rowval, nzval]
        op.kind, op.index_base = OP_CSC, base
        op.colptr, op.rowval, op.nzval = _p(colptr), _p(rowval), _p(nzval)
    else:
        A = np.asfortranarray(dense, dtype=np.float64)
        keep.append(A)
        op.kind, op.A, op.ld = OP_DENSE, _p(A), n
    x = np.zeros(n) if x0 is None else _f64(x0).copy()
    st = GmresStats()
    hraw = np.zeros(max(1, want_hessenberg))
    lib().orc_gmres(C.byref(op), _p(b), _p(x), C.byref(opts), C.byref(st), _p(hraw) if want_hessenberg else None,
                    C.c_int64(want_hessenberg))
    return (x, st, hraw) if want_hessenberg else (x, st)


def getrf(A):
    A = np.asfortranarray(A, dtype=np.float64).copy(order="F")
    n = A.shape[0]
    ipiv = np.empty(n, dtype=np.int64)
    info = C.c_int32(0)
    lib().orc_getrf(C.c_int64(n), _p(A), C.c_int64(n), _p(ipiv), C.byref(info))
    return A, ipiv, info.value


def getrs(LU, ipiv, b):
    b = np.asfortranarray(b, dtype=np.float64).copy(order="F")
    n = LU.shape[0]
    nrhs = 1 if b.ndim == 1 else b.shape[1]
    lib().orc_getrs(C.c_int64(n), C.c_int64(nrhs), _p(LU), C.c_int64(n), _p(ipiv), _p(b), C.c_int64(n))
    return b


def ensemble_solve(N, u0, A, B, alpha=10.0, opts=None):
    opts = opts or default_newton_opts()
    u0, A, B = _f64(u0), _f64(A), _f64(B)
    K = len(A)
    n = 2 * N * N
    u = np.empty((K, n))
    resid = np.empty(K)
    rc = np.empty(K, dtype=np.int32)
    ns = np.empty(K, dtype=np.int32)
    nj = np.empty(K, dtype=np.int32)
    res = EnsResult()
    lib().orc_ensemble_solve(C.c_int32(N), C.c_int32(K), C.c_double(alpha), _p(u0), _p(A), _p(B), C.byref(opts), _p(u),
                             _p(resid), _p(rc), _p(ns), _p(nj), C.byref(res))
    return u, resid, rc, ns, nj, res
