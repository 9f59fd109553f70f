/*
 * oracle.c — TEST INFRASTRUCTURE ONLY (see oracle.h for the pinning statement).
 *
 * CPU restatement of the reference algorithm on the first-order Newton hot path.  Each function
 * cites the reference file:line it follows (paths relative to the NonlinearSolve.jl tree); where
 * the arithmetic lives in an un-vendored dependency the published algorithm is restated and said so.
 * Compile: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off, strict IEEE, no fast-math).
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PAR_THRESHOLD 32768
static int g_threads = 0;

void orc_set_threads(int32_t nthreads) {
  g_threads = nthreads;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
}
int32_t orc_get_threads(void) {
#ifdef _OPENMP
  return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ BLAS-1 helpers
 * (the reference runs these as Julia broadcasts / LinearAlgebra.dot / norm on Array) */
static double v_dot(int64_t n, const double* x, const double* y) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static) if (n > PAR_THRESHOLD)
  for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
  return s;
}
static double v_nrm2(int64_t n, const double* x) { return sqrt(v_dot(n, x, x)); }
static double v_norminf(int64_t n, const double* x) { /* maximum(abs, x): common_defaults.jl:37 ; NaN-propagating like Julia's maximum */
  double m = 0.0;
  int has_nan = 0;
#pragma omp parallel for reduction(max : m) reduction(| : has_nan) schedule(static) if (n > PAR_THRESHOLD)
  for (int64_t i = 0; i < n; ++i) {
    double a = fabs(x[i]);
    if (a != a) has_nan = 1;
    if (a > m) m = a;
  }
  return has_nan ? NAN : m;
}
static void v_axpy(int64_t n, double a, const double* x, double* y) {
#pragma omp parallel for schedule(static) if (n > PAR_THRESHOLD)
  for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
}
static void v_copy(int64_t n, const double* x, double* y) { memcpy(y, x, (size_t)n * sizeof(double)); }
static void v_scal(int64_t n, double a, double* x) {
#pragma omp parallel for schedule(static) if (n > PAR_THRESHOLD)
  for (int64_t i = 0; i < n; ++i) x[i] *= a;
}
static double v_diffnrm2(int64_t n, const double* x, const double* y) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static) if (n > PAR_THRESHOLD)
  for (int64_t i = 0; i < n; ++i) {
    double d = x[i] - y[i];
    s += d * d;
  }
  return sqrt(s);
}

/* ------------------------------------------------------------------ problems */
void orc_problem_init(orc_problem* p, int32_t kind, int32_t N, int64_t n, double A, double B, double alpha, double pscalar,
                      const double* pvec) {
  p->kind = kind;
  p->N = N;
  p->A = A;
  p->B = B;
  p->alpha = alpha;
  p->p = pscalar;
  p->pvec = pvec;
  if (kind == B200_PROB_BRUSS2D) p->n = 2 * (int64_t)N * N;
  else if (kind == B200_PROB_BRUSS3D) p->n = 2 * (int64_t)N * N * N;
  else p->n = n;
}

/* brusselator_f(x,y) = (((x-0.3)^2 + (y-0.6)^2) <= 0.1^2) * 5.0   sparsity_tests__item1.jl:10 */
static inline double bruss_forcing(double x, double y) {
  double r2 = 0.1 * 0.1; /* Julia 0.1^2 == 0.1*0.1 == 0.010000000000000002 */
  double dx = x - 0.3, dy = y - 0.6;
  return ((dx * dx + dy * dy) <= r2) ? 5.0 : 0.0;
}
static inline double grid_coord(int i, int N) { return (double)i / (double)(N - 1); } /* range(0,1,length=N)[i+1] */
static inline double bruss_a(const orc_problem* p) {
  double dx = 1.0 / (double)(p->N - 1); /* step(range(0, stop=1, length=N)) */
  return p->alpha / (dx * dx);          /* alpha = alpha / dx^2   sparsity_tests__item1.jl:15 */
}

/* init_brusselator_2d  sparsity_tests__item1.jl:40-50 ; 3D extension: SURVEY.md §A.2 */
void orc_u0(const orc_problem* p, int32_t mode, double* u) {
  const int N = p->N;
  if (p->kind == B200_PROB_BRUSS2D) {
    for (int j = 0; j < N; ++j)
      for (int i = 0; i < N; ++i) {
        double x = grid_coord(i, N), y = grid_coord(j, N);
        u[i + (int64_t)N * j] = 22.0 * pow(y * (1.0 - y), 1.5);
        u[i + (int64_t)N * j + (int64_t)N * N] = 27.0 * pow(x * (1.0 - x), 1.5);
      }
  } else if (p->kind == B200_PROB_BRUSS3D) {
    const int64_t N2 = (int64_t)N * N, N3 = N2 * N;
    for (int k = 0; k < N; ++k) {
      double z = grid_coord(k, N);
      double fac = (mode == B200_U0_PERTURBED_Z) ? (1.0 + 0.01 * sin(2.0 * M_PI * z)) : 1.0;
      for (int j = 0; j < N; ++j)
        for (int i = 0; i < N; ++i) {
          double x = grid_coord(i, N), y = grid_coord(j, N);
          int64_t c = i + (int64_t)N * j + N2 * k;
          u[c] = 22.0 * pow(y * (1.0 - y), 1.5) * fac;
          u[c + N3] = 27.0 * pow(x * (1.0 - x), 1.5) * fac;
        }
    }
  } else {
    for (int64_t i = 0; i < p->n; ++i) u[i] = 1.0; /* u0 = ones(n): common_rootfind_testing.jl / BASELINE config 1 */
  }
}

/* brusselator_2d_loop  sparsity_tests__item1.jl:13-36, 0-based, idx = i + N j + N^2 s */
static void bruss2d_residual(const orc_problem* p, const double* u, double* du) {
  const int N = p->N;
  const int64_t N2 = (int64_t)N * N;
  const double a = bruss_a(p), A = p->A, B = p->B;
  const double *U = u, *V = u + N2;
#pragma omp parallel for schedule(static) if (N2 > PAR_THRESHOLD)
  for (int j = 0; j < N; ++j) {
    int jp = (j + 1) % N, jm = (j + N - 1) % N;
    double y = grid_coord(j, N);
    for (int i = 0; i < N; ++i) {
      int ip = (i + 1) % N, im = (i + N - 1) % N;
      double x = grid_coord(i, N);
      int64_t c = i + (int64_t)N * j;
      double uc = U[c], vc = V[c];
      double lapu = U[im + (int64_t)N * j] + U[ip + (int64_t)N * j] + U[i + (int64_t)N * jp] + U[i + (int64_t)N * jm] - 4.0 * uc;
      double lapv = V[im + (int64_t)N * j] + V[ip + (int64_t)N * j] + V[i + (int64_t)N * jp] + V[i + (int64_t)N * jm] - 4.0 * vc;
      du[c] = a * lapu + B + uc * uc * vc - (A + 1.0) * uc + bruss_forcing(x, y);
      du[c + N2] = a * lapv + A * uc - uc * uc * vc;
    }
  }
}
/* exact tangent of the loop above == ForwardDiff pushforward (SciMLJacobianOperators.jl:396-414) */
static void bruss2d_jvp(const orc_problem* p, const double* u, const double* d, double* Jv, int transpose) {
  const int N = p->N;
  const int64_t N2 = (int64_t)N * N;
  const double a = bruss_a(p), A = p->A;
  const double *U = u, *V = u + N2, *DU = d, *DV = d + N2;
#pragma omp parallel for schedule(static) if (N2 > PAR_THRESHOLD)
  for (int j = 0; j < N; ++j) {
    int jp = (j + 1) % N, jm = (j + N - 1) % N;
    for (int i = 0; i < N; ++i) {
      int ip = (i + 1) % N, im = (i + N - 1) % N;
      int64_t c = i + (int64_t)N * j;
      double uc = U[c], vc = V[c], du_ = DU[c], dv_ = DV[c];
      double lapu = DU[im + (int64_t)N * j] + DU[ip + (int64_t)N * j] + DU[i + (int64_t)N * jp] + DU[i + (int64_t)N * jm] - 4.0 * du_;
      double lapv = DV[im + (int64_t)N * j] + DV[ip + (int64_t)N * j] + DV[i + (int64_t)N * jp] + DV[i + (int64_t)N * jm] - 4.0 * dv_;
      double j00 = 2.0 * uc * vc - (A + 1.0), j01 = uc * uc, j10 = A - 2.0 * uc * vc, j11 = -(uc * uc);
      if (!transpose) {
        Jv[c] = a * lapu + j00 * du_ + j01 * dv_;
        Jv[c + N2] = a * lapv + j10 * du_ + j11 * dv_;
      } else {
        Jv[c] = a * lapu + j00 * du_ + j10 * dv_;
        Jv[c + N2] = a * lapv + j01 * du_ + j11 * dv_;
      }
    }
  }
}
/* 3D extension: 7-point periodic Laplacian summed as (2D part) + (z part) */
static void bruss3d_residual(const orc_problem* p, const double* u, double* du) {
  const int N = p->N;
  const int64_t N2 = (int64_t)N * N, N3 = N2 * N;
  const double a = bruss_a(p), A = p->A, B = p->B;
  const double *U = u, *V = u + N3;
#pragma omp parallel for schedule(static) collapse(2) if (N3 > PAR_THRESHOLD)
  for (int k = 0; k < N; ++k)
    for (int j = 0; j < N; ++j) {
      int kp = (k + 1) % N, km = (k + N - 1) % N;
      int jp = (j + 1) % N, jm = (j + N - 1) % N;
      double y = grid_coord(j, N);
      for (int i = 0; i < N; ++i) {
        int ip = (i + 1) % N, im = (i + N - 1) % N;
        double x = grid_coord(i, N);
        int64_t row = (int64_t)N * j + N2 * k, c = i + row;
        double uc = U[c], vc = V[c];
        double lapu = (U[im + row] + U[ip + row] + U[i + (int64_t)N * jp + N2 * k] + U[i + (int64_t)N * jm + N2 * k] - 4.0 * uc) +
                      (U[i + (int64_t)N * j + N2 * kp] + U[i + (int64_t)N * j + N2 * km] - 2.0 * uc);
        double lapv = (V[im + row] + V[ip + row] + V[i + (int64_t)N * jp + N2 * k] + V[i + (int64_t)N * jm + N2 * k] - 4.0 * vc) +
                      (V[i + (int64_t)N * j + N2 * kp] + V[i + (int64_t)N * j + N2 * km] - 2.0 * vc);
        du[c] = a * lapu + B + uc * uc * vc - (A + 1.0) * uc + bruss_forcing(x, y);
        du[c + N3] = a * lapv + A * uc - uc * uc * vc;
      }
    }
}
static void bruss3d_jvp(const orc_problem* p, const double* u, const double* d, double* Jv, int transpose) {
  const int N = p->N;
  const int64_t N2 = (int64_t)N * N, N3 = N2 * N;
  const double a = bruss_a(p), A = p->A;
  const double *U = u, *V = u + N3, *DU = d, *DV = d + N3;
#pragma omp parallel for schedule(static) collapse(2) if (N3 > PAR_THRESHOLD)
  for (int k = 0; k < N; ++k)
    for (int j = 0; j < N; ++j) {
      int kp = (k + 1) % N, km = (k + N - 1) % N;
      int jp = (j + 1) % N, jm = (j + N - 1) % N;
      for (int i = 0; i < N; ++i) {
        int ip = (i + 1) % N, im = (i + N - 1) % N;
        int64_t row = (int64_t)N * j + N2 * k, c = i + row;
        double uc = U[c], vc = V[c], du_ = DU[c], dv_ = DV[c];
        double lapu = (DU[im + row] + DU[ip + row] + DU[i + (int64_t)N * jp + N2 * k] + DU[i + (int64_t)N * jm + N2 * k] - 4.0 * du_) +
                      (DU[i + (int64_t)N * j + N2 * kp] + DU[i + (int64_t)N * j + N2 * km] - 2.0 * du_);
        double lapv = (DV[im + row] + DV[ip + row] + DV[i + (int64_t)N * jp + N2 * k] + DV[i + (int64_t)N * jm + N2 * k] - 4.0 * dv_) +
                      (DV[i + (int64_t)N * j + N2 * kp] + DV[i + (int64_t)N * j + N2 * km] - 2.0 * dv_);
        double j00 = 2.0 * uc * vc - (A + 1.0), j01 = uc * uc, j10 = A - 2.0 * uc * vc, j11 = -(uc * uc);
        if (!transpose) {
          Jv[c] = a * lapu + j00 * du_ + j01 * dv_;
          Jv[c + N3] = a * lapv + j10 * du_ + j11 * dv_;
        } else {
          Jv[c] = a * lapu + j00 * du_ + j10 * dv_;
          Jv[c + N3] = a * lapv + j01 * du_ + j11 * dv_;
        }
      }
    }
}

/* y = T x with T = Tridiagonal(-1, 2, -1)   rootfind_tests__item20.jl:7 */
static inline double tri_apply(int64_t n, const double* x, int64_t i) {
  double s = 2.0 * x[i];
  if (i > 0) s -= x[i - 1];
  if (i + 1 < n) s -= x[i + 1];
  return s;
}

void orc_residual(const orc_problem* p, const double* u, double* du) {
  switch (p->kind) {
    case B200_PROB_BRUSS2D: bruss2d_residual(p, u, du); break;
    case B200_PROB_BRUSS3D: bruss3d_residual(p, u, du); break;
    case B200_PROB_QUADRATIC: /* quadratic_f(u,p) = u .* u .- p  common/common_rootfind_testing.jl:15 */
      for (int64_t i = 0; i < p->n; ++i) du[i] = u[i] * u[i] - p->p;
      break;
    case B200_PROB_TRIDIAG_QUAD: /* F(u,p) = u + 0.1 u .* (T u) - p  rootfind_tests__item20.jl:6-9 */
      for (int64_t i = 0; i < p->n; ++i) du[i] = u[i] + 0.1 * u[i] * tri_apply(p->n, u, i) - p->pvec[i];
      break;
    default: break;
  }
}
void orc_jvp(const orc_problem* p, const double* u, const double* v, double* Jv) {
  switch (p->kind) {
    case B200_PROB_BRUSS2D: bruss2d_jvp(p, u, v, Jv, 0); break;
    case B200_PROB_BRUSS3D: bruss3d_jvp(p, u, v, Jv, 0); break;
    case B200_PROB_QUADRATIC:
      for (int64_t i = 0; i < p->n; ++i) Jv[i] = 2.0 * u[i] * v[i];
      break;
    case B200_PROB_TRIDIAG_QUAD: /* JVP(v,u,p) = v + 0.1 (u .* T v + v .* T u)  rootfind_tests__item20.jl:17-20 */
      for (int64_t i = 0; i < p->n; ++i) Jv[i] = v[i] + 0.1 * (u[i] * tri_apply(p->n, v, i) + v[i] * tri_apply(p->n, u, i));
      break;
    default: break;
  }
}
void orc_vjp(const orc_problem* p, const double* u, const double* w, double* JTw) {
  switch (p->kind) {
    case B200_PROB_BRUSS2D: bruss2d_jvp(p, u, w, JTw, 1); break;
    case B200_PROB_BRUSS3D: bruss3d_jvp(p, u, w, JTw, 1); break;
    case B200_PROB_QUADRATIC:
      for (int64_t i = 0; i < p->n; ++i) JTw[i] = 2.0 * u[i] * w[i];
      break;
    case B200_PROB_TRIDIAG_QUAD: { /* J' w = w + 0.1 (T (u .* w) + (T u) .* w) */
      int64_t n = p->n;
      double* uw = (double*)malloc((size_t)n * sizeof(double));
      for (int64_t i = 0; i < n; ++i) uw[i] = u[i] * w[i];
      for (int64_t i = 0; i < n; ++i) JTw[i] = w[i] + 0.1 * (tri_apply(n, uw, i) + tri_apply(n, u, i) * w[i]);
      free(uw);
    } break;
    default: break;
  }
}
/* FiniteDiff.jl finite_difference_jvp! (forward): eps = max(sqrt(eps)*sqrt|x.v|, sqrt(eps)); (f(x+eps v)-f(x))/eps.
 * External package (FiniteDiff compat "2.24", lib/NonlinearSolveBase/Project.toml) — published algorithm restated, unpinned. */
void orc_jvp_fd(const orc_problem* p, const double* u, const double* v, double* Jv) {
  int64_t n = p->n;
  double relstep = sqrt(DBL_EPSILON);
  double tmp = sqrt(fabs(v_dot(n, u, v)));
  double eps = fmax(relstep * tmp, relstep);
  double* x1 = (double*)malloc((size_t)n * sizeof(double));
  double* f0 = (double*)malloc((size_t)n * sizeof(double));
  for (int64_t i = 0; i < n; ++i) x1[i] = u[i] + eps * v[i];
  orc_residual(p, u, f0);
  orc_residual(p, x1, Jv);
  for (int64_t i = 0; i < n; ++i) Jv[i] = (Jv[i] - f0[i]) / eps;
  free(x1);
  free(f0);
}

/* ------------------------------------------------------------------ sparse helpers */
void orc_spmv(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval, int32_t base, const double* x,
              double* y) {
  for (int64_t i = 0; i < n; ++i) y[i] = 0.0;
  for (int64_t c = 0; c < n; ++c) {
    double xc = x[c];
    for (int64_t k = colptr[c] - base; k < colptr[c + 1] - base; ++k) y[rowval[k] - base] += nzval[k] * xc;
  }
}
void orc_spmv_t(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval, int32_t base, const double* x,
                double* y) {
#pragma omp parallel for schedule(static) if (n > PAR_THRESHOLD)
  for (int64_t c = 0; c < n; ++c) {
    double s = 0.0;
    for (int64_t k = colptr[c] - base; k < colptr[c + 1] - base; ++k) s += nzval[k] * x[rowval[k] - base];
    y[c] = s;
  }
}

/* StatefulJacobianOperator mul!  SciMLJacobianOperators.jl:238-243 ; analytic-jac branch :385-386 */
void orc_linop_apply(const orc_linop* op, const double* x, double* y) {
  switch (op->kind) {
    case ORC_OP_JVP: orc_jvp(op->prob, op->u, x, y); break;
    case ORC_OP_JVP_FD: orc_jvp_fd(op->prob, op->u, x, y); break;
    case ORC_OP_CSC: orc_spmv(op->n, op->colptr, op->rowval, op->nzval, op->index_base, x, y); break;
    case ORC_OP_DENSE: {
      int64_t n = op->n;
      for (int64_t i = 0; i < n; ++i) y[i] = 0.0;
      for (int64_t c = 0; c < n; ++c) {
        double xc = x[c];
        const double* col = op->A + c * op->ld;
        for (int64_t i = 0; i < n; ++i) y[i] += col[i] * xc;
      }
    } break;
    default: break;
  }
  if (op->shift != 0.0) /* dampen_jacobian!!(cache, J::AbstractSciMLOperator, D) = J + D  (descent/damped_newton.jl) */
    for (int64_t i = 0; i < op->n; ++i) y[i] += op->shift * x[i];
}

/* ------------------------------------------------------------------ GMRES
 * Restatement of Krylov.jl `gmres!` as reached through LinearSolve.KrylovJL_GMRES
 * (entered at NonlinearSolveBaseLinearSolveExt.jl:26; tolerances from NonlinearSolveFirstOrder/src/solve.jl:203).
 * External package, no version pinned by the reference: Arnoldi with modified Gram-Schmidt, Givens
 * reflections (Krylov.sym_givens), no reorthogonalisation, no restart unless opts->restart > 0,
 * stop when ||r|| <= atol + rtol ||r0||, breakdown when H[k+1,k] <= eps^(3/4).  orth = CGS / CGS2 are
 * NOT reference modes; they restate the product's faster engines so those can be parity-checked too. */
static void sym_givens(double a, double b, double* c, double* s, double* rho) {
  if (b == 0.0) {
    *c = (a == 0.0) ? 1.0 : (a > 0 ? 1.0 : -1.0);
    *s = 0.0;
    *rho = fabs(a);
  } else if (a == 0.0) {
    *c = 0.0;
    *s = (b > 0 ? 1.0 : -1.0);
    *rho = fabs(b);
  } else if (fabs(b) > fabs(a)) {
    double t = a / b;
    *s = (b > 0 ? 1.0 : -1.0) / sqrt(1.0 + t * t);
    *c = *s * t;
    *rho = b / *s;
  } else {
    double t = b / a;
    *c = (a > 0 ? 1.0 : -1.0) / sqrt(1.0 + t * t);
    *s = *c * t;
    *rho = a / *c;
  }
}

void orc_gmres(const orc_linop* op, const double* b, double* x, const b200_gmres_opts* opts, b200_gmres_stats* stats,
               double* h_raw_out, int64_t h_cap) {
  const int64_t n = op->n;
  int64_t itmax = opts->itmax > 0 ? opts->itmax : n;
  const int restart = opts->restart > 0;
  int64_t mem = restart ? opts->restart : (opts->memory > 0 ? opts->memory : 20);
  if (mem > n) mem = n;
  const double btol = pow(DBL_EPSILON, 0.75);

  double* r0 = (double*)malloc((size_t)n * sizeof(double));
  double* w = (double*)malloc((size_t)n * sizeof(double));
  int64_t vcap = mem + 1;
  double** V = (double**)calloc((size_t)vcap, sizeof(double*));
  int64_t rcap = 64;
  double* R = (double*)malloc((size_t)rcap * sizeof(double));
  double *cs = (double*)malloc((size_t)(vcap + 1) * sizeof(double)), *sn = (double*)malloc((size_t)(vcap + 1) * sizeof(double)),
         *z = (double*)malloc((size_t)(vcap + 1) * sizeof(double));
  int64_t cscap = vcap + 1;

  memset(stats, 0, sizeof(*stats));
  int nmatvec = 0;
  if (opts->warm_start) {
    orc_linop_apply(op, x, w);
    ++nmatvec;
    for (int64_t i = 0; i < n; ++i) r0[i] = b[i] - w[i];
  } else {
    for (int64_t i = 0; i < n; ++i) x[i] = 0.0;
    v_copy(n, b, r0);
  }
  double beta = v_nrm2(n, r0);
  double rNorm = beta;
  const double eps_tol = opts->atol + opts->rtol * rNorm;
  stats->rnorm0 = beta;
  stats->tol = eps_tol;
  int64_t iter = 0;
  int solved = rNorm <= eps_tol, tired = iter >= itmax, breakdown = 0, nonfinite = !(rNorm == rNorm) || isinf(rNorm);
  int64_t hraw_pos = 0;
  int npass = 0;

  while (!(solved || tired || breakdown || nonfinite)) {
    ++npass;
    if (!V[0]) V[0] = (double*)malloc((size_t)n * sizeof(double));
    for (int64_t i = 0; i < n; ++i) V[0][i] = r0[i] / rNorm;
    z[0] = rNorm;
    int64_t nr = 0, k = 0;
    int inner_tired = 0;
    while (!(solved || inner_tired || breakdown || nonfinite)) {
      ++k; /* inner_iter */
      if (k + 1 > vcap) { /* grow basis storage (restart == false: Krylov.jl pushes new vectors past `memory`) */
        int64_t ncap = vcap * 2;
        V = (double**)realloc(V, (size_t)ncap * sizeof(double*));
        for (int64_t i = vcap; i < ncap; ++i) V[i] = NULL;
        vcap = ncap;
      }
      if (k + 2 > cscap) {
        cscap = 2 * cscap + 2;
        cs = (double*)realloc(cs, (size_t)cscap * sizeof(double));
        sn = (double*)realloc(sn, (size_t)cscap * sizeof(double));
        z = (double*)realloc(z, (size_t)cscap * sizeof(double));
      }
      if (nr + k + 1 > rcap) {
        while (nr + k + 1 > rcap) rcap *= 2;
        R = (double*)realloc(R, (size_t)rcap * sizeof(double));
      }
      orc_linop_apply(op, V[k - 1], w);
      ++nmatvec;
      double* Rk = R + nr; /* column k, entries 0..k-1, (and Hbis afterwards) */
      if (opts->orth == B200_ORTH_MGS) {
        for (int64_t i = 0; i < k; ++i) {
          Rk[i] = v_dot(n, V[i], w);
          v_axpy(n, -Rk[i], V[i], w);
        }
      } else { /* classical Gram-Schmidt, optionally twice */
        for (int64_t i = 0; i < k; ++i) Rk[i] = v_dot(n, V[i], w);
        for (int64_t i = 0; i < k; ++i) v_axpy(n, -Rk[i], V[i], w);
        if (opts->orth == B200_ORTH_CGS2) {
          double* corr = (double*)malloc((size_t)k * sizeof(double));
          for (int64_t i = 0; i < k; ++i) corr[i] = v_dot(n, V[i], w);
          for (int64_t i = 0; i < k; ++i) {
            v_axpy(n, -corr[i], V[i], w);
            Rk[i] += corr[i];
          }
          free(corr);
        }
      }
      double Hbis = v_nrm2(n, w);
      if (h_raw_out && hraw_pos + k + 1 <= h_cap) {
        for (int64_t i = 0; i < k; ++i) h_raw_out[hraw_pos + i] = Rk[i];
        h_raw_out[hraw_pos + k] = Hbis;
      }
      hraw_pos += k + 1;
      for (int64_t i = 0; i + 1 < k; ++i) { /* previous reflections */
        double Rtmp = cs[i] * Rk[i] + sn[i] * Rk[i + 1];
        Rk[i + 1] = sn[i] * Rk[i] - cs[i] * Rk[i + 1];
        Rk[i] = Rtmp;
      }
      double rho;
      sym_givens(Rk[k - 1], Hbis, &cs[k - 1], &sn[k - 1], &rho);
      Rk[k - 1] = rho;
      double zeta = sn[k - 1] * z[k - 1];
      z[k - 1] = cs[k - 1] * z[k - 1];
      rNorm = fabs(zeta);
      if ((k % 200) == 0 && getenv("ORC_PROGRESS")) fprintf(stderr, "[oracle]   gmres k %d rnorm %.6e (tol %.6e)\n", (int)k, rNorm, eps_tol);
      nr += k;
      nonfinite = !(rNorm == rNorm) || isinf(rNorm) || !(Hbis == Hbis);
      solved = rNorm <= eps_tol;
      inner_tired = restart ? (k >= mem || iter + k >= itmax) : (iter + k >= itmax);
      breakdown = Hbis <= btol;
      if (!(solved || inner_tired || breakdown || nonfinite)) {
        if (!V[k]) V[k] = (double*)malloc((size_t)n * sizeof(double));
        for (int64_t i = 0; i < n; ++i) V[k][i] = w[i] / Hbis;
        z[k] = zeta;
      }
    }
    /* back substitution R y = z, then x += V y */
    if (!nonfinite) {
      double* y = (double*)malloc((size_t)k * sizeof(double));
      for (int64_t i = k - 1; i >= 0; --i) {
        double s = z[i];
        for (int64_t j = i + 1; j < k; ++j) s -= R[j * (j + 1) / 2 + i] * y[j];
        double d = R[i * (i + 1) / 2 + i];
        y[i] = (d == 0.0) ? 0.0 : s / d;
      }
      for (int64_t i = 0; i < k; ++i) v_axpy(n, y[i], V[i], x);
      free(y);
    }
    iter += k;
    tired = iter >= itmax;
    if (restart && !(solved || tired || breakdown || nonfinite)) {
      orc_linop_apply(op, x, w);
      ++nmatvec;
      for (int64_t i = 0; i < n; ++i) r0[i] = b[i] - w[i];
      rNorm = v_nrm2(n, r0);
    }
    if (!restart) break;
  }
  stats->iters = (int32_t)iter;
  stats->nmatvec = nmatvec;
  stats->restarts = npass > 0 ? npass - 1 : 0;
  stats->rnorm = rNorm;
  stats->status = nonfinite ? B200_LS_NONFINITE : solved ? B200_LS_SOLVED : breakdown ? B200_LS_BREAKDOWN : B200_LS_MAXITERS;
  for (int64_t i = 0; i < vcap; ++i) free(V[i]);
  free(V);
  free(R);
  free(cs);
  free(sn);
  free(z);
  free(r0);
  free(w);
}

/* ------------------------------------------------------------------ Jacobian columns, pattern, colouring */
/* structural + numerical entries of column c (unsorted, may contain duplicates for N < 3) */
static int jac_column(const orc_problem* p, const double* u, int64_t c, int64_t* rows, double* vals) {
  int cnt = 0;
  if (p->kind == B200_PROB_BRUSS2D || p->kind == B200_PROB_BRUSS3D) {
    const int N = p->N;
    const int dim = (p->kind == B200_PROB_BRUSS2D) ? 2 : 3;
    const int64_t N2 = (int64_t)N * N, NC = (dim == 2) ? N2 : N2 * N;
    const double a = u ? bruss_a(p) : 1.0;
    int s = (int)(c / NC);
    int64_t cell = c % NC;
    int i = (int)(cell % N), j = (int)((cell / N) % N), k = (dim == 3) ? (int)(cell / N2) : 0;
    int ip = (i + 1) % N, im = (i + N - 1) % N, jp = (j + 1) % N, jm = (j + N - 1) % N, kp = (k + 1) % N, km = (k + N - 1) % N;
    double uc = u ? u[cell] : 1.0, vc = u ? u[cell + NC] : 1.0;
    double diag_lap = (dim == 2) ? -4.0 * a : -6.0 * a;
    double dself, dcross;
    if (s == 0) {
      dself = 2.0 * uc * vc - (p->A + 1.0); /* d f_u / d u */
      dcross = p->A - 2.0 * uc * vc;        /* d f_v / d u */
    } else {
      dself = -(uc * uc); /* d f_v / d v */
      dcross = uc * uc;   /* d f_u / d v */
    }
    int64_t off = (int64_t)s * NC, offx = (int64_t)(1 - s) * NC;
    rows[cnt] = cell + off; vals[cnt++] = diag_lap + dself;
    rows[cnt] = im + (int64_t)N * j + N2 * k + off; vals[cnt++] = a;
    rows[cnt] = ip + (int64_t)N * j + N2 * k + off; vals[cnt++] = a;
    rows[cnt] = i + (int64_t)N * jm + N2 * k + off; vals[cnt++] = a;
    rows[cnt] = i + (int64_t)N * jp + N2 * k + off; vals[cnt++] = a;
    if (dim == 3) {
      rows[cnt] = i + (int64_t)N * j + N2 * km + off; vals[cnt++] = a;
      rows[cnt] = i + (int64_t)N * j + N2 * kp + off; vals[cnt++] = a;
    }
    rows[cnt] = cell + offx; vals[cnt++] = dcross;
  } else if (p->kind == B200_PROB_QUADRATIC) {
    rows[cnt] = c; vals[cnt++] = u ? 2.0 * u[c] : 1.0;
  } else if (p->kind == B200_PROB_TRIDIAG_QUAD) {
    /* F_i = u_i + 0.1 u_i (2u_i - u_{i-1} - u_{i+1}) - p_i */
    int64_t n = p->n;
    rows[cnt] = c; vals[cnt++] = u ? 1.0 + 0.1 * (tri_apply(n, u, c) + 2.0 * u[c]) : 1.0;
    if (c > 0) { rows[cnt] = c - 1; vals[cnt++] = u ? -0.1 * u[c - 1] : 1.0; }
    if (c + 1 < n) { rows[cnt] = c + 1; vals[cnt++] = u ? -0.1 * u[c + 1] : 1.0; }
  }
  /* sort by row, merge duplicates */
  for (int x = 1; x < cnt; ++x) {
    int64_t r = rows[x]; double v = vals[x]; int y = x - 1;
    while (y >= 0 && rows[y] > r) { rows[y + 1] = rows[y]; vals[y + 1] = vals[y]; --y; }
    rows[y + 1] = r; vals[y + 1] = v;
  }
  int m = 0;
  for (int x = 0; x < cnt; ++x) {
    if (m > 0 && rows[m - 1] == rows[x]) vals[m - 1] += vals[x];
    else { rows[m] = rows[x]; vals[m] = vals[x]; ++m; }
  }
  return m;
}

/* global structural pattern, what TracerSparsityDetector / jac_prototype supply (jacobian.jl:286-353) */
int64_t orc_pattern_nnz(const orc_problem* p) {
  int64_t rows[16]; double vals[16]; int64_t nnz = 0;
  for (int64_t c = 0; c < p->n; ++c) nnz += jac_column(p, NULL, c, rows, vals);
  return nnz;
}
void orc_pattern(const orc_problem* p, int32_t base, int64_t* colptr, int64_t* rowval) {
  int64_t rows[16]; double vals[16]; int64_t pos = 0;
  for (int64_t c = 0; c < p->n; ++c) {
    colptr[c] = pos + base;
    int m = jac_column(p, NULL, c, rows, vals);
    for (int x = 0; x < m; ++x) rowval[pos++] = rows[x] + base;
  }
  colptr[p->n] = pos + base;
}

/* DI.jacobian! with dense AutoForwardDiff (jacobian.jl:244-247): exact derivatives, column-major */
void orc_dense_jac(const orc_problem* p, const double* u, double* J, int64_t ld) {
  int64_t n = p->n, rows[16]; double vals[16];
  for (int64_t c = 0; c < n; ++c) {
    memset(J + c * ld, 0, (size_t)n * sizeof(double));
    int m = jac_column(p, u, c, rows, vals);
    for (int x = 0; x < m; ++x) J[c * ld + rows[x]] = vals[x];
  }
}

/* SparseMatrixColorings 0.4 GreedyColoringAlgorithm(LargestFirst()) with a column partition, as chosen at
 * NonlinearSolveBaseSparseMatrixColoringsExt.jl:13-28.  External package: published algorithm restated (unpinned) —
 * partial distance-2 colouring of the bipartite graph, vertices visited by decreasing column degree (stable sort),
 * smallest admissible colour, colours 1-based. */
void orc_coloring_column(int64_t n, const int64_t* colptr, const int64_t* rowval, int32_t base, int32_t order, int64_t* colors,
                         int64_t* ncolors) {
  int64_t nnz = colptr[n] - base;
  /* row structure (CSR of the pattern) */
  int64_t nrows = n;
  int64_t* rowptr = (int64_t*)calloc((size_t)nrows + 1, sizeof(int64_t));
  int64_t* colidx = (int64_t*)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int64_t));
  for (int64_t k = 0; k < nnz; ++k) rowptr[rowval[k] - base + 1]++;
  for (int64_t r = 0; r < nrows; ++r) rowptr[r + 1] += rowptr[r];
  int64_t* fill = (int64_t*)malloc((size_t)nrows * sizeof(int64_t));
  memcpy(fill, rowptr, (size_t)nrows * sizeof(int64_t));
  for (int64_t c = 0; c < n; ++c)
    for (int64_t k = colptr[c] - base; k < colptr[c + 1] - base; ++k) colidx[fill[rowval[k] - base]++] = c;
  /* vertex order */
  int64_t* perm = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  for (int64_t c = 0; c < n; ++c) perm[c] = c;
  if (order == B200_ORDER_LARGEST_FIRST) { /* stable counting sort by degree, descending */
    int64_t maxdeg = 0;
    for (int64_t c = 0; c < n; ++c) { int64_t d = colptr[c + 1] - colptr[c]; if (d > maxdeg) maxdeg = d; }
    int64_t* cnt = (int64_t*)calloc((size_t)maxdeg + 2, sizeof(int64_t));
    for (int64_t c = 0; c < n; ++c) cnt[maxdeg - (colptr[c + 1] - colptr[c]) + 1]++;
    for (int64_t d = 0; d <= maxdeg; ++d) cnt[d + 1] += cnt[d];
    for (int64_t c = 0; c < n; ++c) perm[cnt[maxdeg - (colptr[c + 1] - colptr[c])]++] = c;
    free(cnt);
  }
  for (int64_t c = 0; c < n; ++c) colors[c] = 0;
  int64_t fcap = 64, maxcolor = 0;
  int64_t* forbidden = (int64_t*)malloc((size_t)fcap * sizeof(int64_t));
  for (int64_t i = 0; i < fcap; ++i) forbidden[i] = -1;
  for (int64_t t = 0; t < n; ++t) {
    int64_t v = perm[t];
    for (int64_t k = colptr[v] - base; k < colptr[v + 1] - base; ++k) {
      int64_t r = rowval[k] - base;
      for (int64_t q = rowptr[r]; q < rowptr[r + 1]; ++q) {
        int64_t x = colidx[q];
        if (colors[x] != 0) forbidden[colors[x]] = v;
      }
    }
    int64_t col = 1;
    while (col < fcap && forbidden[col] == v) ++col;
    if (col + 1 >= fcap) {
      int64_t nf = fcap * 2;
      forbidden = (int64_t*)realloc(forbidden, (size_t)nf * sizeof(int64_t));
      for (int64_t i = fcap; i < nf; ++i) forbidden[i] = -1;
      fcap = nf;
    }
    colors[v] = col;
    if (col > maxcolor) maxcolor = col;
  }
  *ncolors = maxcolor;
  free(forbidden); free(perm); free(fill); free(colidx); free(rowptr);
}

/* DI.jacobian! with AutoSparse(AutoForwardDiff) (jacobian.jl:244-247 after :286-353): one exact JVP per colour
 * with the seed = sum of the unit vectors of that colour's columns, then decompression into nzval. */
void orc_sparse_jac_fill(const orc_problem* p, const double* u, const int64_t* colptr, const int64_t* rowval, int32_t base,
                         const int64_t* colors, int64_t ncolors, double* nzval) {
  int64_t n = p->n;
  double* seed = (double*)malloc((size_t)n * sizeof(double));
  double* comp = (double*)malloc((size_t)n * sizeof(double));
  for (int64_t col = 1; col <= ncolors; ++col) {
    for (int64_t c = 0; c < n; ++c) seed[c] = (colors[c] == col) ? 1.0 : 0.0;
    orc_jvp(p, u, seed, comp);
    for (int64_t c = 0; c < n; ++c)
      if (colors[c] == col)
        for (int64_t k = colptr[c] - base; k < colptr[c + 1] - base; ++k) nzval[k] = comp[rowval[k] - base];
  }
  free(seed); free(comp);
}

/* ------------------------------------------------------------------ dense LU (LAPACK getrf/getrs semantics)
 * The reference reaches LAPACK through LinearSolve's default / LUFactorization (linear_solve.jl:100-117). */
void orc_getrf(int64_t n, double* A, int64_t ld, int64_t* ipiv, int32_t* info) {
  const int64_t nb = 64;
  *info = 0;
  for (int64_t k0 = 0; k0 < n; k0 += nb) {
    int64_t kb = (n - k0 < nb) ? n - k0 : nb;
    /* panel: unblocked left-to-right with partial pivoting, swaps applied across the full row */
    for (int64_t k = k0; k < k0 + kb; ++k) {
      int64_t piv = k; double pmax = fabs(A[k * ld + k]);
      for (int64_t i = k + 1; i < n; ++i) { double a = fabs(A[k * ld + i]); if (a > pmax) { pmax = a; piv = i; } }
      ipiv[k] = piv + 1;
      if (A[k * ld + piv] == 0.0) { if (*info == 0) *info = (int32_t)(k + 1); continue; }
      if (piv != k) for (int64_t c = 0; c < n; ++c) { double t = A[c * ld + k]; A[c * ld + k] = A[c * ld + piv]; A[c * ld + piv] = t; }
      double inv = 1.0 / A[k * ld + k];
      for (int64_t i = k + 1; i < n; ++i) A[k * ld + i] *= inv;
      for (int64_t c = k + 1; c < k0 + kb; ++c) { /* update within panel */
        double akc = A[c * ld + k];
        if (akc != 0.0) for (int64_t i = k + 1; i < n; ++i) A[c * ld + i] -= A[k * ld + i] * akc;
      }
    }
    int64_t k1 = k0 + kb;
    if (k1 >= n) break;
    /* U12 = L11^{-1} A12 ; A22 -= L21 U12 */
#pragma omp parallel for schedule(static)
    for (int64_t c = k1; c < n; ++c) {
      double* col = A + c * ld;
      for (int64_t k = k0; k < k1; ++k) {
        double ukc = col[k];
        if (ukc != 0.0) for (int64_t i = k + 1; i < k1; ++i) col[i] -= A[k * ld + i] * ukc;
      }
      for (int64_t k = k0; k < k1; ++k) {
        double ukc = col[k];
        if (ukc != 0.0) {
          const double* lk = A + k * ld;
          for (int64_t i = k1; i < n; ++i) col[i] -= lk[i] * ukc;
        }
      }
    }
  }
}
void orc_getrs(int64_t n, int64_t nrhs, const double* A, int64_t ld, const int64_t* ipiv, double* B, int64_t ldb) {
  for (int64_t r = 0; r < nrhs; ++r) {
    double* b = B + r * ldb;
    for (int64_t k = 0; k < n; ++k) { int64_t pv = ipiv[k] - 1; if (pv != k) { double t = b[k]; b[k] = b[pv]; b[pv] = t; } }
    for (int64_t k = 0; k < n; ++k) { double bk = b[k]; if (bk != 0.0) for (int64_t i = k + 1; i < n; ++i) b[i] -= A[k * ld + i] * bk; }
    for (int64_t k = n - 1; k >= 0; --k) { b[k] /= A[k * ld + k]; double bk = b[k]; for (int64_t i = 0; i < k; ++i) b[i] -= A[k * ld + i] * bk; }
  }
}

/* ------------------------------------------------------------------ Newton driver
 * GeneralizedFirstOrderAlgorithm: __init (NonlinearSolveFirstOrder/src/solve.jl:140-301), step! (:325-465),
 * driver loop (NonlinearSolveBase/src/solve.jl:360-387, 835-858), NewtonDescent (descent/newton.jl:97-141),
 * Dogleg (descent/dogleg.jl:86-151), GenericTrustRegionScheme/Simple (trust_region.jl:204-258, 396-430, 511-513),
 * EisenstatWalkerForcing2 (eisenstat_walker.jl:42-87), termination (termination_conditions.jl:134-179, 243-336, 414-453). */
typedef struct {
  int mode; double abstol; double best; double initial; int nsteps; int retcode;
  double obj_trace[100]; double step_trace[32]; double* best_u; int64_t n;
} term_cache;

static int term_check(term_cache* tc, const double* fu, const double* u, const double* uprev) {
  int64_t n = tc->n;
  double objective = v_norminf(n, fu);
  if (tc->mode == B200_TERM_ABS_NORM) { /* check_convergence(::AbsNormModes) :366-368 */
    if (objective <= tc->abstol) { tc->retcode = B200_RC_SUCCESS; return 1; }
    return 0;
  }
  if (!isfinite(objective)) { tc->retcode = B200_RC_UNSTABLE; return 1; } /* :256-259 */
  if (tc->mode == B200_TERM_ABS_NORM_SAFE_BEST && objective < tc->best) { /* :269-274 */
    tc->best = objective;
    v_copy(n, u, tc->best_u);
  }
  if (objective <= tc->abstol) { tc->retcode = B200_RC_SUCCESS; return 1; } /* :277-280 */
  tc->nsteps += 1;
  tc->obj_trace[(tc->nsteps - 1) % 100] = objective; /* mod1(nsteps, 100) */
  if (objective <= 3.0 * tc->abstol && tc->nsteps > 100) { /* patience :286-303 */
    double mn = INFINITY, mx = -INFINITY;
    for (int i = 0; i < 100; ++i) { if (tc->obj_trace[i] < mn) mn = tc->obj_trace[i]; if (tc->obj_trace[i] > mx) mx = tc->obj_trace[i]; }
    if (mn < 1.3 * mx) { tc->retcode = B200_RC_STALLED; return 1; }
  }
  /* stall test, max_stalled_steps = 32 for the default mode (:306-331, :384-386) */
  double du_norm = v_diffnrm2(n, u, uprev);
  tc->step_trace[(tc->nsteps - 1) % 32] = du_norm;
  if (tc->nsteps > 32) {
    double mx = 0.0;
    for (int i = 0; i < 32; ++i) if (tc->step_trace[i] > mx) mx = tc->step_trace[i];
    if (mx <= tc->abstol) { tc->retcode = B200_RC_STALLED; return 1; }
  }
  tc->retcode = B200_RC_FAILURE;
  return 0;
}

void orc_newton_solve(const orc_problem* p, const double* u0, const b200_newton_opts* o, double* u_out, double* fu_out,
                      b200_newton_result* res, b200_trace_rec* trace, int32_t trace_cap) {
  const int64_t n = p->n;
  const double abstol = o->abstol > 0 ? o->abstol : 3.0e-13; /* common_defaults.jl:44-48 */
  const double reltol = o->reltol > 0 ? o->reltol : 3.0e-13;
  const int maxiters = o->maxiters > 0 ? o->maxiters : 1000;
  const int tr_on = o->globalization == B200_GLOBALIZATION_TRUST_REGION;
  memset(res, 0, sizeof(*res));

  double* u = (double*)malloc((size_t)n * 8); double* fu = (double*)malloc((size_t)n * 8);
  double* u_cache = (double*)malloc((size_t)n * 8); double* du = (double*)calloc((size_t)n, 8);
  double* xlin = (double*)calloc((size_t)n, 8);
  double *u_trial = NULL, *fu_trial = NULL, *Jdu = NULL, *JTfu = NULL, *du_c = NULL, *c1 = NULL, *c2 = NULL;
  double *c1v = NULL, *c2v = NULL; /* line-search trial iterate / residual */
  if (o->globalization == B200_GLOBALIZATION_LINESEARCH) { c1v = (double*)malloc((size_t)n * 8); c2v = (double*)malloc((size_t)n * 8); }
  if (tr_on) {
    u_trial = (double*)malloc((size_t)n * 8); fu_trial = (double*)malloc((size_t)n * 8); Jdu = (double*)malloc((size_t)n * 8);
    JTfu = (double*)malloc((size_t)n * 8); du_c = (double*)malloc((size_t)n * 8); c1 = (double*)malloc((size_t)n * 8); c2 = (double*)malloc((size_t)n * 8);
  }
  v_copy(n, u0, u);
  orc_residual(p, u, fu); /* evaluate_f(prob,u) in __init does not bump nf (solve.jl:194) */
  v_copy(n, u, u_cache);

  term_cache tc; memset(&tc, 0, sizeof(tc));
  tc.mode = o->termination; tc.abstol = abstol; tc.n = n;
  tc.initial = v_norminf(n, fu);
  tc.best = (tc.mode == B200_TERM_ABS_NORM) ? INFINITY : tc.initial;
  tc.best_u = (tc.mode == B200_TERM_ABS_NORM_SAFE_BEST) ? (double*)malloc((size_t)n * 8) : NULL;
  if (tc.best_u) v_copy(n, u, tc.best_u);

  /* linear solver state */
  b200_gmres_opts gopts = o->gmres;
  if (gopts.atol <= 0) gopts.atol = abstol; /* linsolve_kwargs = (; abstol, reltol)  solve.jl:203 */
  if (gopts.rtol <= 0) gopts.rtol = reltol;
  double* Jdense = NULL; double* LU = NULL; int64_t* ipiv = NULL;
  int64_t *colptr = NULL, *rowval = NULL, *colors = NULL, ncolors = 0; double* nzval = NULL;
  if (o->linsolve == B200_LINSOLVE_DENSE_LU) {
    Jdense = (double*)malloc((size_t)n * n * 8); LU = (double*)malloc((size_t)n * n * 8); ipiv = (int64_t*)malloc((size_t)n * 8);
    res->njacs += 1; /* jac_prototype === nothing: DI.jacobian at init (jacobian.jl:103-117) */
  } else if (o->linsolve == B200_LINSOLVE_SPARSE_GMRES) {
    int64_t nnz = orc_pattern_nnz(p);
    colptr = (int64_t*)malloc((size_t)(n + 1) * 8); rowval = (int64_t*)malloc((size_t)nnz * 8);
    colors = (int64_t*)malloc((size_t)n * 8); nzval = (double*)malloc((size_t)nnz * 8);
    orc_pattern(p, 1, colptr, rowval);
    orc_coloring_column(n, colptr, rowval, 1, B200_ORDER_LARGEST_FIRST, colors, &ncolors);
  }
  orc_linop op; memset(&op, 0, sizeof(op));
  op.n = n; op.prob = p; op.u = u;
  op.kind = (o->linsolve == B200_LINSOLVE_SPARSE_GMRES) ? ORC_OP_CSC : (o->jvp_mode == B200_JVP_FINITE_DIFF ? ORC_OP_JVP_FD : ORC_OP_JVP);
  op.colptr = colptr; op.rowval = rowval; op.nzval = nzval; op.index_base = 1;

  /* trust region init (trust_region.jl:204-258, 330-346): Simple scheme */
  double trust_region = 0, max_tr = 0; int shrink_counter = 0;
  /* per-scheme defaults: trust_region.jl:330-384 (a value of 0 means "default", :320-328) */
  const int sch = o->tr_scheme;
  const double step_thr = o->tr_step_threshold > 0 ? o->tr_step_threshold : (sch == B200_TR_HEI ? 0.0 : sch == B200_TR_YUAN ? 1.0 / 1000 : 1.0 / 10000);
  const double shrink_thr = o->tr_shrink_threshold > 0 ? o->tr_shrink_threshold : (sch == B200_TR_HEI ? 0.0 : sch == B200_TR_NLSOLVE ? 1.0 / 20 : 0.25);
  const double expand_thr = o->tr_expand_threshold > 0 ? o->tr_expand_threshold : (sch == B200_TR_NLSOLVE ? 0.9 : sch == B200_TR_HEI ? 0.0 : 0.75);
  const double shrink_fac = o->tr_shrink_factor > 0 ? o->tr_shrink_factor : (sch == B200_TR_NLSOLVE ? 0.5 : sch == B200_TR_HEI ? 0.0 : 0.25);
  const double expand_fac = o->tr_expand_factor > 0 ? o->tr_expand_factor : 2.0;
  const int max_shrink = o->max_shrink_times > 0 ? o->max_shrink_times : 32;
  /* get_parameters  trust_region.jl:372-379 */
  double tp1 = 0, tp2 = 0, tp3 = 0, tp4 = 0;
  if (sch == B200_TR_NLSOLVE) { tp1 = 0.5; }
  else if (sch == B200_TR_HEI) { tp1 = 5.0; tp2 = 0.1; tp3 = 0.15; tp4 = 0.15; }
  else if (sch == B200_TR_YUAN) { tp1 = 2.0; tp2 = 1.0 / 6; tp3 = 6.0; }
  else if (sch == B200_TR_FAN) { tp1 = 0.1; tp2 = 0.25; tp3 = 12.0; tp4 = 1.0e18; }
  if (tr_on) {
    double fu_norm = v_nrm2(n, fu), u0_norm = v_nrm2(n, u), umin = INFINITY, umax = -INFINITY;
    for (int64_t i = 0; i < n; ++i) { if (u[i] < umin) umin = u[i]; if (u[i] > umax) umax = u[i]; }
    if (o->tr_max_trust_radius > 0) max_tr = o->tr_max_trust_radius;
    else max_tr = (sch == B200_TR_SIMPLE || sch == B200_TR_NOCEDAL_WRIGHT) ? fmax(fu_norm, umax - umin) : INFINITY; /* :330-337 */
    if (o->tr_initial_trust_radius > 0) trust_region = o->tr_initial_trust_radius;
    else if (sch == B200_TR_NLSOLVE) trust_region = u0_norm > 0 ? u0_norm : 1.0; /* :339-346 */
    else if (sch == B200_TR_HEI) trust_region = 1.0;
    else if (sch == B200_TR_FAN) trust_region = pow(fu_norm, 0.99) / 10.0;
    else trust_region = max_tr / 11.0;
    if (sch == B200_TR_YUAN) { orc_vjp(p, u, fu, JTfu); trust_region = tp1 * v_nrm2(n, JTfu); } /* :232-234 */
  }
  /* PseudoTransient = DampedNewtonDescent + SwitchedEvolutionRelaxation (pseudo_transient.jl:37-56, 105-170;
     descent/damped_newton.jl:290-296 `:simple` mode): solve (J + alpha^-1 I) x = fu, du = -x;
     alpha^-1 starts at 1/alpha_initial and is multiplied by ||f_n||_2 / ||f_{n-1}||_2 before every solve (ratio 1 at the first) */
  const int pt_on = o->descent == B200_DESCENT_PSEUDO_TRANSIENT;
  double alpha_inv = pt_on ? 1.0 / (o->pt_alpha_initial > 0 ? o->pt_alpha_initial : 1.0e-3) : 0.0;
  double pt_res_norm = v_nrm2(n, fu);
  /* forcing init (eisenstat_walker.jl:90-98) */
  double eta = o->ew_eta0, rnorm = v_nrm2(n, fu), rnorm_prev = rnorm;

  int retcode = B200_RC_DEFAULT, force_stop = 0, make_new_jacobian = 1, nsteps = 0, ntrace = 0;
  int have_factor = 0;

  while (!force_stop && nsteps < maxiters) {
    int recompute_forced = 0;
  step_again:;
    int new_jacobian;
    (void)recompute_forced;
    if (make_new_jacobian) { /* J = cache.jac_cache(u)  solve.jl:338-344 */
      new_jacobian = 1;
      if (o->linsolve == B200_LINSOLVE_DENSE_LU) { res->njacs += 1; orc_dense_jac(p, u, Jdense, n); }
      else if (o->linsolve == B200_LINSOLVE_SPARSE_GMRES) { res->njacs += 1; orc_sparse_jac_fill(p, u, colptr, rowval, 1, colors, ncolors, nzval); }
    } else new_jacobian = 0;

    if (pt_on && !recompute_forced) {
      const double rn = v_nrm2(n, fu);
      alpha_inv *= rn / pt_res_norm;
      pt_res_norm = rn;
      op.shift = alpha_inv;
    }
    if (o->forcing == B200_FORCING_EW2 && o->linsolve != B200_LINSOLVE_DENSE_LU) { /* pre_step_forcing!  :42-80 */
      if (nsteps == 0) { eta = o->ew_eta0; rnorm = rnorm_prev = v_nrm2(n, fu); }
      else {
        double eta_prev = eta;
        eta = o->ew_gamma * pow(rnorm / rnorm_prev, o->ew_alpha);
        if (o->ew_safeguard) { double sg = o->ew_gamma * pow(eta_prev, o->ew_alpha); if (sg > o->ew_safeguard_threshold && sg > eta) eta = sg; }
        if (eta < 0.0) eta = 0.0;
        if (eta > o->ew_eta_max) eta = o->ew_eta_max;
      }
      gopts.rtol = eta; /* LinearSolve.update_tolerances!(lincache; reltol = eta) */
    }

    /* descent: Newton step  J x = fu ; du = -x   (newton.jl:97-141) */
    int lin_success = 1; b200_gmres_stats gs; memset(&gs, 0, sizeof(gs));
    res->nsolve += 1;
    if (o->linsolve == B200_LINSOLVE_DENSE_LU) {
      int reuse = !new_jacobian;
      if (!(reuse && have_factor)) { /* update_A! for factorisations (NonlinearSolveBaseLinearSolveExt.jl:81-86) */
        res->nfactors += 1; int32_t info; /* copyto!(lincache.A, J) then lu!  (…LinearSolveExt.jl:102-111) */
        memcpy(LU, Jdense, (size_t)n * n * 8);
        if (pt_on) for (int64_t i = 0; i < n; ++i) LU[i * n + i] += alpha_inv; /* dampen_jacobian!!: diagonal += alpha^-1 */
        orc_getrf(n, LU, n, ipiv, &info); have_factor = 1;
        if (info != 0) lin_success = 0;
      }
      v_copy(n, fu, xlin);
      if (lin_success) orc_getrs(n, 1, LU, n, ipiv, xlin, n);
    } else {
      /* `linu` = the du buffer, read as the initial guess only when warm_start is on (descent_buffer_init.jl:20-23) */
      v_copy(n, du, xlin);
      orc_gmres(&op, fu, xlin, &gopts, &gs, NULL, 0);
      res->njvp += gs.nmatvec;
      if (getenv("ORC_PROGRESS")) fprintf(stderr, "[oracle] newton step %d: gmres status %d iters %d rnorm %.6e tol %.6e\n", (int)nsteps, (int)gs.status, (int)gs.iters, gs.rnorm, gs.tol);
      if (gs.status == B200_LS_NONFINITE || gs.status == B200_LS_OUT_OF_MEMORY) lin_success = 0; /* retcode Failure */
    }
    if (!lin_success) { /* solve.jl:367-382 */
      if (new_jacobian) { retcode = B200_RC_INTERNAL_LINSOLVE_FAILED; force_stop = 1; goto step_done; }
      make_new_jacobian = 1; recompute_forced = 1; goto step_again;
    }
    for (int64_t i = 0; i < n; ++i) du[i] = -xlin[i]; /* @. du *= -1   newton.jl:138 */

    double dJJd = NAN; /* extras.duJtJdu */
    if (tr_on) { /* Dogleg  dogleg.jl:86-151 */
      double nrm_newton = v_nrm2(n, du);
      if (!(nrm_newton <= trust_region)) {
        /* Cauchy direction: du_c = -J' fu (steepest.jl:60-80); J is current at u (a rejected step keeps u) */
        orc_vjp(p, u, fu, du_c);
        v_scal(n, -1.0, du_c);
        double l_grad = v_nrm2(n, du_c);
        orc_jvp(p, u, du_c, Jdu);
        double quad = v_dot(n, Jdu, Jdu);
        double d_cauchy = (l_grad * l_grad * l_grad) / quad;
        if (d_cauchy >= trust_region) {
          double lam = trust_region / l_grad;
          for (int64_t i = 0; i < n; ++i) du[i] = lam * du_c[i];
          dJJd = lam * lam * quad;
        } else {
          for (int64_t i = 0; i < n; ++i) { c1[i] = (d_cauchy / l_grad) * du_c[i]; c2[i] = du[i] - c1[i]; }
          double a = v_dot(n, c2, c2), b = 2.0 * v_dot(n, c1, c2), c = d_cauchy * d_cauchy - trust_region * trust_region;
          double aux = fmax(0.0, b * b - 4.0 * a * c);
          double tau = (-b + sqrt(aux)) / (2.0 * a);
          for (int64_t i = 0; i < n; ++i) du[i] = c1[i] + tau * c2[i];
        }
      }
    }
    if (o->forcing == B200_FORCING_EW2 && o->linsolve != B200_LINSOLVE_DENSE_LU) { rnorm_prev = rnorm; rnorm = v_nrm2(n, fu); } /* post_step_forcing! :83-87 */
    make_new_jacobian = 1;
    int accepted = 1;
    double ls_alpha = 1.0;
    if (o->globalization == B200_GLOBALIZATION_LINESEARCH) {
      /* solve.jl:392-408 with LineSearch.jl BackTracking (external package, no version pinned: the LineSearches.jl
       * BackTracking algorithm, cubic interpolation, restated; "parity unpinned"): phi(a) = ||f(u + a du)||^2 / 2,
       * phi'(0) = <fu, J du> by one JVP; accept when phi(a) <= phi(0) + c1 a phi'(0). */
      const double c1 = o->ls_c1 > 0 ? o->ls_c1 : 1e-4, rho_hi = o->ls_rho_hi > 0 ? o->ls_rho_hi : 0.5, rho_lo = o->ls_rho_lo > 0 ? o->ls_rho_lo : 0.1;
      const int ls_max = o->ls_maxiters > 0 ? o->ls_maxiters : 1000;
      double* ut = c1v; double* ft = c2v;
      double nf0 = v_nrm2(n, fu);
      const double phi0 = 0.5 * nf0 * nf0;
      orc_jvp(p, u, du, ft);
      const double dphi0 = v_dot(n, fu, ft);
      double a1 = 1.0, a2 = 1.0, phx0 = phi0, phx1;
#define ORC_PHI(alpha, out) do { for (int64_t i_ = 0; i_ < n; ++i_) ut[i_] = u[i_] + (alpha) * du[i_]; orc_residual(p, ut, ft); res->nf += 1; \
                                 double t_ = v_nrm2(n, ft); (out) = 0.5 * t_ * t_; } while (0)
      ORC_PHI(a1, phx1);
      int itf = 0;
      while (!isfinite(phx1) && itf < 50) { ++itf; a1 = a2; a2 = a1 / 2.0; ORC_PHI(a2, phx1); }
      int it = 0, ls_failed = 0;
      while (phx1 > phi0 + c1 * a2 * dphi0) {
        if (++it > ls_max) { ls_failed = 1; break; }
        double at;
        if (it == 1) at = -(dphi0 * a2 * a2) / (2.0 * (phx1 - phi0 - dphi0 * a2));
        else {
          const double div = 1.0 / (a1 * a1 * a2 * a2 * (a2 - a1));
          const double ca = (a1 * a1 * (phx1 - phi0 - dphi0 * a2) - a2 * a2 * (phx0 - phi0 - dphi0 * a1)) * div;
          const double cb = (-a1 * a1 * a1 * (phx1 - phi0 - dphi0 * a2) + a2 * a2 * a2 * (phx0 - phi0 - dphi0 * a1)) * div;
          if (fabs(ca) <= 1e-14 * fabs(cb) || ca == 0.0) at = dphi0 / (2.0 * cb);
          else { const double dd = fmax(cb * cb - 3.0 * ca * dphi0, 0.0); at = (-cb + sqrt(dd)) / (3.0 * ca); }
        }
        a1 = a2;
        at = fmin(at, a2 * rho_hi);
        a2 = fmax(at, a2 * rho_lo);
        phx0 = phx1;
        ORC_PHI(a2, phx1);
      }
#undef ORC_PHI
      if (ls_failed) { retcode = B200_RC_INTERNAL_LINESEARCH_FAILED; force_stop = 1; }
      v_axpy(n, a2, du, u);                    /* @bb axpy!(alpha, du, u)   solve.jl:403 */
      orc_residual(p, u, fu); res->nf += 1;  /* evaluate_f!               solve.jl:407 */
      accepted = 1; ls_alpha = a2;
    } else if (!tr_on) { /* solve.jl:436-445 */
      v_axpy(n, 1.0, du, u);
      orc_residual(p, u, fu); res->nf += 1;
    } else { /* GenericTrustRegionScheme solve!  trust_region.jl:396-430, 511-513 */
      for (int64_t i = 0; i < n; ++i) u_trial[i] = u[i] + du[i];
      orc_residual(p, u_trial, fu_trial); res->nf += 1;
      if (dJJd != dJJd) { orc_jvp(p, u, du, Jdu); dJJd = v_dot(n, Jdu, Jdu); }
      orc_vjp(p, u, fu, JTfu);
      double nt = v_nrm2(n, fu_trial), nc = v_nrm2(n, fu);
      double num = (nt * nt - nc * nc) / 2.0;
      double denom = v_dot(n, du, JTfu) + dJJd / 2.0;
      double rho = num / denom;
      accepted = rho > step_thr;
      const double dun = v_nrm2(n, du); /* internalnorm(du) */
      if (sch == B200_TR_SIMPLE) { /* trust_region.jl:431-440 */
        if (rho < shrink_thr) { trust_region *= shrink_fac; shrink_counter += 1; }
        else { shrink_counter = 0; if (rho > expand_thr && rho > step_thr) trust_region = expand_fac * trust_region; }
      } else if (sch == B200_TR_NLSOLVE) { /* :441-455 */
        if (rho < shrink_thr) { trust_region *= shrink_fac; shrink_counter += 1; }
        else {
          shrink_counter = 0;
          if (rho >= expand_thr) trust_region = expand_fac * dun;
          else if (rho >= tp1) trust_region = fmax(trust_region, expand_fac * dun);
        }
      } else if (sch == B200_TR_NOCEDAL_WRIGHT) { /* :456-466 */
        if (rho < shrink_thr) { trust_region = shrink_fac * dun; shrink_counter += 1; }
        else { shrink_counter = 0; if (rho > expand_thr && fabs(dun - trust_region) < 1.0e-6 * trust_region) trust_region = expand_fac * trust_region; }
      } else if (sch == B200_TR_HEI) { /* :467-476 with rfunc_adaptive_trust_region :383-391 (r, c2, M, g1, g2, beta) = (rho, shrink_thr, p1, p3, p4, p2) */
        const double r = rho, c2_ = shrink_thr, M = tp1, g1 = tp3, g2 = tp4, beta = tp2;
        const double rf = (r >= c2_) ? (2.0 * (M - 1.0 - g2) * atan(r - c2_) + (1.0 + g2)) / M_PI : (1.0 - g1 - beta) * (exp(r - c2_) + beta / (1.0 - g1 - beta));
        const double tr_new = rf * dun;
        if (tr_new < trust_region) shrink_counter += 1; else shrink_counter = 0;
        trust_region = tr_new;
      } else if (sch == B200_TR_YUAN) { /* :477-490 */
        if (rho < shrink_thr) { tp1 = tp2 * tp1; shrink_counter += 1; }
        else { if (rho >= expand_thr && 2.0 * dun > trust_region) tp1 = tp3 * tp1; shrink_counter = 0; }
        orc_vjp(p, u_trial, fu_trial, JTfu);
        trust_region = tp1 * v_nrm2(n, JTfu);
      } else if (sch == B200_TR_FAN) { /* :491-499 */
        if (rho < shrink_thr) { tp1 *= tp2; shrink_counter += 1; }
        else { shrink_counter = 0; if (rho > expand_thr) tp1 = fmin(tp1 * tp3, tp4); }
        trust_region = tp1 * pow(nt, 0.99);
      }
      if (trust_region > max_tr) trust_region = max_tr;
      if (accepted) { v_copy(n, u_trial, u); v_copy(n, fu_trial, fu); }
      else make_new_jacobian = 0;
      if (shrink_counter > max_shrink) { retcode = B200_RC_SHRINK_THRESHOLD_EXCEEDED; force_stop = 1; }
    }
    if (term_check(&tc, fu, u, u_cache)) { retcode = tc.retcode; force_stop = 1; } /* check_and_update! :414-426 */
    if (trace && ntrace < trace_cap) {
      b200_trace_rec* t = &trace[ntrace];
      t->iter = nsteps + 1; t->lin_iters = gs.iters; t->lin_status = gs.status; t->accepted = accepted;
      t->fnorm_inf = v_norminf(n, fu); t->step_norm2 = v_diffnrm2(n, u, u_cache); t->lin_rnorm = gs.rnorm; t->trust_radius = (o->globalization == B200_GLOBALIZATION_LINESEARCH) ? ls_alpha : pt_on ? 1.0 / alpha_inv : trust_region;
    }
    ++ntrace;
    v_copy(n, u, u_cache);
  step_done:
    res->nsteps += 1; nsteps += 1; /* CommonSolve.step!  NonlinearSolveBase/src/solve.jl:835-858 */
  }
  if (retcode == B200_RC_DEFAULT) retcode = (nsteps >= maxiters) ? B200_RC_MAXITERS : B200_RC_SUCCESS; /* :372-376 */
  /* update_from_termination_cache! for Best modes (:440-453): roll back to the best iterate */
  if (tc.best_u) {
    int same = 1;
    for (int64_t i = 0; i < n; ++i) if (u[i] != tc.best_u[i]) { same = 0; break; }
    if (!same) { v_copy(n, tc.best_u, u); orc_residual(p, u, fu); res->nf += 1; }
  }
  res->retcode = retcode; res->ntrace = ntrace < trace_cap ? ntrace : trace_cap;
  res->resid_inf = v_norminf(n, fu);
  if (u_out) v_copy(n, u, u_out);
  if (fu_out) v_copy(n, fu, fu_out);
  free(u); free(fu); free(u_cache); free(du); free(xlin); free(u_trial); free(fu_trial); free(Jdu); free(JTfu); free(du_c); free(c1); free(c2); free(c1v); free(c2v);
  free(tc.best_u); free(Jdense); free(LU); free(ipiv); free(colptr); free(rowval); free(colors); free(nzval);
}

/* ------------------------------------------------------------------ ensemble
 * EnsembleProblem semantics as used at test/PolyAlgorithms/core_tests__item6.jl:3-20: K independent solves, trajectory m
 * gets its own parameters (remake(prob; p = ...)), results collected in order; EnsembleThreads == omp parallel for. */
void orc_ensemble_solve(int32_t N, int32_t nprob, double alpha, const double* u0, const double* A, const double* B,
                        const b200_newton_opts* opts, double* u_out, double* resid_inf, int32_t* retcodes, int32_t* nsteps,
                        int32_t* njvp, b200_ens_result* result) {
  const int64_t n = 2 * (int64_t)N * N;
#pragma omp parallel for schedule(dynamic, 1)
  for (int32_t m = 0; m < nprob; ++m) {
    orc_problem p; orc_problem_init(&p, B200_PROB_BRUSS2D, N, 0, A[m], B[m], alpha, 0.0, NULL);
    b200_newton_result r;
    double* fu = (double*)malloc((size_t)n * 8);
    orc_newton_solve(&p, u0 + (int64_t)m * n, opts, u_out + (int64_t)m * n, fu, &r, NULL, 0);
    resid_inf[m] = r.resid_inf; retcodes[m] = r.retcode; nsteps[m] = r.nsteps; njvp[m] = r.njvp;
    free(fu);
  }
  if (result) {
    memset(result, 0, sizeof(*result));
    result->nprob = nprob;
    for (int32_t m = 0; m < nprob; ++m) {
      if (retcodes[m] == B200_RC_SUCCESS) result->nsuccess += 1;
      if (nsteps[m] > result->max_nsteps) result->max_nsteps = nsteps[m];
      result->total_nsteps += nsteps[m]; result->total_njvp += njvp[m];
      if (resid_inf[m] > result->worst_resid_inf || resid_inf[m] != resid_inf[m]) result->worst_resid_inf = resid_inf[m];
    }
  }
}

/* ------------------------------------------------------------------ timed CPU sample for bench.py (cpu_baseline / --impl reference)
 * `count` Arnoldi iterations of the reference's inner loop (operator apply + Gram-Schmidt against the current basis +
 * normalise + append) starting from a basis of k0 vectors.  The k0 start vectors are pseudo-random unit vectors rather than
 * a true Krylov basis: the arithmetic and memory traffic per iteration are identical, which is all a throughput sample
 * needs, and it avoids spending minutes of CPU time building the first k0 Krylov vectors.  Returns seconds in *seconds. */
int32_t orc_arnoldi_sample(const orc_problem* p, const double* u, int32_t k0, int32_t count, int32_t orth, double* seconds) {
  const int64_t n = p->n;
  const int64_t kmax = (int64_t)k0 + count + 1;
  double** V = (double**)calloc((size_t)kmax, sizeof(double*));
  double* w = (double*)malloc((size_t)n * 8);
  double* h = (double*)malloc((size_t)kmax * 8);
  if (!V || !w || !h) return -1;
  for (int64_t c = 0; c <= k0; ++c) {
    V[c] = (double*)malloc((size_t)n * 8);
    if (!V[c]) return -1;
    const double s = 1.0 / sqrt((double)n / 3.0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      uint64_t x = (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ull + (uint64_t)(c + 1) * 0xBF58476D1CE4E5B9ull;
      x ^= x >> 31; x *= 0x94D049BB133111EBull; x ^= x >> 29;
      V[c][i] = (((double)(x >> 11) / 9007199254740992.0) * 2.0 - 1.0) * s;
    }
  }
#ifdef _OPENMP
  double t0 = omp_get_wtime();
#else
  double t0 = 0.0;
#endif
  for (int32_t it = 0; it < count; ++it) {
    const int64_t k = (int64_t)k0 + it + 1; /* basis vectors V[0..k-1] */
    orc_jvp(p, u, V[k - 1], w);
    if (orth == B200_ORTH_MGS) {
      for (int64_t i = 0; i < k; ++i) { h[i] = v_dot(n, V[i], w); v_axpy(n, -h[i], V[i], w); }
    } else {
      const int passes = (orth == B200_ORTH_CGS2) ? 2 : 1;
      for (int ps = 0; ps < passes; ++ps) {
        for (int64_t i = 0; i < k; ++i) h[i] = v_dot(n, V[i], w);
        for (int64_t i = 0; i < k; ++i) v_axpy(n, -h[i], V[i], w);
      }
    }
    const double nw = v_nrm2(n, w);
    V[k] = (double*)malloc((size_t)n * 8);
    if (!V[k]) return -1;
    const double inv = nw > 0 ? 1.0 / nw : 0.0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) V[k][i] = w[i] * inv;
  }
#ifdef _OPENMP
  *seconds = omp_get_wtime() - t0;
#else
  *seconds = 0.0;
#endif
  for (int64_t c = 0; c < kmax; ++c) free(V[c]);
  free(V); free(w); free(h);
  return 0;
}
