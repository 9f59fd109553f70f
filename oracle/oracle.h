/*
 * oracle.h — TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C + OpenMP) of the reference
 * algorithm for the first-order Newton hot path of NonlinearSolve.jl.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product (libb200newton.so) never links, imports or calls anything in oracle/.
 *
 * PARITY PINNING.  The Julia reference cannot run in this environment (no julia binary, no
 * network) and none of its arithmetic dependencies (Krylov.jl, LinearSolve.jl, ForwardDiff,
 * SparseMatrixColorings) are vendored under /root/reference.  The oracle is therefore pinned
 * against what the reference's own tests hold for this path:
 *   - sparsity_tests__item1.jl:54-93   2D Brusselator N=32, ||resid||inf < 1e-8 at abstol=1e-8
 *   - rootfind_tests__item20.jl:32-54  custom jvp + GMRES, tridiagonal quadratic, < 1e-6 at abstol=1e-13
 *   - rootfind_tests__item1.jl:8-49    u.*u .- 2 -> sqrt(2), err < 1e-9
 *   - core_tests__item3.jl:40-58       JVP/VJP vs analytic Jacobian, atol 1e-5
 *   - operator_jacobian.jl:17-30       linear tridiagonal, sol.u ~ W\b
 *   - core_tests__item6.jl:14-20       ensemble: all successful retcodes
 *   - rootfind_tests__item5.jl:29-33   PseudoTransient(alpha_initial = 10) on u.*u .- 2, err < 1e-9
 *   - rootfind tests over RadiusUpdateSchemes: every scheme solves u.*u .- 2 to err < 1e-9
 * and against an independent NumPy/SciPy restatement (tests/golden/make_golden.py, fixtures in
 * tests/golden/).  GMRES iteration counts, Hessenberg entries, sparsity index arrays, colour
 * vectors and NLStats on the Brusselator are **parity unpinned** in the reference itself (no
 * reference test reads them); for those the oracle restates the published algorithms of
 * Krylov.jl (gmres.jl: MGS Arnoldi + Givens reflections, no restart) and SparseMatrixColorings
 * 0.4 (GreedyColoringAlgorithm(LargestFirst()), column colouring) and IS the parity target.
 * Also unpinned (external packages, no version recorded): LineSearch.jl / LineSearches.jl BackTracking (cubic
 * interpolation) behind `linesearch = BackTracking()`; Krylov.jl's left / right preconditioning (M, N with ldiv = false).
 */
#ifndef ORACLE_H
#define ORACLE_H
#include "../include/b200newton.h" /* option/result structs and enums only */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_problem {
  int32_t kind; /* B200_PROB_* (no callback kind) */
  int32_t N;
  int64_t n;
  double A, B, alpha;
  double p;            /* quadratic: f = u.^2 .- p */
  const double* pvec;  /* tridiag_quad: f = u + 0.1 u.*(T u) - pvec */
} orc_problem;

enum { ORC_OP_JVP = 0, ORC_OP_JVP_FD = 1, ORC_OP_CSC = 2, ORC_OP_DENSE = 3 };
typedef struct orc_linop {
  int32_t kind;
  int32_t index_base;
  int64_t n;
  const orc_problem* prob;
  const double* u;
  const int64_t* colptr;
  const int64_t* rowval;
  const double* nzval;
  const double* A;
  int64_t ld;
  double shift; /* operator is A + shift I (DampedNewtonDescent: J + D) */
} orc_linop;

void orc_set_threads(int32_t nthreads);
int32_t orc_get_threads(void);

void orc_problem_init(orc_problem* p, int32_t kind, int32_t N, int64_t n, double A, double B, double alpha, double pscalar,
                      const double* pvec);
void orc_u0(const orc_problem* p, int32_t mode, double* u);
void orc_residual(const orc_problem* p, const double* u, double* du);
void orc_jvp(const orc_problem* p, const double* u, const double* v, double* Jv);
void orc_jvp_fd(const orc_problem* p, const double* u, const double* v, double* Jv);
void orc_vjp(const orc_problem* p, const double* u, const double* w, double* JTw);

void orc_linop_apply(const orc_linop* op, const double* x, double* y);
/* hess_out (optional): packed upper-triangular R after Givens, column k at offset k(k+1)/2; h_raw_out (optional): raw
 * Hessenberg columns, column k (0-based) holds k+2 entries at offset k(k+3)/2. */
void orc_gmres(const orc_linop* op, const double* b, double* x, const b200_gmres_opts* opts, b200_gmres_stats* stats,
               double* h_raw_out, int64_t h_cap);

void orc_dense_jac(const orc_problem* p, const double* u, double* J, int64_t ld);
void orc_getrf(int64_t n, double* A, int64_t ld, int64_t* ipiv, int32_t* info);
void orc_getrs(int64_t n, int64_t nrhs, const double* A, int64_t ld, const int64_t* ipiv, double* B, int64_t ldb);

int64_t orc_pattern_nnz(const orc_problem* p);
void orc_pattern(const orc_problem* p, int32_t index_base, int64_t* colptr, int64_t* rowval);
void orc_coloring_column(int64_t n, const int64_t* colptr, const int64_t* rowval, int32_t index_base, int32_t order,
                         int64_t* colors, int64_t* ncolors);
void orc_sparse_jac_fill(const orc_problem* p, const double* u, const int64_t* colptr, const int64_t* rowval,
                         int32_t index_base, const int64_t* colors, int64_t ncolors, double* nzval);
void orc_spmv(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval, int32_t index_base,
              const double* x, double* y);
void orc_spmv_t(int64_t n, const int64_t* colptr, const int64_t* rowval, const double* nzval, int32_t index_base,
                const double* x, double* y);

void orc_newton_solve(const orc_problem* p, const double* u0, const b200_newton_opts* opts, double* u_out, double* fu_out,
                      b200_newton_result* result, b200_trace_rec* trace, int32_t trace_cap);

void orc_ensemble_solve(int32_t N, int32_t nprob, double alpha, const double* u0, const double* A, const double* B,
                        const b200_newton_opts* opts, double* u_out, double* resid_inf, int32_t* retcodes, int32_t* nsteps,
                        int32_t* njvp, b200_ens_result* result);

/* bench helper: time `count` Arnoldi iterations starting from a k0-vector basis (see oracle.c) */
int32_t orc_arnoldi_sample(const orc_problem* p, const double* u, int32_t k0, int32_t count, int32_t orth, double* seconds);

#ifdef __cplusplus
}
#endif
#endif
