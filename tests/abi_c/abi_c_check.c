/* abi_c_check.c — a plain-C host of libb200newton.so: the closest executable stand-in for the Julia `@ccall` glue
 * (julia/B200Newton) in an image without a Julia runtime.  Compiled by gcc against include/b200newton.h only.
 *
 *   1. BASELINE config 1 (reference plumbing): f(u) = u.^2 .- p, u0 = ones(1000), p = 2, NewtonRaphson() with the dense-LU
 *      default and with KrylovJL_GMRES; expected root sqrt(2) to 1e-9, iterates 1 -> 1.5 -> 17/12 -> ... (4 steps at abstol 1e-9)
 *      (lib/NonlinearSolveFirstOrder/test/rootfind_tests__item1.jl:8-49).
 *   2. The ensemble path through the C-ABI collective: K trajectories of the 2D Brusselator sharded
 *      over every visible GPU from ONE process (b200_nccl_init_all), solved per device, gathered with b200_ens_allgather and
 *      reduced with b200_ens_allreduce_stats; checked bit for bit against device 0 solving all K alone.
 *      (test/PolyAlgorithms/core_tests__item6.jl:3-20: every trajectory successful, results in trajectory order.)
 * Prints ABI_C_OK and exits 0 on success.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200newton.h"

#define CHECK(ctx, call)                                                                              \
  do {                                                                                                \
    int32_t s_ = (call);                                                                              \
    if (s_ != B200_OK) {                                                                              \
      fprintf(stderr, "%s:%d: %s -> %d (%s)\n", __FILE__, __LINE__, #call, (int)s_, (ctx) ? b200_last_error(ctx) : ""); \
      return 1;                                                                                       \
    }                                                                                                 \
  } while (0)
#define REQUIRE(cond)                                                                                 \
  do {                                                                                                \
    if (!(cond)) { fprintf(stderr, "%s:%d: requirement failed: %s\n", __FILE__, __LINE__, #cond); return 1; } \
  } while (0)

static int config1(b200_ctx* ctx, int linsolve) {
  const int64_t n = 1000;
  b200_problem* prob = NULL;
  b200_newton* nw = NULL;
  b200_newton_opts o;
  b200_newton_result r;
  double* u0 = NULL;
  double* u_dev = NULL;
  double* host = (double*)malloc(sizeof(double) * n);
  CHECK(ctx, b200_problem_create_quadratic(ctx, n, 2.0, &prob));
  b200_newton_opts_default(&o);
  o.linsolve = linsolve;
  o.abstol = 1e-9; /* the reference test's abstol */
  CHECK(ctx, b200_newton_create(prob, &o, &nw));
  CHECK(ctx, b200_malloc(ctx, sizeof(double) * n, (void**)&u0));
  CHECK(ctx, b200_fill(ctx, n, 1.0, u0));
  CHECK(ctx, b200_newton_reinit(nw, u0));
  /* step!(cache) twice: 1 -> 1.5 -> 17/12, then solve!(cache) */
  int32_t done = 0;
  CHECK(ctx, b200_newton_step(nw, &done));
  CHECK(ctx, b200_newton_u(nw, &u_dev));
  CHECK(ctx, b200_memcpy_d2h(ctx, host, u_dev, sizeof(double) * n));
  REQUIRE(!done && fabs(host[0] - 1.5) < 1e-15 && fabs(host[n - 1] - 1.5) < 1e-15);
  CHECK(ctx, b200_newton_step(nw, &done));
  CHECK(ctx, b200_newton_u(nw, &u_dev));
  CHECK(ctx, b200_memcpy_d2h(ctx, host, u_dev, sizeof(double) * n));
  REQUIRE(fabs(host[7] - 17.0 / 12.0) < 1e-14);
  CHECK(ctx, b200_newton_solve(nw, &r));
  CHECK(ctx, b200_newton_u(nw, &u_dev));
  CHECK(ctx, b200_memcpy_d2h(ctx, host, u_dev, sizeof(double) * n));
  double err = 0.0;
  for (int64_t i = 0; i < n; ++i) err = fmax(err, fabs(host[i] - sqrt(2.0)));
  printf("config1 linsolve=%d: retcode=%d nsteps=%d nf=%d njacs=%d nfactors=%d nsolve=%d err=%.3e resid_inf=%.3e\n", linsolve, r.retcode, r.nsteps, r.nf,
         r.njacs, r.nfactors, r.nsolve, err, r.resid_inf);
  /* 1 -> 1.5 -> 1.41667 -> 1.4142157 -> 1.41421356237: ||f||_inf = 4.5e-12 <= abstol after the 4th step */
  REQUIRE(r.retcode == B200_RC_SUCCESS && err < 1e-9 && r.nsteps == 4 && r.nsolve == 4 && r.nf == 4);
  if (linsolve == B200_LINSOLVE_DENSE_LU) REQUIRE(r.nfactors == 4 && r.njacs == 5);
  CHECK(ctx, b200_free(ctx, u0));
  CHECK(ctx, b200_newton_destroy(nw));
  CHECK(ctx, b200_problem_destroy(prob));
  free(host);
  return 0;
}

/* trajectory m: A_m = 3.4 + 0.1 (m mod 64)/64, B_m = 1 + 0.05 floor(m/64)/128  (BASELINE config 5) */
static void ens_params(int K, int m0, double* A, double* B) {
  for (int m = 0; m < K; ++m) {
    const int g = m0 + m;
    A[m] = 3.4 + 0.1 * (g % 64) / 64.0;
    B[m] = 1.0 + 0.05 * (g / 64) / 128.0;
  }
}

static int ensemble(int ndev) {
  enum { N = 8, KPER = 6 };
  const int64_t n = 2 * N * N;
  const int K = KPER * ndev;
  b200_ctx* ctx[8];
  b200_comm* comm[8];
  b200_ensemble* ens[8];
  double *u0[8], *A[8], *B[8], *uloc[8], *uall[8];
  b200_ens_result local[8], global[8];
  b200_newton_opts o;
  b200_newton_opts_default(&o);
  o.abstol = 1e-8;
  double* hu0 = (double*)malloc(sizeof(double) * n * K);
  double* hA = (double*)malloc(sizeof(double) * K);
  double* hB = (double*)malloc(sizeof(double) * K);
  for (int d = 0; d < ndev; ++d) CHECK(NULL, b200_ctx_create(d, NULL, &ctx[d]));
  { /* reference initial condition of the 2D Brusselator, replicated per trajectory */
    b200_problem* p = NULL;
    double* tmp = NULL;
    CHECK(ctx[0], b200_problem_create_bruss2d(ctx[0], N, 3.4, 1.0, 10.0, &p));
    CHECK(ctx[0], b200_malloc(ctx[0], sizeof(double) * n, (void**)&tmp));
    CHECK(ctx[0], b200_problem_u0(p, B200_U0_REFERENCE, tmp));
    CHECK(ctx[0], b200_memcpy_d2h(ctx[0], hu0, tmp, sizeof(double) * n));
    for (int m = 1; m < K; ++m) memcpy(hu0 + (size_t)m * n, hu0, sizeof(double) * n);
    CHECK(ctx[0], b200_free(ctx[0], tmp));
    CHECK(ctx[0], b200_problem_destroy(p));
  }
  CHECK(ctx[0], b200_nccl_init_all(ctx, ndev, comm));
  for (int d = 0; d < ndev; ++d) {
    ens_params(KPER, d * KPER, hA, hB);
    CHECK(ctx[d], b200_ens_create(ctx[d], N, KPER, 10.0, &o, &ens[d]));
    CHECK(ctx[d], b200_malloc(ctx[d], sizeof(double) * n * KPER, (void**)&u0[d]));
    CHECK(ctx[d], b200_malloc(ctx[d], sizeof(double) * KPER, (void**)&A[d]));
    CHECK(ctx[d], b200_malloc(ctx[d], sizeof(double) * KPER, (void**)&B[d]));
    CHECK(ctx[d], b200_malloc(ctx[d], sizeof(double) * n * KPER, (void**)&uloc[d]));
    CHECK(ctx[d], b200_malloc(ctx[d], sizeof(double) * n * K, (void**)&uall[d]));
    CHECK(ctx[d], b200_memcpy_h2d(ctx[d], u0[d], hu0, sizeof(double) * n * KPER));
    CHECK(ctx[d], b200_memcpy_h2d(ctx[d], A[d], hA, sizeof(double) * KPER));
    CHECK(ctx[d], b200_memcpy_h2d(ctx[d], B[d], hB, sizeof(double) * KPER));
  }
  for (int d = 0; d < ndev; ++d)
    CHECK(ctx[d], b200_ens_solve(ens[d], u0[d], A[d], B[d], uloc[d], NULL, NULL, NULL, NULL, &local[d]));
  CHECK(ctx[0], b200_nccl_group_start());
  for (int d = 0; d < ndev; ++d) {
    CHECK(ctx[d], b200_ens_allgather(comm[d], uloc[d], n * KPER, uall[d]));
    CHECK(ctx[d], b200_ens_allreduce_stats_begin(comm[d], &local[d]));
  }
  CHECK(ctx[0], b200_nccl_group_end());
  for (int d = 0; d < ndev; ++d) CHECK(ctx[d], b200_ens_allreduce_stats_finish(comm[d], &global[d]));
  /* reference: device 0 solves all K trajectories alone */
  b200_ensemble* eall = NULL;
  double *u0a = NULL, *Aa = NULL, *Ba = NULL, *ua = NULL;
  b200_ens_result rall;
  ens_params(K, 0, hA, hB);
  CHECK(ctx[0], b200_ens_create(ctx[0], N, K, 10.0, &o, &eall));
  CHECK(ctx[0], b200_malloc(ctx[0], sizeof(double) * n * K, (void**)&u0a));
  CHECK(ctx[0], b200_malloc(ctx[0], sizeof(double) * K, (void**)&Aa));
  CHECK(ctx[0], b200_malloc(ctx[0], sizeof(double) * K, (void**)&Ba));
  CHECK(ctx[0], b200_malloc(ctx[0], sizeof(double) * n * K, (void**)&ua));
  CHECK(ctx[0], b200_memcpy_h2d(ctx[0], u0a, hu0, sizeof(double) * n * K));
  CHECK(ctx[0], b200_memcpy_h2d(ctx[0], Aa, hA, sizeof(double) * K));
  CHECK(ctx[0], b200_memcpy_h2d(ctx[0], Ba, hB, sizeof(double) * K));
  CHECK(ctx[0], b200_ens_solve(eall, u0a, Aa, Ba, ua, NULL, NULL, NULL, NULL, &rall));
  double* ref = (double*)malloc(sizeof(double) * n * K);
  double* got = (double*)malloc(sizeof(double) * n * K);
  CHECK(ctx[0], b200_memcpy_d2h(ctx[0], ref, ua, sizeof(double) * n * K));
  for (int d = 0; d < ndev; ++d) {
    CHECK(ctx[d], b200_memcpy_d2h(ctx[d], got, uall[d], sizeof(double) * n * K));
    REQUIRE(memcmp(ref, got, sizeof(double) * n * K) == 0); /* trajectory order, bit for bit, on every rank */
    REQUIRE(global[d].nprob == K && global[d].nsuccess == K && global[d].total_nsteps == rall.total_nsteps && global[d].total_njvp == rall.total_njvp &&
            global[d].max_nsteps == rall.max_nsteps && global[d].worst_resid_inf == rall.worst_resid_inf);
  }
  printf("ensemble: %d devices x %d trajectories gathered through b200_ens_allgather: nsuccess=%d/%d total_nsteps=%lld total_njvp=%lld worst_resid=%.3e\n", ndev,
         KPER, global[0].nsuccess, K, (long long)global[0].total_nsteps, (long long)global[0].total_njvp, global[0].worst_resid_inf);
  REQUIRE(rall.nsuccess == K && rall.worst_resid_inf < 1e-8);
  for (int d = 0; d < ndev; ++d) {
    CHECK(ctx[d], b200_nccl_destroy(comm[d]));
    CHECK(ctx[d], b200_ens_destroy(ens[d]));
    CHECK(ctx[d], b200_free(ctx[d], u0[d])); CHECK(ctx[d], b200_free(ctx[d], A[d])); CHECK(ctx[d], b200_free(ctx[d], B[d]));
    CHECK(ctx[d], b200_free(ctx[d], uloc[d])); CHECK(ctx[d], b200_free(ctx[d], uall[d]));
  }
  CHECK(ctx[0], b200_ens_destroy(eall));
  CHECK(ctx[0], b200_free(ctx[0], u0a)); CHECK(ctx[0], b200_free(ctx[0], Aa)); CHECK(ctx[0], b200_free(ctx[0], Ba)); CHECK(ctx[0], b200_free(ctx[0], ua));
  for (int d = 0; d < ndev; ++d) CHECK(NULL, b200_ctx_destroy(ctx[d]));
  free(hu0); free(hA); free(hB); free(ref); free(got);
  return 0;
}

int main(int argc, char** argv) {
  int32_t ndev = 0;
  if (b200_device_count(&ndev) != B200_OK || ndev < 1) {
    fprintf(stderr, "no CUDA device: the library has no CPU path (b200_device_count -> %d devices)\n", (int)ndev);
    return 3;
  }
  if (argc > 1) ndev = atoi(argv[1]) < ndev ? atoi(argv[1]) : ndev;
  if (ndev > 8) ndev = 8;
  b200_ctx* ctx = NULL;
  CHECK(NULL, b200_ctx_create(0, NULL, &ctx));
  printf("libb200newton version %d, %d device(s)\n", (int)b200_version(), (int)ndev);
  if (config1(ctx, B200_LINSOLVE_DENSE_LU)) return 1;
  if (config1(ctx, B200_LINSOLVE_GMRES)) return 1;
  CHECK(ctx, b200_ctx_destroy(ctx));
  int32_t ver = 0;
  if (b200_nccl_version(&ver) == B200_OK) {
    printf("NCCL %d\n", (int)ver);
    if (ensemble(ndev)) return 1;
  } else {
    printf("NCCL not loadable: collective step skipped\n");
    return 4;
  }
  printf("ABI_C_OK\n");
  return 0;
}
