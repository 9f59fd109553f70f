// newton.cu — the first-order nonlinear driver on the device (kernels K5/K6).
//
// Mirrors GeneralizedFirstOrderAlgorithm: __init (lib/NonlinearSolveFirstOrder/src/solve.jl:140-301), step! (:325-465),
// the driver loop (lib/NonlinearSolveBase/src/solve.jl:360-387, 835-858), NewtonDescent (descent/newton.jl:97-141),
// Dogleg (descent/dogleg.jl:86-151), GenericTrustRegionScheme / RadiusUpdateSchemes.Simple (trust_region.jl:204-258,
// 320-384, 396-430, 511-513), EisenstatWalkerForcing2 (eisenstat_walker.jl:42-87) and the default termination mode
// AbsNormSafeBestTerminationMode (termination_conditions.jl:134-179, 243-336, 376-393, 414-453).
//
// All n-vectors stay in HBM for the whole solve; per step the host reads back one small record of scalars
// (||f||_inf, ||du||_2^2) produced by the epilogues of the update / residual kernels, and takes the (scalar) control
// decisions of the reference on the host.
#include "common.cuh"
#include <math.h>
#include <algorithm>
#include <chrono>

extern "C" int32_t b200_gmres_keep_hessenberg(b200_gmres* gm, int64_t capacity);

namespace {
struct TermCache {  // NonlinearTerminationModeCache (termination_conditions.jl:60-98)
  int mode;
  int norm_kind;        // B200_NORM_*
  int max_stalled;      // window of the step-norm stall test (0: disabled)
  double abstol, reltol, best, initial, u0_norm;
  int nsteps;
  int retcode;
  double obj_trace[100];
  double step_trace[128];
};
// what the modes read of (du = f(u), u): filled on demand by term_quantities()
struct TermQuant {
  double f_inf, f_norm, sum_norm, rel_viol;  // ||f||_inf ; internalnorm(f) ; internalnorm(f + u) ; #{ |f_i| > reltol |u_i + f_i| }
};
}  // namespace

struct b200_newton {
  b200_ctx* ctx;
  b200_problem* prob;
  b200_newton_opts o;
  int64_t n;
  double abstol, reltol;
  int maxiters;
  double *u, *fu, *u_cache, *du, *xlin, *best_u;
  double *u_trial, *fu_trial, *Jdu, *JTfu, *du_c, *c1, *c2;
  b200_gmres* gm;
  b200_linop op;
  b200_linop prec;  // built-in preconditioner (opts.precond), re-pointed at the current iterate every step
  b200_mg* mg;      // multigrid hierarchy when opts.precond is MULTIGRID_*
  // dense
  double* Jdense;
  int64_t* ipiv;
  double* qr_work;    // pivoted-QR rescue of a singular LU (allocated on first use)
  int32_t* qr_jpvt;
  // sparse
  b200_sparse_jac* sj;
  double* nzval;
  b200_sparse_lu* slu;  // LINSOLVE_SPARSE_LU: band factorisation of the assembled Jacobian
  // LevenbergMarquardt: J'J + lambda D'D (factored in place), the running diagonal D'D, velocity / acceleration, previous velocity
  double *lmA, *lm_dtd, *lm_v, *lm_a, *lm_vold, *lm_rhs;
  int qn_since_du, qn_since_dfu, qn_nresets;  // Broyden: NoChangeInStateReset counters, resets so far
  double *lr_U, *lr_V, *lr_c;                // LimitedMemoryBroyden: J^-1 = lr_alpha I + U V' (n x lr_m each, circular), coefficient scratch
  int lr_m, lr_idx;
  double lr_alpha;
  double lm_lambda, lm_lambda_factor, lm_norm_v_old, lm_loss_old;
  // state
  TermCache tc;
  b200_newton_result res;
  std::vector<b200_trace_rec> trace;
  int retcode, force_stop, make_new_jacobian, nsteps, have_factor, initialised;
  double trust_region, max_tr;
  int shrink_counter;
  double eta, rnorm, rnorm_prev;
  double tp1, tp2, tp3, tp4;      // radius-update scheme parameters (get_parameters, trust_region.jl:372-379)
  double alpha_inv, pt_res_norm;  // PseudoTransient: 1/alpha and the residual 2-norm of the previous step (SER)
  double fnorm_inf;  // ||f(u)||_inf of the current iterate
  double total_time; // accumulated wall time of the steps (maxtime)
  double bytes;
};

namespace {
// ||x||_2 on the host (synchronising)
int32_t h_nrm2(b200_newton* nw, const double* x, double* out) { return b200_nrm2(nw->ctx, nw->n, x, out); }
int32_t h_dot(b200_newton* nw, const double* x, const double* y, double* out) { return b200_dot(nw->ctx, nw->n, x, y, out); }

bool term_is_safe(int mode) {
  return mode == B200_TERM_ABS_NORM_SAFE_BEST || mode == B200_TERM_ABS_NORM_SAFE || mode == B200_TERM_REL_NORM_SAFE || mode == B200_TERM_REL_NORM_SAFE_BEST;
}
bool term_is_best(int mode) { return mode == B200_TERM_ABS_NORM_SAFE_BEST || mode == B200_TERM_REL_NORM_SAFE_BEST; }
bool term_is_rel(int mode) { return mode == B200_TERM_REL_NORM_SAFE || mode == B200_TERM_REL_NORM_SAFE_BEST; }

// the reductions over (fu, u) a mode needs beyond ||fu||_inf (which the residual kernel's epilogue already produced)
int32_t term_quantities(b200_newton* nw, const double* fu, const double* u, double f_inf, TermQuant* q) {
  b200_ctx* ctx = nw->ctx;
  const int64_t n = nw->n;
  const int mode = nw->tc.mode;
  const bool l2 = nw->tc.norm_kind == B200_NORM_L2;
  q->f_inf = f_inf; q->f_norm = f_inf; q->sum_norm = 0.0; q->rel_viol = 0.0;
  const bool need_sum = mode == B200_TERM_NORM || mode == B200_TERM_REL_NORM || term_is_rel(mode);
  const bool need_viol = mode == B200_TERM_REL;
  const bool need_f2 = l2 && mode != B200_TERM_ABS && mode != B200_TERM_REL;
  if (!need_sum && !need_viol && !need_f2) return B200_OK;
  if (need_f2) B200_TRY(b200i_reduce_sum_dev(ctx, n, fu, nullptr, RED_SUMSQ, ctx->d_scalars + 8));
  if (need_sum) B200_TRY(b200i_reduce_sum_dev(ctx, n, fu, u, l2 ? RED_SUMSQ2 : RED_MAXABS2, ctx->d_scalars + 9));
  if (need_viol) B200_TRY(b200i_reduce_sum_dev(ctx, n, fu, u, RED_RELVIOL, ctx->d_scalars + 10, nw->tc.reltol));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalars + 8, ctx->d_scalars + 8, sizeof(double) * 3, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (need_f2) q->f_norm = sqrt(ctx->h_scalars[8]);
  if (need_sum) q->sum_norm = l2 ? sqrt(ctx->h_scalars[9]) : ctx->h_scalars[9];
  if (need_viol) q->rel_viol = ctx->h_scalars[10];
  return B200_OK;
}

// check_convergence (termination_conditions.jl:339-372) for the plain modes and the Safe-mode state machine (:243-336);
// `du_norm` = ||u - uprev||_2
int term_check(b200_newton* nw, const TermQuant& q, double du_norm, bool* new_best) {
  TermCache& tc = nw->tc;
  *new_best = false;
  switch (tc.mode) {
    case B200_TERM_ABS_NORM: if (q.f_norm <= tc.abstol) { tc.retcode = B200_RC_SUCCESS; return 1; } return 0;
    case B200_TERM_ABS: if (q.f_inf <= tc.abstol) { tc.retcode = B200_RC_SUCCESS; return 1; } return 0;          // all(|du| <= abstol)
    case B200_TERM_NORM: if (q.f_norm <= tc.abstol || q.f_norm <= tc.reltol * q.sum_norm) { tc.retcode = B200_RC_SUCCESS; return 1; } return 0;
    case B200_TERM_REL_NORM: if (q.f_norm <= tc.reltol * q.sum_norm) { tc.retcode = B200_RC_SUCCESS; return 1; } return 0;
    case B200_TERM_REL: if (q.rel_viol == 0.0) { tc.retcode = B200_RC_SUCCESS; return 1; } return 0;          // all(|du| <= reltol |u + du|)
    default: break;
  }
  const bool rel = term_is_rel(tc.mode);
  double objective, criteria;
  if (!rel) { objective = q.f_norm; criteria = tc.abstol; }
  else { objective = q.f_norm / (q.sum_norm + (nextafter(tc.reltol, INFINITY) - tc.reltol)); criteria = tc.reltol; }  // + eps(reltol)
  if (!std::isfinite(objective)) { tc.retcode = B200_RC_UNSTABLE; return 1; }
  if (term_is_best(tc.mode) && objective < tc.best) { tc.best = objective; *new_best = true; }
  if (objective <= criteria) { tc.retcode = B200_RC_SUCCESS; return 1; }
  tc.nsteps += 1;
  tc.obj_trace[(tc.nsteps - 1) % 100] = objective;
  if (objective <= 3.0 * criteria && tc.nsteps > 100) {  // patience_objective_multiplier = 3, patience_steps = 100, min_max_factor = 1.3
    double mn = INFINITY, mx = -INFINITY;
    for (int i = 0; i < 100; ++i) { mn = std::min(mn, tc.obj_trace[i]); mx = std::max(mx, tc.obj_trace[i]); }
    if (mn < 1.3 * mx) { tc.retcode = B200_RC_STALLED; return 1; }
  }
  if (tc.max_stalled > 0) {
    tc.step_trace[(tc.nsteps - 1) % tc.max_stalled] = du_norm;
    if (tc.nsteps > tc.max_stalled) {
      double mx = 0.0;
      for (int i = 0; i < tc.max_stalled; ++i) mx = std::max(mx, tc.step_trace[i]);
      const bool stalled = rel ? (mx <= tc.reltol * (mx + tc.u0_norm)) : (mx <= tc.abstol);
      if (stalled) { tc.retcode = B200_RC_STALLED; return 1; }
    }
  }
  tc.retcode = B200_RC_FAILURE;
  return 0;
}

int32_t alloc_vec(b200_ctx* ctx, int64_t n, double** p) {
  CUDA_TRY(ctx, cudaMalloc(p, sizeof(double) * ((n + 1) & ~(int64_t)1)));
  return B200_OK;
}
}  // namespace

extern "C" {
void b200_newton_opts_default(b200_newton_opts* o) {
  memset(o, 0, sizeof(*o));
  o->abstol = 0.0;  // => 3e-13
  o->reltol = 0.0;
  o->maxiters = 1000;
  o->linsolve = B200_LINSOLVE_GMRES;
  o->jvp_mode = B200_JVP_EXACT;
  o->globalization = B200_GLOBALIZATION_NONE;
  o->forcing = B200_FORCING_NONE;
  o->termination = B200_TERM_ABS_NORM_SAFE_BEST;
  o->store_trace = 1;
  o->fused_step = 1;
  b200_gmres_opts_default(&o->gmres);
  o->gmres.atol = 0.0;  // inherit the nonlinear tolerances (solve.jl:203)
  o->gmres.rtol = 0.0;
  o->ew_eta0 = 0.5; o->ew_eta_max = 0.9; o->ew_gamma = 0.9; o->ew_alpha = 2.0; o->ew_safeguard_threshold = 0.1; o->ew_safeguard = 1;
  o->max_shrink_times = 32;
}

int32_t b200_newton_destroy(b200_newton* nw) {
  B200_DEVICE_GUARD(nw ? nw->ctx : nullptr);
  if (!nw) return B200_OK;
  b200_ctx* ctx = nw->ctx;
  cudaStreamSynchronize(ctx->stream);
  double* vecs[] = {nw->u, nw->fu, nw->u_cache, nw->du, nw->xlin, nw->best_u, nw->u_trial, nw->fu_trial, nw->Jdu, nw->JTfu, nw->du_c, nw->c1, nw->c2,
                    nw->Jdense, nw->nzval, nw->lmA, nw->lm_dtd, nw->lm_v, nw->lm_a, nw->lm_vold, nw->lm_rhs, nw->lr_U, nw->lr_V, nw->lr_c};
  for (double* v : vecs) if (v) cudaFree(v);
  if (nw->ipiv) cudaFree(nw->ipiv);
  if (nw->qr_work) cudaFree(nw->qr_work);
  if (nw->qr_jpvt) cudaFree(nw->qr_jpvt);
  if (nw->gm) b200_gmres_destroy(nw->gm);
  if (nw->mg) b200i_mg_destroy(nw->mg);
  if (nw->sj) b200_sparse_jac_destroy(nw->sj);
  if (nw->slu) b200_sparse_lu_destroy(nw->slu);
  delete nw;
  return B200_OK;
}

int32_t b200_newton_create(b200_problem* prob, const b200_newton_opts* opts, b200_newton** out) {
  B200_DEVICE_GUARD(prob ? prob->ctx : nullptr);
  b200_ctx* ctx = prob->ctx;
  B200_REQUIRE(ctx, opts && out, "newton_create: bad arguments");
  B200_REQUIRE(ctx, opts->tr_scheme >= B200_TR_SIMPLE && opts->tr_scheme <= B200_TR_BASTIN, "newton_create: unknown radius update scheme");
  B200_REQUIRE(ctx, opts->termination >= B200_TERM_ABS_NORM_SAFE_BEST && opts->termination <= B200_TERM_REL_NORM_SAFE_BEST, "newton_create: unknown termination mode");
  B200_REQUIRE(ctx, opts->term_norm == B200_NORM_INF || opts->term_norm == B200_NORM_L2, "newton_create: unknown termination norm");
  B200_REQUIRE(ctx, opts->term_max_stalled_steps <= 128, "newton_create: term_max_stalled_steps must be <= 128");
  B200_REQUIRE(ctx, opts->descent == B200_DESCENT_NEWTON || (opts->descent == B200_DESCENT_PSEUDO_TRANSIENT && opts->globalization != B200_GLOBALIZATION_TRUST_REGION) ||
                        (opts->descent == B200_DESCENT_LEVENBERG_MARQUARDT && opts->globalization == B200_GLOBALIZATION_NONE && opts->linsolve == B200_LINSOLVE_DENSE_LU) ||
                        (opts->descent == B200_DESCENT_BROYDEN && opts->globalization == B200_GLOBALIZATION_NONE &&
                         (prob->n <= 65535 || opts->qn_init_jacobian == B200_QN_INIT_LOW_RANK || opts->qn_update_rule == B200_QN_UPDATE_KLEMENT) &&
                         (opts->qn_init_jacobian == B200_QN_INIT_IDENTITY || opts->qn_init_jacobian == B200_QN_INIT_LOW_RANK ||
                          (opts->qn_init_jacobian == B200_QN_INIT_TRUE_JACOBIAN && opts->linsolve == B200_LINSOLVE_DENSE_LU)) &&
                         (opts->qn_update_rule == B200_QN_UPDATE_GOOD_BROYDEN || opts->qn_update_rule == B200_QN_UPDATE_BAD_BROYDEN ||
                          (opts->qn_update_rule == B200_QN_UPDATE_KLEMENT && opts->qn_init_jacobian == B200_QN_INIT_IDENTITY))),
               "newton_create: descent must be Newton, PseudoTransient (without a trust region), LevenbergMarquardt (dense concrete Jacobian, its own trust region) or "
               "Broyden (no globalisation, n <= 65535, init_jacobian = true_jacobian needs the dense LU)");
  B200_REQUIRE(ctx, opts->precond == B200_PRECOND_NONE ||
                        ((opts->linsolve == B200_LINSOLVE_GMRES || opts->linsolve == B200_LINSOLVE_SPARSE_GMRES) &&
                         (prob->kind == B200_PROB_BRUSS2D || prob->kind == B200_PROB_BRUSS3D)),
               "newton_create: the built-in preconditioners need a Krylov linsolve and a built-in Brusselator problem");
  B200_REQUIRE(ctx, opts->precond >= B200_PRECOND_NONE && opts->precond <= B200_PRECOND_MULTIGRID_RIGHT, "newton_create: unknown preconditioner");
  b200_newton* nw = new b200_newton();
  nw->ctx = ctx; nw->prob = prob; nw->o = *opts; nw->n = prob->n;
  nw->abstol = opts->abstol > 0 ? opts->abstol : 3.0e-13;  // common_defaults.jl:44-48
  nw->reltol = opts->reltol > 0 ? opts->reltol : 3.0e-13;
  nw->maxiters = opts->maxiters > 0 ? opts->maxiters : 1000;
  nw->u = nw->fu = nw->u_cache = nw->du = nw->xlin = nw->best_u = nullptr;
  nw->u_trial = nw->fu_trial = nw->Jdu = nw->JTfu = nw->du_c = nw->c1 = nw->c2 = nullptr;
  nw->qr_work = nullptr; nw->qr_jpvt = nullptr;
  nw->lmA = nw->lm_dtd = nw->lm_v = nw->lm_a = nw->lm_vold = nw->lm_rhs = nullptr;
  nw->lr_U = nw->lr_V = nw->lr_c = nullptr; nw->lr_m = nw->lr_idx = 0; nw->lr_alpha = 1.0;
  nw->slu = nullptr; nw->mg = nullptr; nw->gm = nullptr; nw->Jdense = nullptr; nw->ipiv = nullptr; nw->sj = nullptr; nw->nzval = nullptr;
  nw->initialised = 0;
  const int64_t n = nw->n;
  int32_t s = B200_OK;
  auto A = [&](double** p) { if (s == B200_OK) s = alloc_vec(ctx, n, p); };
  A(&nw->u); A(&nw->fu); A(&nw->u_cache); A(&nw->du); A(&nw->xlin);
  if (term_is_best(opts->termination)) A(&nw->best_u);
  if (opts->globalization == B200_GLOBALIZATION_TRUST_REGION) {
    A(&nw->u_trial); A(&nw->fu_trial); A(&nw->Jdu); A(&nw->JTfu); A(&nw->du_c); A(&nw->c1); A(&nw->c2);
  }
  if (opts->globalization == B200_GLOBALIZATION_LINESEARCH) { A(&nw->u_trial); A(&nw->fu_trial); A(&nw->Jdu); }
  if (opts->descent == B200_DESCENT_LEVENBERG_MARQUARDT) {
    A(&nw->u_trial); A(&nw->fu_trial); A(&nw->Jdu); A(&nw->lm_dtd); A(&nw->lm_v); A(&nw->lm_a); A(&nw->lm_vold); A(&nw->lm_rhs);
    if (s == B200_OK && cudaMalloc(&nw->lmA, sizeof(double) * n * n) != cudaSuccess) { cudaGetLastError(); s = ctx->fail(B200_ERR_NOMEM, "LevenbergMarquardt: J'J does not fit in device memory", __FILE__, __LINE__); }
  }
  if (opts->descent == B200_DESCENT_BROYDEN) {  // the stored inverse and the update rule's / reset condition's vectors share the LM slots
    A(&nw->lm_dtd); A(&nw->lm_v); A(&nw->lm_a); A(&nw->lm_vold); A(&nw->lm_rhs);
    if (opts->qn_update_rule == B200_QN_UPDATE_KLEMENT) {
      // diagonal structure: the approximate Jacobian is the n-vector lm_dtd, the rule's residual cache lm_v — nothing n x n
    } else if (opts->qn_init_jacobian == B200_QN_INIT_LOW_RANK) {
      nw->lr_m = std::max(1, std::min(opts->qn_threshold > 0 ? opts->qn_threshold : 10, nw->maxiters));  // threshold = min(threshold, maxiters)
      if (s == B200_OK && (cudaMalloc(&nw->lr_U, sizeof(double) * n * nw->lr_m) != cudaSuccess || cudaMalloc(&nw->lr_V, sizeof(double) * n * nw->lr_m) != cudaSuccess ||
                           cudaMalloc(&nw->lr_c, sizeof(double) * nw->lr_m) != cudaSuccess)) {
        cudaGetLastError();
        s = ctx->fail(B200_ERR_NOMEM, "LimitedMemoryBroyden: the low-rank factors do not fit in device memory", __FILE__, __LINE__);
      }
    } else if (s == B200_OK && cudaMalloc(&nw->lmA, sizeof(double) * n * n) != cudaSuccess) { cudaGetLastError(); s = ctx->fail(B200_ERR_NOMEM, "Broyden: the stored inverse Jacobian (n x n) does not fit in device memory", __FILE__, __LINE__); }
  }
  if (s != B200_OK) { b200_newton_destroy(nw); return s; }
  memset(&nw->op, 0, sizeof(nw->op));
  nw->op.ctx = ctx; nw->op.n = n;
  if (opts->precond == B200_PRECOND_MULTIGRID_LEFT || opts->precond == B200_PRECOND_MULTIGRID_RIGHT) {
    s = b200i_mg_create(prob, &nw->mg);
    if (s != B200_OK) { b200_newton_destroy(nw); return s; }
  }
  if (opts->linsolve == B200_LINSOLVE_GMRES || opts->linsolve == B200_LINSOLVE_SPARSE_GMRES) {
    b200_gmres_opts g = opts->gmres;
    if (g.atol <= 0) g.atol = nw->abstol;  // linsolve_kwargs = (; abstol, reltol)   solve.jl:203
    if (g.rtol <= 0) g.rtol = nw->reltol;
    nw->o.gmres = g;
    s = b200_gmres_create(ctx, n, &g, &nw->gm);
    if (s != B200_OK) { b200_newton_destroy(nw); return s; }
  }
  if (opts->linsolve == B200_LINSOLVE_GMRES) {
    nw->op.kind = LINOP_PROBLEM; nw->op.prob = prob; nw->op.u = nw->u; nw->op.jvp_mode = opts->jvp_mode;
  } else if (opts->linsolve == B200_LINSOLVE_DENSE_LU) {
    if (cudaMalloc(&nw->Jdense, sizeof(double) * n * n) != cudaSuccess || cudaMalloc(&nw->ipiv, sizeof(int64_t) * n) != cudaSuccess) {
      cudaGetLastError();
      b200_newton_destroy(nw);
      return ctx->fail(B200_ERR_NOMEM, "dense Jacobian does not fit in device memory", __FILE__, __LINE__);
    }
  } else if (opts->linsolve == B200_LINSOLVE_SPARSE_GMRES || opts->linsolve == B200_LINSOLVE_SPARSE_LU) {
    // jac_prototype + colouring once at init (jacobian.jl:286-353; colouring ext :13-28)
    int64_t nnz = 0;
    s = b200_pattern_nnz(prob, &nnz);
    std::vector<int64_t> colptr(n + 1), rowval(nnz), colors(n);
    int64_t ncolors = 0;
    if (s == B200_OK) s = b200_pattern(prob, 1, colptr.data(), rowval.data());
    if (s == B200_OK) s = b200_coloring_column(n, colptr.data(), rowval.data(), 1, B200_ORDER_LARGEST_FIRST, colors.data(), &ncolors);
    if (s == B200_OK) s = b200_sparse_jac_create(prob, colptr.data(), rowval.data(), 1, colors.data(), ncolors, &nw->sj);
    if (s == B200_OK && cudaMalloc(&nw->nzval, sizeof(double) * nnz) != cudaSuccess) { cudaGetLastError(); s = B200_ERR_NOMEM; }
    // sparse direct route (linsolve = nothing on a sparse prototype): symbolic phase once, like LinearSolve's cache
    if (s == B200_OK && opts->linsolve == B200_LINSOLVE_SPARSE_LU) s = b200_sparse_lu_create(ctx, n, colptr.data(), rowval.data(), 1, &nw->slu);
    if (s != B200_OK) { b200_newton_destroy(nw); return s; }
    nw->op.kind = LINOP_SPARSE_JAC; nw->op.sj = nw->sj; nw->op.nzval = nw->nzval;
  } else {
    b200_newton_destroy(nw);
    return ctx->fail(B200_ERR_INVALID, "unknown linsolve kind", __FILE__, __LINE__);
  }
  *out = nw;
  return B200_OK;
}

// reinit!(cache, u0): everything __init computes from u0 (solve.jl:191-284; termination_conditions.jl:134-179)
int32_t b200_newton_reinit(b200_newton* nw, const double* u0_dev) {
  B200_DEVICE_GUARD(nw ? nw->ctx : nullptr);
  b200_ctx* ctx = nw->ctx;
  const int64_t n = nw->n;
  const b200_newton_opts& o = nw->o;
  if (u0_dev != nw->u) CUDA_TRY(ctx, cudaMemcpyAsync(nw->u, u0_dev, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_scalars, 0, sizeof(double) * 4, ctx->stream));
  B200_TRY(b200i_residual_norm(nw->prob, nw->u, nw->fu, ctx->d_scalars));  // evaluate_f(prob,u): nf is NOT bumped (solve.jl:194)
  CUDA_TRY(ctx, cudaMemcpyAsync(nw->u_cache, nw->u, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemsetAsync(nw->du, 0, sizeof(double) * n, ctx->stream));  // descent caches start with a defined du
  if (nw->best_u) CUDA_TRY(ctx, cudaMemcpyAsync(nw->best_u, nw->u, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  B200_TRY(b200i_fetch_scalars(ctx, 1));
  nw->fnorm_inf = ctx->h_scalars[0];
  memset(&nw->tc, 0, sizeof(nw->tc));
  nw->tc.mode = o.termination;
  nw->tc.norm_kind = o.term_norm;
  nw->tc.max_stalled = !term_is_safe(o.termination) ? 0 : (o.term_max_stalled_steps < 0 ? 0 : (o.term_max_stalled_steps == 0 ? 32 : o.term_max_stalled_steps));
  nw->tc.abstol = nw->abstol;
  nw->tc.reltol = nw->reltol;
  nw->total_time = 0.0;
  {  // SciMLBase.reinit!(::NonlinearTerminationModeCache, du, u)   termination_conditions.jl:180-215
    TermQuant q;
    B200_TRY(term_quantities(nw, nw->fu, nw->u, nw->fnorm_inf, &q));
    if (!term_is_safe(o.termination)) nw->tc.best = INFINITY;
    else if (!term_is_rel(o.termination)) nw->tc.best = q.f_norm;
    else {
      nw->tc.best = q.f_norm / (q.sum_norm + 2.220446049250313e-16);   // eps(TT)
      if (nw->tc.max_stalled > 0) B200_TRY(h_nrm2(nw, nw->u, &nw->tc.u0_norm));
    }
    nw->tc.initial = nw->tc.best;
  }
  memset(&nw->res, 0, sizeof(nw->res));
  nw->trace.clear();
  nw->retcode = B200_RC_DEFAULT; nw->force_stop = 0; nw->make_new_jacobian = 1; nw->nsteps = 0; nw->have_factor = 0;
  nw->shrink_counter = 0; nw->bytes = 2.0 * 8.0 * n;
  if (o.linsolve == B200_LINSOLVE_DENSE_LU) nw->res.njacs += 1;  // jac_prototype === nothing: DI.jacobian at init (jacobian.jl:103-117)
  if (o.globalization == B200_GLOBALIZATION_TRUST_REGION) {  // trust_region.jl:204-258, 330-346 (Simple scheme)
    double fu_norm, u0_norm, umin, umax;
    const int sch = o.tr_scheme;
    B200_TRY(h_nrm2(nw, nw->fu, &fu_norm));
    B200_TRY(h_nrm2(nw, nw->u, &u0_norm));
    B200_TRY(b200_extrema(ctx, n, nw->u, &umin, &umax));
    nw->tp1 = nw->tp2 = nw->tp3 = nw->tp4 = 0.0;
    if (sch == B200_TR_NLSOLVE) { nw->tp1 = 0.5; }
    else if (sch == B200_TR_HEI) { nw->tp1 = 5.0; nw->tp2 = 0.1; nw->tp3 = 0.15; nw->tp4 = 0.15; }
    else if (sch == B200_TR_YUAN) { nw->tp1 = 2.0; nw->tp2 = 1.0 / 6; nw->tp3 = 6.0; }
    else if (sch == B200_TR_FAN) { nw->tp1 = 0.1; nw->tp2 = 0.25; nw->tp3 = 12.0; nw->tp4 = 1.0e18; }
    else if (sch == B200_TR_BASTIN) { nw->tp1 = 2.5; nw->tp2 = 0.25; }
    if (o.tr_max_trust_radius > 0) nw->max_tr = o.tr_max_trust_radius;  // max_trust_radius  :330-337
    else nw->max_tr = (sch == B200_TR_SIMPLE || sch == B200_TR_NOCEDAL_WRIGHT) ? std::max(fu_norm, umax - umin) : INFINITY;  // Hei, Yuan, Fan, NLsolve, Bastin: unbounded
    if (o.tr_initial_trust_radius > 0) nw->trust_region = o.tr_initial_trust_radius;  // initial_trust_radius  :339-346
    else if (sch == B200_TR_NLSOLVE) nw->trust_region = u0_norm > 0 ? u0_norm : 1.0;
    else if (sch == B200_TR_HEI || sch == B200_TR_BASTIN) nw->trust_region = 1.0;
    else if (sch == B200_TR_FAN) nw->trust_region = pow(fu_norm, 0.99) / 10.0;
    else nw->trust_region = nw->max_tr / 11.0;
    if (sch == B200_TR_YUAN) {  // itr = p1 ||J' fu||  :232-234
      double g;
      B200_TRY(b200_vjp(nw->prob, nw->u, nw->fu, nw->JTfu));
      B200_TRY(h_nrm2(nw, nw->JTfu, &g));
      nw->trust_region = nw->tp1 * g;
    }
  }
  if (o.descent == B200_DESCENT_PSEUDO_TRANSIENT) {  // SwitchedEvolutionRelaxationCache init / reinit! (pseudo_transient.jl:105-130)
    nw->alpha_inv = 1.0 / (o.pt_alpha_initial > 0 ? o.pt_alpha_initial : 1.0e-3);
    B200_TRY(h_nrm2(nw, nw->fu, &nw->pt_res_norm));
  }
  if (o.descent == B200_DESCENT_LEVENBERG_MARQUARDT) {  // LevenbergMarquardtDampingCache / TrustRegionCache init + reinit!  levenberg_marquardt.jl:71-117, 226-262
    nw->lm_lambda = o.lm_damping_initial > 0 ? o.lm_damping_initial : 1.0;
    nw->lm_lambda_factor = o.lm_damping_increase > 0 ? o.lm_damping_increase : 2.0;
    B200_TRY(b200_fill(ctx, n, o.lm_min_damping_D > 0 ? o.lm_min_damping_D : 1.0e-8, nw->lm_dtd));
    B200_TRY(b200_copy(ctx, n, nw->u, nw->lm_vold));   // `@bb v = copy(u)`: the previous velocity starts as u0
    B200_TRY(b200_fill(ctx, n, 0.0, nw->du));
    nw->lm_norm_v_old = INFINITY;
    nw->lm_loss_old = INFINITY;
  }
  if (o.descent == B200_DESCENT_BROYDEN) {  // BroydenUpdateRuleCache.dfu = copy(fu); NoChangeInStateResetCache.dfu = copy(fu); counters 0
    B200_TRY(b200_copy(ctx, n, nw->fu, nw->lm_v));
    B200_TRY(b200_copy(ctx, n, nw->fu, nw->lm_a));
    B200_TRY(b200_fill(ctx, n, 0.0, nw->du));
    nw->qn_since_du = nw->qn_since_dfu = nw->qn_nresets = 0;
  }
  nw->op.shift = 0.0;
  nw->eta = o.ew_eta0;
  if (o.forcing == B200_FORCING_EW2) {
    B200_TRY(h_nrm2(nw, nw->fu, &nw->rnorm));
    nw->rnorm_prev = nw->rnorm;
  }
  nw->initialised = 1;
  return B200_OK;
}

// One step of LevenbergMarquardt() (levenberg_marquardt.jl; descent/damped_newton.jl :normal_form; descent/geodesic_acceleration.jl:
// 95-135; step! solve.jl:325-465).  All n-vectors and both n x n matrices stay on the device; the host takes the scalar
// decisions from a handful of norms / dots.
static int32_t lm_step(b200_newton* nw) {
  b200_ctx* ctx = nw->ctx;
  const int64_t n = nw->n;
  const b200_newton_opts& o = nw->o;
  const double inc = o.lm_damping_increase > 0 ? o.lm_damping_increase : 2.0, dec = o.lm_damping_decrease > 0 ? o.lm_damping_decrease : 3.0;
  const double h = o.lm_finite_diff_step > 0 ? o.lm_finite_diff_step : 0.1, alpha_geo = o.lm_alpha_geodesic > 0 ? o.lm_alpha_geodesic : 0.75;
  const double b_uphill = o.lm_b_uphill > 0 ? o.lm_b_uphill : (o.lm_b_uphill == 0 ? 1.0 : 0.0);
  if (nw->make_new_jacobian) {  // J = cache.jac_cache(u)
    nw->res.njacs += 1;
    B200_TRY(b200_dense_jac_fill(nw->prob, nw->u, nw->Jdense, n));
  }
  // ---- DampedNewtonDescent, normal form: A = J'J + lambda D'D with D'D = running max of diag(J'J); v = -(A^-1 J' f)
  B200_TRY(b200i_gram(ctx, n, nw->Jdense, n, nw->lmA, n));
  B200_TRY(b200i_lm_damp(ctx, n, nw->lmA, n, nw->lm_dtd, nw->lm_lambda));
  int32_t info = 0;
  nw->res.nfactors += 1;
  B200_TRY(b200_getrf(ctx, n, nw->lmA, n, nw->ipiv, &info));
  int descent_ok = 1, tr_ok = 0, accepted = 0;
  double objective = nw->fnorm_inf, du_norm = 0.0;
  if (info != 0) {  // the damped normal matrix is singular: LinearSolve failure with a current Jacobian
    if (nw->make_new_jacobian) { nw->retcode = B200_RC_INTERNAL_LINSOLVE_FAILED; nw->force_stop = 1; return B200_OK; }
    nw->make_new_jacobian = 1;
    return lm_step(nw);
  }
  B200_TRY(b200_gemv(ctx, 1, n, n, nw->Jdense, n, nw->fu, nw->lm_rhs));          // J' f
  B200_TRY(b200_copy(ctx, n, nw->lm_rhs, nw->lm_v));
  nw->res.nsolve += 1;
  B200_TRY(b200_getrs(ctx, n, 1, nw->lmA, n, nw->ipiv, nw->lm_v, n));
  B200_TRY(b200_scal(ctx, n, -1.0, nw->lm_v));
  double norm_v;
  B200_TRY(h_nrm2(nw, nw->lm_v, &norm_v));
  if (!o.lm_disable_geodesic) {
    // geodesic acceleration: fu_cache = (2/h) ((f(u + h v) - f(u)) / h - J v) ; a = -(A^-1 J' fu_cache) with the same factorisation
    B200_TRY(b200_copy(ctx, n, nw->u, nw->u_trial));
    B200_TRY(b200_axpy(ctx, n, h, nw->lm_v, nw->u_trial));
    B200_TRY(b200_residual(nw->prob, nw->u_trial, nw->fu_trial));              // evaluate_f!! inside the descent: NLStats.nf is not bumped
    B200_TRY(b200_gemv(ctx, 0, n, n, nw->Jdense, n, nw->lm_v, nw->Jdu));        // J v
    B200_TRY(b200_axpy(ctx, n, -1.0, nw->fu, nw->fu_trial));
    B200_TRY(b200_axpby(ctx, n, -2.0 / h, nw->Jdu, 2.0 / (h * h), nw->fu_trial));
    B200_TRY(b200_gemv(ctx, 1, n, n, nw->Jdense, n, nw->fu_trial, nw->lm_a));
    nw->res.nsolve += 1;
    B200_TRY(b200_getrs(ctx, n, 1, nw->lmA, n, nw->ipiv, nw->lm_a, n));
    B200_TRY(b200_scal(ctx, n, -1.0, nw->lm_a));
    double norm_a;
    B200_TRY(h_nrm2(nw, nw->lm_a, &norm_a));
    if (2.0 * norm_a <= norm_v * alpha_geo) {
      B200_TRY(b200_copy(ctx, n, nw->lm_v, nw->du));
      B200_TRY(b200_axpy(ctx, n, 0.5, nw->lm_a, nw->du));                       // du = v + a / 2
    } else {
      descent_ok = 0;  // the step is not taken; du keeps its previous value (geodesic_acceleration.jl:131-133)
    }
  } else {
    nw->res.nsolve += 0;
    B200_TRY(b200_copy(ctx, n, nw->lm_v, nw->du));
  }
  if (descent_ok) {
    nw->make_new_jacobian = 1;
    // ---- LevenbergMarquardtTrustRegion: beta = cos(v, v_old); accept iff (1 - beta)^b_uphill * ||f(u + du)|| <= loss_old
    double vdot;
    B200_TRY(h_dot(nw, nw->lm_v, nw->lm_vold, &vdot));
    const double beta = vdot / (norm_v * nw->lm_norm_v_old);
    B200_TRY(b200_copy(ctx, n, nw->u, nw->u_trial));
    B200_TRY(b200_axpy(ctx, n, 1.0, nw->du, nw->u_trial));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_scalars, 0, sizeof(double) * 2, ctx->stream));
    B200_TRY(b200i_residual_norm(nw->prob, nw->u_trial, nw->fu_trial, ctx->d_scalars));
    nw->res.nf += 1;
    B200_TRY(b200i_fetch_scalars(ctx, 1));
    const double trial_inf = ctx->h_scalars[0];
    double loss;
    B200_TRY(h_nrm2(nw, nw->fu_trial, &loss));
    tr_ok = (pow(1.0 - beta, b_uphill) * loss <= nw->lm_loss_old) ? 1 : 0;     // loss_old is never updated by the reference (stays Inf)
    if (tr_ok) {
      nw->lm_norm_v_old = norm_v;
      B200_TRY(b200_copy(ctx, n, nw->lm_v, nw->lm_vold));
      B200_TRY(h_nrm2(nw, nw->du, &du_norm));
      std::swap(nw->u, nw->u_trial);
      std::swap(nw->fu, nw->fu_trial);
      objective = trial_inf;
      accepted = 1;
    } else {
      nw->make_new_jacobian = 0;
    }
    nw->fnorm_inf = objective;
    bool new_best = false;
    TermQuant tq;
    B200_TRY(term_quantities(nw, nw->fu, nw->u, objective, &tq));
    if (term_check(nw, tq, du_norm, &new_best)) { nw->retcode = nw->tc.retcode; nw->force_stop = 1; }
    if (new_best && nw->best_u) CUDA_TRY(ctx, cudaMemcpyAsync(nw->best_u, nw->u, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    nw->make_new_jacobian = 0;
  }
  if (o.store_trace) {
    b200_trace_rec t;
    memset(&t, 0, sizeof(t));
    t.iter = nw->nsteps + 1; t.accepted = accepted; t.fnorm_inf = objective; t.step_norm2 = du_norm; t.trust_radius = nw->lm_lambda;  // the damping USED by this step
    t.lin_status = descent_ok;
    nw->trace.push_back(t);
  }
  // callback_into_cache! (levenberg_marquardt.jl:176-185): lambda shrinks after a step both the descent and the trust region accepted
  if (descent_ok && tr_ok) nw->lm_lambda_factor = 1.0 / dec;
  nw->lm_lambda *= nw->lm_lambda_factor;
  nw->lm_lambda_factor = inc;
  return B200_OK;
}

// One step of Broyden() (NonlinearSolveQuasiNewton/src/solve.jl:293-486, broyden.jl:124-144, reset_conditions.jl:52-88,
// initialization.jl:78-105).  The stored inverse J^-1 (n x n, column-major) lives in HBM; a step is GEMV-shaped:
//   du = -(J^-1 f)                               1 pass over J^-1
//   J^-1 += ((du - J^-1 df) / denom) w'          good Broyden: w = J^-T du, denom = <du, J^-1 df>   -> 2 GEMV + 1 GER = 4 passes
//                                                bad Broyden : w = df,      denom = ||df||^2        -> 1 GEMV + 1 GER = 3 passes
// The host takes the scalar decisions (reset counters, max_resets, termination) from a handful of reductions.
static int32_t broyden_init_inverse(b200_newton* nw) {
  b200_ctx* ctx = nw->ctx;
  const int64_t n = nw->n;
  double* Jinv = nw->lmA;
  if (nw->o.qn_init_jacobian == B200_QN_INIT_TRUE_JACOBIAN) {  // J^-1 = J \ I through the LU (Utils.linsolve_identity!!)
    nw->res.njacs += 1;
    B200_TRY(b200_dense_jac_fill(nw->prob, nw->u, nw->Jdense, n));
    int32_t info = 0;
    nw->res.nfactors += 1;
    B200_TRY(b200_getrf(ctx, n, nw->Jdense, n, nw->ipiv, &info));
    if (info != 0) { nw->retcode = B200_RC_INTERNAL_LINSOLVE_FAILED; nw->force_stop = 1; return B200_OK; }
    B200_TRY(b200i_scaled_identity(ctx, n, Jinv, n, 1.0));
    B200_TRY(b200_getrs(ctx, n, n, nw->Jdense, n, nw->ipiv, Jinv, n));
    return B200_OK;
  }
  double alpha = nw->o.qn_alpha;
  if (!(alpha > 0)) {  // Utils.initial_jacobian_scaling_alpha(nothing, u, fu, L2): 2 ||f|| / max(||u||, 1); 1 below ||f|| = 1e-5
    double fn, un;
    B200_TRY(h_nrm2(nw, nw->fu, &fn));
    B200_TRY(h_nrm2(nw, nw->u, &un));
    alpha = (fn < 1.0e-5) ? 1.0 : (2.0 * fn) / std::max(un, 1.0);
  }
  if (nw->o.qn_init_jacobian == B200_QN_INIT_LOW_RANK) {  // BroydenLowRankJacobian: idx = 0, alpha = inv(scaling)   initialization.jl:176-199
    nw->lr_idx = 0;
    nw->lr_alpha = 1.0 / alpha;
    return B200_OK;
  }
  return b200i_scaled_identity(ctx, n, Jinv, n, 1.0 / alpha);
}

// y = J^-1 x (transpose = 0) or J^-T x (1): the dense stored inverse, or alpha x + U (V' x) resp. alpha x + V (U' x) for the low-rank form
static int32_t broyden_apply(b200_newton* nw, int transpose, const double* x, double* y) {
  b200_ctx* ctx = nw->ctx;
  const int64_t n = nw->n;
  if (nw->o.qn_init_jacobian != B200_QN_INIT_LOW_RANK) return b200_gemv(ctx, transpose, n, n, nw->lmA, n, x, y);
  const int k = std::min(nw->lr_idx, nw->lr_m);
  if (k > 0) {
    const double *L = transpose ? nw->lr_V : nw->lr_U, *R = transpose ? nw->lr_U : nw->lr_V;
    B200_TRY(b200_gemv(ctx, 1, n, k, R, n, x, nw->lr_c));   // c = R' x   (k dot products)
    B200_TRY(b200_gemv(ctx, 0, n, k, L, n, nw->lr_c, y));   // y = L c
    return b200_axpy(ctx, n, nw->lr_alpha, x, y);
  }
  return b200_axpby(ctx, n, nw->lr_alpha, x, 0.0, y);
}

static int32_t broyden_count(b200_newton* nw, const double* x, const double* y, double tol, double* out) {
  b200_ctx* ctx = nw->ctx;
  B200_TRY(b200i_reduce_sum_dev(ctx, nw->n, x, y, y ? RED_COUNT_DIFF_LE : RED_COUNT_LE, ctx->d_scalars, tol));
  B200_TRY(b200i_fetch_scalars(ctx, 1));
  *out = ctx->h_scalars[0];
  return B200_OK;
}

// One step of Klement() with its default diagonal structure (klement.jl:30-49, 128-140; reset_conditions.jl:103-120; solve.jl:293-486):
// J is an n-vector, so descent, update and the reset test are elementwise kernels and one counting reduction — any n.
static int32_t klement_step(b200_newton* nw) {
  b200_ctx* ctx = nw->ctx;
  const int64_t n = nw->n;
  const b200_newton_opts& o = nw->o;
  double *J = nw->lm_dtd, *fu_cache = nw->lm_v;
  const int max_resets = o.qn_max_resets > 0 ? o.qn_max_resets : 100;
  auto init_diag = [&]() -> int32_t {
    double alpha = o.qn_alpha;
    if (!(alpha > 0)) {
      double fn, un;
      B200_TRY(h_nrm2(nw, nw->fu, &fn));
      B200_TRY(h_nrm2(nw, nw->u, &un));
      alpha = (fn < 1.0e-5) ? 1.0 : (2.0 * fn) / std::max(un, 1.0);
    }
    return b200_fill(ctx, n, alpha, J);   // J = one.(fu) .* alpha, NOT inverted (store_inverse_jacobian = false)
  };
  int reset = 0;
  if (nw->nsteps == 0) {
    B200_TRY(init_diag());
  } else {
    double zeros;  // IllConditionedJacobianReset on a Diagonal: any(iszero, diag(J))
    B200_TRY(broyden_count(nw, J, nullptr, 0.0, &zeros));
    if (zeros > 0) {
      reset = 1;
      if (++nw->qn_nresets >= max_resets) { nw->retcode = B200_RC_CONVERGENCE_FAILURE; nw->force_stop = 1; return B200_OK; }
      B200_TRY(init_diag());
    }
  }
  B200_TRY(b200i_klement_descent(ctx, n, J, nw->fu, nw->du));
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_scalars, 0, sizeof(double) * 2, ctx->stream));
  B200_TRY(b200i_axpy_norm(ctx, n, 1.0, nw->du, nw->u, ctx->d_scalars + 1));
  B200_TRY(b200i_residual_norm(nw->prob, nw->u, nw->fu, ctx->d_scalars));
  nw->res.nf += 1;
  nw->res.nsolve += 1;   // the diagonal system goes through NativeJLLinearSolveCache (linear_solve.jl:130-134): nsolve and nfactors both +1
  nw->res.nfactors += 1;
  B200_TRY(b200i_fetch_scalars(ctx, 2));
  const double objective = ctx->h_scalars[0], du_norm = sqrt(ctx->h_scalars[1]);
  nw->fnorm_inf = objective;
  nw->bytes += 8.0 * (double)n * 8.0;
  bool new_best = false;
  TermQuant tq;
  B200_TRY(term_quantities(nw, nw->fu, nw->u, objective, &tq));
  if (term_check(nw, tq, du_norm, &new_best)) { nw->retcode = nw->tc.retcode; nw->force_stop = 1; }
  if (new_best && nw->best_u) CUDA_TRY(ctx, cudaMemcpyAsync(nw->best_u, nw->u, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  if (o.store_trace) {
    b200_trace_rec t;
    memset(&t, 0, sizeof(t));
    t.iter = nw->nsteps + 1; t.accepted = 1; t.fnorm_inf = objective; t.step_norm2 = du_norm; t.lin_status = reset;
    nw->trace.push_back(t);
  }
  if (nw->force_stop) return B200_OK;
  return b200i_klement_update(ctx, n, J, nw->fu, fu_cache, nw->du);
}

static int32_t broyden_step(b200_newton* nw) {
  if (nw->o.qn_update_rule == B200_QN_UPDATE_KLEMENT) return klement_step(nw);
  b200_ctx* ctx = nw->ctx;
  const int64_t n = nw->n;
  const b200_newton_opts& o = nw->o;
  double *Jinv = nw->lmA, *dfu_rule = nw->lm_v, *dfu_reset = nw->lm_a, *Jd = nw->lm_vold, *w = nw->lm_rhs, *c = nw->lm_dtd;
  const double tol = o.qn_reset_tolerance > 0 ? o.qn_reset_tolerance : 1.8189894035458565e-12;  // eps^(3/4)
  const int max_resets = o.qn_max_resets > 0 ? o.qn_max_resets : 100;
  int reset = 0;
  if (nw->nsteps == 0) {
    B200_TRY(broyden_init_inverse(nw));
    if (nw->force_stop) return B200_OK;
  } else {
    // NoChangeInStateReset(nsteps = 3): ANY component of du (then of f - f_prev) at or below the tolerance counts as "no change"
    double cnt;
    B200_TRY(broyden_count(nw, nw->du, nullptr, tol, &cnt));
    if (cnt > 0) {
      if (++nw->qn_since_du >= 3) { nw->qn_since_du = nw->qn_since_dfu = 0; reset = 1; }
    } else {
      nw->qn_since_du = nw->qn_since_dfu = 0;
    }
    if (!reset) {
      B200_TRY(broyden_count(nw, nw->fu, dfu_reset, tol, &cnt));
      if (cnt > 0) {
        if (++nw->qn_since_dfu >= 3) { nw->qn_since_dfu = nw->qn_since_du = 0; reset = 1; }
      } else {
        nw->qn_since_dfu = nw->qn_since_du = 0;
      }
      B200_TRY(b200_copy(ctx, n, nw->fu, dfu_reset));
    }
    if (reset) {
      if (++nw->qn_nresets >= max_resets) { nw->retcode = B200_RC_CONVERGENCE_FAILURE; nw->force_stop = 1; return B200_OK; }
      B200_TRY(broyden_init_inverse(nw));
      if (nw->force_stop) return B200_OK;
    }
  }
  // ---- NewtonDescent on the stored inverse: du = -(J^-1 f) ; u += du ; f = f(u)
  B200_TRY(broyden_apply(nw, 0, nw->fu, nw->du));
  B200_TRY(b200_scal(ctx, n, -1.0, nw->du));
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_scalars, 0, sizeof(double) * 2, ctx->stream));
  B200_TRY(b200i_axpy_norm(ctx, n, 1.0, nw->du, nw->u, ctx->d_scalars + 1));
  B200_TRY(b200i_residual_norm(nw->prob, nw->u, nw->fu, ctx->d_scalars));
  nw->res.nf += 1;
  B200_TRY(b200i_fetch_scalars(ctx, 2));
  const double objective = ctx->h_scalars[0], du_norm = sqrt(ctx->h_scalars[1]);
  nw->fnorm_inf = objective;
  const double pass_bytes = (o.qn_init_jacobian == B200_QN_INIT_LOW_RANK) ? 16.0 * (double)n * std::min(nw->lr_idx, nw->lr_m) : 8.0 * (double)n * (double)n;  // one product with the stored inverse
  nw->bytes += pass_bytes;
  bool new_best = false;
  TermQuant tq;
  B200_TRY(term_quantities(nw, nw->fu, nw->u, objective, &tq));
  if (term_check(nw, tq, du_norm, &new_best)) { nw->retcode = nw->tc.retcode; nw->force_stop = 1; }
  if (new_best && nw->best_u) CUDA_TRY(ctx, cudaMemcpyAsync(nw->best_u, nw->u, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
  if (o.store_trace) {
    b200_trace_rec t;
    memset(&t, 0, sizeof(t));
    t.iter = nw->nsteps + 1; t.accepted = 1; t.fnorm_inf = objective; t.step_norm2 = du_norm; t.lin_status = reset;  // lin_status: J^-1 was re-initialised before this step
    nw->trace.push_back(t);
  }
  if (nw->force_stop) return B200_OK;  // the reference skips the update once the step has stopped the solve
  // ---- update rule
  B200_TRY(b200_axpby(ctx, n, 1.0, nw->fu, -1.0, dfu_rule));             // dfu = fu - dfu
  B200_TRY(broyden_apply(nw, 0, dfu_rule, Jd));                          // J^-1 dfu
  double denom;
  const double* rmul;
  if (o.qn_update_rule == B200_QN_UPDATE_GOOD_BROYDEN) {
    B200_TRY(broyden_apply(nw, 1, nw->du, w));                           // J^-T du
    B200_TRY(h_dot(nw, nw->du, Jd, &denom));
    rmul = w;
  } else {
    double nd;
    B200_TRY(h_nrm2(nw, dfu_rule, &nd));
    denom = nd * nd;
    rmul = dfu_rule;
  }
  const double inv = 1.0 / (denom == 0.0 ? 1.0e-5 : denom);
  B200_TRY(b200_copy(ctx, n, nw->du, c));
  B200_TRY(b200_axpy(ctx, n, -1.0, Jd, c));
  B200_TRY(b200_scal(ctx, n, inv, c));                                   // (du - J^-1 dfu) / denom
  if (o.qn_init_jacobian == B200_QN_INIT_LOW_RANK) {                      // mul!(J, u, v', true, true): the pair goes into slot idx mod m
    const int slot = nw->lr_idx % nw->lr_m;
    B200_TRY(b200_copy(ctx, n, c, nw->lr_U + (int64_t)slot * n));
    B200_TRY(b200_copy(ctx, n, rmul, nw->lr_V + (int64_t)slot * n));
    nw->lr_idx += 1;
  } else {
    B200_TRY(b200i_ger(ctx, n, Jinv, n, c, rmul));                       // J^-1 += c rmul'
  }
  B200_TRY(b200_copy(ctx, n, nw->fu, dfu_rule));
  nw->bytes += pass_bytes * (o.qn_update_rule == B200_QN_UPDATE_GOOD_BROYDEN ? 4.0 : 3.0);
  return B200_OK;
}

static int32_t newton_step_inner(b200_newton* nw) {
  if (nw->o.descent == B200_DESCENT_LEVENBERG_MARQUARDT) return lm_step(nw);
  if (nw->o.descent == B200_DESCENT_BROYDEN) return broyden_step(nw);
  b200_ctx* ctx = nw->ctx;
  const int64_t n = nw->n;
  const b200_newton_opts& o = nw->o;
  const double Bv = 8.0 * (double)n;
  const bool tr_on = o.globalization == B200_GLOBALIZATION_TRUST_REGION;
  const bool krylov = o.linsolve == B200_LINSOLVE_GMRES || o.linsolve == B200_LINSOLVE_SPARSE_GMRES;
  const int sch = o.tr_scheme;  // per-scheme defaults  trust_region.jl:348-381
  const double step_thr = o.tr_step_threshold > 0 ? o.tr_step_threshold : (sch == B200_TR_HEI ? 0.0 : sch == B200_TR_YUAN ? 1.0 / 1000 : sch == B200_TR_BASTIN ? 1.0 / 20 : 1.0 / 10000);
  const double shrink_thr = o.tr_shrink_threshold > 0 ? o.tr_shrink_threshold : (sch == B200_TR_HEI ? 0.0 : (sch == B200_TR_NLSOLVE || sch == B200_TR_BASTIN) ? 1.0 / 20 : 0.25);
  const double expand_thr = o.tr_expand_threshold > 0 ? o.tr_expand_threshold : ((sch == B200_TR_NLSOLVE || sch == B200_TR_BASTIN) ? 0.9 : sch == B200_TR_HEI ? 0.0 : 0.75);
  const double shrink_fac = o.tr_shrink_factor > 0 ? o.tr_shrink_factor : (sch == B200_TR_NLSOLVE ? 0.5 : sch == B200_TR_HEI ? 0.0 : sch == B200_TR_BASTIN ? 1.0 / 20 : 0.25);
  const double expand_fac = o.tr_expand_factor > 0 ? o.tr_expand_factor : 2.0;
  const int max_shrink = o.max_shrink_times > 0 ? o.max_shrink_times : 32;

  for (int attempt = 0; attempt < 2; ++attempt) {
    int new_jacobian;
    if (nw->make_new_jacobian) {  // J = cache.jac_cache(u)   solve.jl:338-344
      new_jacobian = 1;
      if (o.linsolve == B200_LINSOLVE_DENSE_LU) {
        nw->res.njacs += 1;
        B200_TRY(b200_dense_jac_fill(nw->prob, nw->u, nw->Jdense, n));  // written straight into the LU workspace (K10: no copyto!)
        nw->have_factor = 0;
      } else if (o.linsolve == B200_LINSOLVE_SPARSE_GMRES || o.linsolve == B200_LINSOLVE_SPARSE_LU) {
        nw->res.njacs += 1;
        B200_TRY(b200_sparse_jac_fill(nw->sj, nw->u, nw->nzval));
        nw->have_factor = 0;
      }
    } else {
      new_jacobian = 0;
    }
    if (o.descent == B200_DESCENT_PSEUDO_TRANSIENT && attempt == 0) {
      // SER (pseudo_transient.jl:157-170): alpha^-1 *= ||f_n|| / ||f_{n-1}||  (2-norm); A = J + alpha^-1 I
      double rn;
      B200_TRY(h_nrm2(nw, nw->fu, &rn));
      nw->alpha_inv *= rn / nw->pt_res_norm;
      nw->pt_res_norm = rn;
      nw->op.shift = nw->alpha_inv;
    }
    if (o.descent == B200_DESCENT_PSEUDO_TRANSIENT && o.linsolve == B200_LINSOLVE_DENSE_LU && new_jacobian)
      B200_TRY(b200i_diag_shift(ctx, n, nw->Jdense, n, nw->alpha_inv));  // dampen_jacobian!!: J[i,i] += alpha^-1
    if (o.forcing == B200_FORCING_EW2 && krylov) {  // pre_step_forcing!   eisenstat_walker.jl:42-80
      if (nw->nsteps == 0) {
        nw->eta = o.ew_eta0;
        B200_TRY(h_nrm2(nw, nw->fu, &nw->rnorm));
        nw->rnorm_prev = nw->rnorm;
      } else {
        const double eta_prev = nw->eta;
        nw->eta = o.ew_gamma * pow(nw->rnorm / nw->rnorm_prev, o.ew_alpha);
        if (o.ew_safeguard) {
          const double sg = o.ew_gamma * pow(eta_prev, o.ew_alpha);
          if (sg > o.ew_safeguard_threshold && sg > nw->eta) nw->eta = sg;
        }
        nw->eta = std::min(std::max(nw->eta, 0.0), o.ew_eta_max);
      }
      B200_TRY(b200_gmres_set_tolerances(nw->gm, -1.0, nw->eta));  // LinearSolve.update_tolerances!(lincache; reltol = eta)
    }
    // ---- descent: J x = fu ; du = -x   (newton.jl:97-141)
    int lin_success = 1;
    b200_gmres_stats gs;
    memset(&gs, 0, sizeof(gs));
    nw->res.nsolve += 1;
    if (o.linsolve == B200_LINSOLVE_SPARSE_LU) {
      if (!nw->have_factor) {
        nw->res.nfactors += 1;
        int32_t info = 0;
        if (o.descent == B200_DESCENT_PSEUDO_TRANSIENT) return ctx->fail(B200_ERR_UNSUPPORTED, "PseudoTransient with the sparse direct solver is not offered (use dense LU or GMRES)", __FILE__, __LINE__);
        B200_TRY(b200_sparse_lu_factor(nw->slu, nw->nzval, &info));
        nw->have_factor = 1;
        if (info != 0) lin_success = 0;
      }
      if (lin_success) B200_TRY(b200_sparse_lu_solve(nw->slu, nw->fu, nw->xlin));
    } else if (!krylov) {
      if (!nw->have_factor) {  // update_A! for a factorisation on a fresh A (…LinearSolveExt.jl:81-86)
        nw->res.nfactors += 1;
        int32_t info = 0;
        B200_TRY(b200_getrf(ctx, n, nw->Jdense, n, nw->ipiv, &info));
        nw->have_factor = 1;
        if (info != 0) lin_success = 0;
      }
      CUDA_TRY(ctx, cudaMemcpyAsync(nw->xlin, nw->fu, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
      if (lin_success) B200_TRY(b200_getrs(ctx, n, 1, nw->Jdense, n, nw->ipiv, nw->xlin, n));
      else if (n <= 4096) {
        // singular LU: LinearSolve's default dense solver falls back to a column-pivoted QR (linear_solve.jl:48-55).  The
        // factorisation destroyed J: refill it, solve in the least-squares sense (basic solution of the numerical-rank system).
        B200_TRY(b200_dense_jac_fill(nw->prob, nw->u, nw->Jdense, n));
        if (o.descent == B200_DESCENT_PSEUDO_TRANSIENT) B200_TRY(b200i_diag_shift(ctx, n, nw->Jdense, n, nw->alpha_inv));
        if (!nw->qr_work) {
          CUDA_TRY(ctx, cudaMalloc(&nw->qr_work, sizeof(double) * (3 * n + 2)));
          CUDA_TRY(ctx, cudaMalloc(&nw->qr_jpvt, sizeof(int32_t) * (n + 2)));
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(nw->qr_work + 2 * n, nw->fu, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
        int32_t rank = 0;
        B200_TRY(b200i_qrcp_solve(ctx, n, nw->Jdense, n, nw->qr_work + 2 * n, nw->xlin, nw->qr_work, nw->qr_jpvt, &rank));
        nw->have_factor = 0;  // Jdense now holds the QR factors: the next step refactorises whatever happens
        lin_success = rank > 0 ? 1 : 0;
      }
    } else {
      // `linu` aliases the du buffer: it is the initial guess only when warm_start is requested
      if (o.gmres.warm_start) CUDA_TRY(ctx, cudaMemcpyAsync(nw->xlin, nw->du, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
      if (o.precond != B200_PRECOND_NONE) {  // precs(A, p): rebuilt from the current iterate, like update_A! does for Pl / Pr
        memset(&nw->prec, 0, sizeof(nw->prec));
        nw->prec.ctx = ctx; nw->prec.n = n; nw->prec.prob = nw->prob; nw->prec.u = nw->u;
        if (nw->mg) {
          nw->prec.kind = LINOP_MULTIGRID; nw->prec.mg = nw->mg;
          if (new_jacobian) B200_TRY(b200i_mg_setup(nw->mg, nw->u));  // coarse operators follow the linearisation point
        } else {
          nw->prec.kind = LINOP_BLOCK_JACOBI;
        }
        const bool left = o.precond == B200_PRECOND_BLOCK_JACOBI_LEFT || o.precond == B200_PRECOND_MULTIGRID_LEFT;
        B200_TRY(b200_gmres_set_precond(nw->gm, left ? &nw->prec : nullptr, left ? nullptr : &nw->prec));
      }
      B200_TRY(b200_gmres_solve(nw->gm, &nw->op, nw->fu, nw->xlin, &gs));
      nw->res.njvp += gs.nmatvec;
      nw->bytes += gs.bytes;
      if (gs.status == B200_LS_NONFINITE || gs.status == B200_LS_OUT_OF_MEMORY) lin_success = 0;  // LinearSolve retcode Failure
    }
    if (!lin_success) {  // solve.jl:367-382
      if (new_jacobian) { nw->retcode = B200_RC_INTERNAL_LINSOLVE_FAILED; nw->force_stop = 1; return B200_OK; }
      nw->make_new_jacobian = 1;
      continue;  // retry once with a fresh Jacobian
    }
    // du = -x   (@. du *= -1, newton.jl:138)
    B200_TRY(b200_axpby(ctx, n, -1.0, nw->xlin, 0.0, nw->du));

    double dJJd = NAN;
    if (tr_on) {  // Dogleg   dogleg.jl:86-151
      double nrm_newton;
      B200_TRY(h_nrm2(nw, nw->du, &nrm_newton));
      if (!(nrm_newton <= nw->trust_region)) {
        B200_TRY(b200_vjp(nw->prob, nw->u, nw->fu, nw->du_c));  // du_c = -J' fu  (steepest.jl:60-80)
        B200_TRY(b200_scal(ctx, n, -1.0, nw->du_c));
        double l_grad, quad;
        B200_TRY(h_nrm2(nw, nw->du_c, &l_grad));
        B200_TRY(b200_jvp(nw->prob, nw->u, nw->du_c, nw->Jdu));
        B200_TRY(h_dot(nw, nw->Jdu, nw->Jdu, &quad));
        const double d_cauchy = (l_grad * l_grad * l_grad) / quad;
        if (d_cauchy >= nw->trust_region) {
          const double lam = nw->trust_region / l_grad;
          B200_TRY(b200_axpby(ctx, n, lam, nw->du_c, 0.0, nw->du));
          dJJd = lam * lam * quad;
        } else {
          B200_TRY(b200_axpby(ctx, n, d_cauchy / l_grad, nw->du_c, 0.0, nw->c1));  // c1 = (d_cauchy/l_grad) du_c
          B200_TRY(b200_copy(ctx, n, nw->du, nw->c2));
          B200_TRY(b200_axpy(ctx, n, -1.0, nw->c1, nw->c2));                         // c2 = du_newton - c1
          B200_TRY(b200i_reduce_sum_dev(ctx, n, nw->c2, nw->c2, RED_DOT, ctx->d_scalars));
          B200_TRY(b200i_reduce_sum_dev(ctx, n, nw->c1, nw->c2, RED_DOT, ctx->d_scalars + 1));
          B200_TRY(b200i_fetch_scalars(ctx, 2));
          const double a = ctx->h_scalars[0], b2 = ctx->h_scalars[1];
          const double b = 2.0 * b2, c = d_cauchy * d_cauchy - nw->trust_region * nw->trust_region;
          const double aux = std::max(0.0, b * b - 4.0 * a * c);
          const double tau = (-b + sqrt(aux)) / (2.0 * a);
          B200_TRY(b200_copy(ctx, n, nw->c1, nw->du));
          B200_TRY(b200_axpy(ctx, n, tau, nw->c2, nw->du));
        }
      }
    }
    if (o.forcing == B200_FORCING_EW2 && krylov) {  // post_step_forcing!  eisenstat_walker.jl:83-87 (fu is still the old residual)
      nw->rnorm_prev = nw->rnorm;
      B200_TRY(h_nrm2(nw, nw->fu, &nw->rnorm));
    }
    nw->make_new_jacobian = 1;
    int accepted = 1;
    double objective, du_norm;
    double ls_alpha = 1.0;
    if (o.globalization == B200_GLOBALIZATION_LINESEARCH) {
      // solve.jl:392-408 with LineSearch.jl BackTracking (cubic interpolation; external package restated, see oracle.c):
      // phi(a) = ||f(u + a du)||^2 / 2, phi'(0) = <fu, J du> by one JVP; every trial is one axpby + one residual + one norm.
      const double c1 = o.ls_c1 > 0 ? o.ls_c1 : 1e-4, rho_hi = o.ls_rho_hi > 0 ? o.ls_rho_hi : 0.5, rho_lo = o.ls_rho_lo > 0 ? o.ls_rho_lo : 0.1;
      const int ls_max = o.ls_maxiters > 0 ? o.ls_maxiters : 1000;
      double nf0, dphi0;
      B200_TRY(h_nrm2(nw, nw->fu, &nf0));
      const double phi0 = 0.5 * nf0 * nf0;
      B200_TRY(b200_jvp(nw->prob, nw->u, nw->du, nw->Jdu));
      B200_TRY(h_dot(nw, nw->fu, nw->Jdu, &dphi0));
      auto phi = [&](double alpha, double* out) -> int32_t {
        B200_TRY(b200_copy(ctx, n, nw->u, nw->u_trial));
        B200_TRY(b200_axpy(ctx, n, alpha, nw->du, nw->u_trial));
        B200_TRY(b200_residual(nw->prob, nw->u_trial, nw->fu_trial));
        nw->res.nf += 1;
        double t;
        B200_TRY(h_nrm2(nw, nw->fu_trial, &t));
        *out = 0.5 * t * t;
        return B200_OK;
      };
      double a1 = 1.0, a2 = 1.0, phx0 = phi0, phx1;
      B200_TRY(phi(a1, &phx1));
      int itf = 0;
      while (!std::isfinite(phx1) && itf < 50) { ++itf; a1 = a2; a2 = a1 / 2.0; B200_TRY(phi(a2, &phx1)); }
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
          else { const double dd = std::max(cb * cb - 3.0 * ca * dphi0, 0.0); at = (-cb + sqrt(dd)) / (3.0 * ca); }
        }
        a1 = a2;
        at = std::min(at, a2 * rho_hi);
        a2 = std::max(at, a2 * rho_lo);
        phx0 = phx1;
        B200_TRY(phi(a2, &phx1));
      }
      if (ls_failed) { nw->retcode = B200_RC_INTERNAL_LINESEARCH_FAILED; nw->force_stop = 1; }
      ls_alpha = a2;
      CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_scalars, 0, sizeof(double) * 2, ctx->stream));
      B200_TRY(b200i_axpy_norm(ctx, n, a2, nw->du, nw->u, ctx->d_scalars + 1));   // @bb axpy!(alpha, du, u)   solve.jl:403
      B200_TRY(b200i_residual_norm(nw->prob, nw->u, nw->fu, ctx->d_scalars));     // evaluate_f!               solve.jl:407
      nw->res.nf += 1;
      B200_TRY(b200i_fetch_scalars(ctx, 2));
      objective = ctx->h_scalars[0];
      du_norm = sqrt(ctx->h_scalars[1]);
    } else if (!tr_on) {  // solve.jl:436-445 : u += du ; fu = f(u) ; with ||du||^2 and ||fu||_inf produced by the kernels' epilogues
      CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_scalars, 0, sizeof(double) * 2, ctx->stream));
      B200_TRY(b200i_axpy_norm(ctx, n, 1.0, nw->du, nw->u, ctx->d_scalars + 1));
      B200_TRY(b200i_residual_norm(nw->prob, nw->u, nw->fu, ctx->d_scalars));
      nw->res.nf += 1;
      B200_TRY(b200i_fetch_scalars(ctx, 2));
      objective = ctx->h_scalars[0];
      du_norm = sqrt(ctx->h_scalars[1]);
      nw->bytes += 5.0 * Bv;
    } else {  // GenericTrustRegionScheme solve!   trust_region.jl:396-430, 511-513
      B200_TRY(b200_copy(ctx, n, nw->u, nw->u_trial));
      B200_TRY(b200_axpy(ctx, n, 1.0, nw->du, nw->u_trial));
      CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_scalars, 0, sizeof(double) * 2, ctx->stream));
      B200_TRY(b200i_residual_norm(nw->prob, nw->u_trial, nw->fu_trial, ctx->d_scalars));
      nw->res.nf += 1;
      // the six scalars of the acceptance test are independent: reduced on the device one after the other, fetched ONCE
      // (round 1 synchronised the host after each of them)
      const bool need_dJJd = dJJd != dJJd;
      if (need_dJJd) {
        B200_TRY(b200_jvp(nw->prob, nw->u, nw->du, nw->Jdu));
        B200_TRY(b200i_reduce_sum_dev(ctx, n, nw->Jdu, nw->Jdu, RED_DOT, ctx->d_scalars + 1));
      }
      B200_TRY(b200_vjp(nw->prob, nw->u, nw->fu, nw->JTfu));
      B200_TRY(b200i_reduce_sum_dev(ctx, n, nw->fu_trial, nullptr, RED_SUMSQ, ctx->d_scalars + 2));
      B200_TRY(b200i_reduce_sum_dev(ctx, n, nw->fu, nullptr, RED_SUMSQ, ctx->d_scalars + 3));
      B200_TRY(b200i_reduce_sum_dev(ctx, n, nw->du, nw->JTfu, RED_DOT, ctx->d_scalars + 4));
      B200_TRY(b200i_reduce_sum_dev(ctx, n, nw->du, nullptr, RED_SUMSQ, ctx->d_scalars + 5));
      B200_TRY(b200i_fetch_scalars(ctx, 6));
      const double trial_inf = ctx->h_scalars[0];
      if (need_dJJd) dJJd = ctx->h_scalars[1];
      const double nt = sqrt(ctx->h_scalars[2]), nc = sqrt(ctx->h_scalars[3]), dg = ctx->h_scalars[4];
      const double num = (nt * nt - nc * nc) / 2.0;
      const double denom = dg + dJJd / 2.0;
      const double rho = num / denom;
      accepted = rho > step_thr;
      const double dun = sqrt(ctx->h_scalars[5]);  // internalnorm(du)
      double& tr = nw->trust_region;
      if (sch == B200_TR_SIMPLE) {  // trust_region.jl:431-440
        if (rho < shrink_thr) { tr *= shrink_fac; nw->shrink_counter += 1; }
        else { nw->shrink_counter = 0; if (rho > expand_thr && rho > step_thr) tr = expand_fac * tr; }
      } else if (sch == B200_TR_NLSOLVE) {  // :441-455
        if (rho < shrink_thr) { tr *= shrink_fac; nw->shrink_counter += 1; }
        else {
          nw->shrink_counter = 0;
          if (rho >= expand_thr) tr = expand_fac * dun;
          else if (rho >= nw->tp1) tr = std::max(tr, expand_fac * dun);
        }
      } else if (sch == B200_TR_NOCEDAL_WRIGHT) {  // :456-466
        if (rho < shrink_thr) { tr = shrink_fac * dun; nw->shrink_counter += 1; }
        else { nw->shrink_counter = 0; if (rho > expand_thr && fabs(dun - tr) < 1.0e-6 * tr) tr = expand_fac * tr; }
      } else if (sch == B200_TR_HEI) {  // :467-476, rfunc_adaptive_trust_region :383-391
        const double M = nw->tp1, g1 = nw->tp3, g2 = nw->tp4, beta = nw->tp2;
        const double rf = (rho >= shrink_thr) ? (2.0 * (M - 1.0 - g2) * atan(rho - shrink_thr) + (1.0 + g2)) / M_PI
                                              : (1.0 - g1 - beta) * (exp(rho - shrink_thr) + beta / (1.0 - g1 - beta));
        const double tr_new = rf * dun;
        if (tr_new < tr) nw->shrink_counter += 1; else nw->shrink_counter = 0;
        tr = tr_new;
      } else if (sch == B200_TR_YUAN) {  // :477-490
        if (rho < shrink_thr) { nw->tp1 = nw->tp2 * nw->tp1; nw->shrink_counter += 1; }
        else { if (rho >= expand_thr && 2.0 * dun > tr) nw->tp1 = nw->tp3 * nw->tp1; nw->shrink_counter = 0; }
        double g;
        B200_TRY(b200_vjp(nw->prob, nw->u_trial, nw->fu_trial, nw->JTfu));
        B200_TRY(h_nrm2(nw, nw->JTfu, &g));
        tr = nw->tp1 * g;
      } else if (sch == B200_TR_FAN) {  // :491-499
        if (rho < shrink_thr) { nw->tp1 *= nw->tp2; nw->shrink_counter += 1; }
        else { nw->shrink_counter = 0; if (rho > expand_thr) nw->tp1 = std::min(nw->tp1 * nw->tp3, nw->tp4); }
        tr = nw->tp1 * pow(nt, 0.99);
      } else if (sch == B200_TR_BASTIN) {  // :500-520, retrospective ratio at the trial point with the step just taken
        if (rho > step_thr) {
          double d1, d2;
          B200_TRY(b200_jvp(nw->prob, nw->u_trial, nw->du, nw->Jdu));
          B200_TRY(b200_vjp(nw->prob, nw->u_trial, nw->fu_trial, nw->JTfu));
          B200_TRY(h_dot(nw, nw->JTfu, nw->JTfu, &d1));
          B200_TRY(b200_vjp(nw->prob, nw->u_trial, nw->Jdu, nw->JTfu));
          B200_TRY(h_dot(nw, nw->JTfu, nw->JTfu, &d2));
          const double rho_retro = num / (d1 + d2 / 2.0);
          if (rho_retro >= expand_thr) tr = nw->tp1 * dun;
          nw->shrink_counter = 0;
        } else { tr *= nw->tp2; nw->shrink_counter += 1; }
      }
      nw->trust_region = std::min(nw->trust_region, nw->max_tr);
      if (accepted) {
        du_norm = dun;
        std::swap(nw->u, nw->u_trial);      // copyto!(cache.u, u_new) as a pointer swap
        std::swap(nw->fu, nw->fu_trial);
        nw->op.u = nw->u;
        objective = trial_inf;
      } else {
        nw->make_new_jacobian = 0;
        objective = nw->fnorm_inf;
        du_norm = 0.0;
      }
      if (nw->shrink_counter > max_shrink) { nw->retcode = B200_RC_SHRINK_THRESHOLD_EXCEEDED; nw->force_stop = 1; }
    }
    nw->fnorm_inf = objective;
    bool new_best = false;
    TermQuant tq;
    B200_TRY(term_quantities(nw, nw->fu, nw->u, objective, &tq));
    if (term_check(nw, tq, du_norm, &new_best)) { nw->retcode = nw->tc.retcode; nw->force_stop = 1; }  // check_and_update!
    if (new_best && nw->best_u) CUDA_TRY(ctx, cudaMemcpyAsync(nw->best_u, nw->u, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    if (o.store_trace) {
      b200_trace_rec t;
      t.iter = nw->nsteps + 1; t.lin_iters = gs.iters; t.lin_status = gs.status; t.accepted = accepted;
      t.fnorm_inf = objective; t.step_norm2 = du_norm; t.lin_rnorm = gs.rnorm; t.trust_radius = (o.globalization == B200_GLOBALIZATION_LINESEARCH) ? ls_alpha : (o.descent == B200_DESCENT_PSEUDO_TRANSIENT) ? 1.0 / nw->alpha_inv : nw->trust_region;
      nw->trace.push_back(t);
    }
    // copyto!(u_cache, u) (solve.jl:460) is only ever read back as `uprev` for the stall norm, which the update kernel
    // already produced; the copy is skipped.
    return B200_OK;
  }
  return B200_OK;
}

// CommonSolve.step! (NonlinearSolveBase/src/solve.jl:835-858): one step, the counters, and the wall-clock limit
static int32_t newton_timed_step(b200_newton* nw) {
  const bool limited = nw->o.maxtime > 0.0;
  std::chrono::steady_clock::time_point t0;
  if (limited) t0 = std::chrono::steady_clock::now();
  B200_TRY(newton_step_inner(nw));
  nw->res.nsteps += 1;
  nw->nsteps += 1;
  if (limited) {
    CUDA_TRY(nw->ctx, cudaStreamSynchronize(nw->ctx->stream));
    nw->total_time += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!nw->force_stop && nw->retcode == B200_RC_DEFAULT && nw->total_time >= nw->o.maxtime) { nw->retcode = B200_RC_MAXTIME; nw->force_stop = 1; }
  }
  return B200_OK;
}

int32_t b200_newton_step(b200_newton* nw, int32_t* terminated_host) {
  B200_DEVICE_GUARD(nw ? nw->ctx : nullptr);
  b200_ctx* ctx = nw->ctx;
  B200_REQUIRE(ctx, nw->initialised, "newton_step before newton_reinit");
  if (!(nw->force_stop || nw->nsteps >= nw->maxiters)) {  // not_terminated  abstract_types.jl:722-724
    B200_TRY(newton_timed_step(nw));
  }
  if (terminated_host) *terminated_host = (nw->force_stop || nw->nsteps >= nw->maxiters) ? 1 : 0;
  return B200_OK;
}

int32_t b200_newton_result_get(b200_newton* nw, b200_newton_result* r) {
  *r = nw->res;
  r->retcode = nw->retcode;
  r->resid_inf = nw->fnorm_inf;
  r->ntrace = (int32_t)nw->trace.size();
  r->bytes = nw->bytes;
  return B200_OK;
}

int32_t b200_newton_solve(b200_newton* nw, b200_newton_result* result) {
  B200_DEVICE_GUARD(nw ? nw->ctx : nullptr);
  b200_ctx* ctx = nw->ctx;
  B200_REQUIRE(ctx, nw->initialised, "newton_solve before newton_reinit");
  while (!nw->force_stop && nw->nsteps < nw->maxiters) B200_TRY(newton_timed_step(nw));
  if (nw->retcode == B200_RC_DEFAULT) nw->retcode = (nw->nsteps >= nw->maxiters) ? B200_RC_MAXITERS : B200_RC_SUCCESS;  // solve.jl:372-376
  // update_from_termination_cache! for Best modes: roll back to the best iterate (termination_conditions.jl:440-453)
  if (nw->best_u) {
    int32_t same = 1;
    B200_TRY(b200_equal(ctx, nw->n, nw->u, nw->best_u, &same));
    if (!same) {
      CUDA_TRY(ctx, cudaMemcpyAsync(nw->u, nw->best_u, sizeof(double) * nw->n, cudaMemcpyDeviceToDevice, ctx->stream));
      CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_scalars, 0, sizeof(double), ctx->stream));
      B200_TRY(b200i_residual_norm(nw->prob, nw->u, nw->fu, ctx->d_scalars));
      nw->res.nf += 1;
      B200_TRY(b200i_fetch_scalars(ctx, 1));
      nw->fnorm_inf = ctx->h_scalars[0];
    }
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (result) B200_TRY(b200_newton_result_get(nw, result));
  return B200_OK;
}

int32_t b200_newton_u(b200_newton* nw, double** u_dev) { *u_dev = nw->u; return B200_OK; }
int32_t b200_newton_fu(b200_newton* nw, double** fu_dev) { *fu_dev = nw->fu; return B200_OK; }

int32_t b200_newton_trace(b200_newton* nw, b200_trace_rec* recs, int32_t cap, int32_t* count) {
  int32_t m = (int32_t)std::min<size_t>(nw->trace.size(), (size_t)std::max(cap, 0));
  for (int32_t i = 0; i < m; ++i) recs[i] = nw->trace[i];
  if (count) *count = m;
  return B200_OK;
}

int32_t b200_newton_solve_host(b200_newton* nw, const double* u0_host, double* u_host, double* resid_host, b200_newton_result* result) {
  B200_DEVICE_GUARD(nw ? nw->ctx : nullptr);
  b200_ctx* ctx = nw->ctx;
  const size_t bytes = sizeof(double) * nw->n;
  CUDA_TRY(ctx, cudaMemcpyAsync(nw->u, u0_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  B200_TRY(b200_newton_reinit(nw, nw->u));
  B200_TRY(b200_newton_solve(nw, result));
  if (u_host) CUDA_TRY(ctx, cudaMemcpyAsync(u_host, nw->u, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (resid_host) CUDA_TRY(ctx, cudaMemcpyAsync(resid_host, nw->fu, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
}  // extern "C"
