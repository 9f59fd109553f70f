// ens_batched.cu — batched per-CTA Newton-GMRES engine for the ensemble path (placeholder until the kernel lands).
#include "common.cuh"

int32_t b200i_ens_batched_supported(int32_t N, const b200_newton_opts* o) { (void)N; (void)o; return 0; }
int32_t b200i_ens_batched_solve(b200_ctx* ctx, int32_t, int32_t, double, const b200_newton_opts*, const double*, const double*, const double*,
                                double*, double*, int32_t*, int32_t*, int32_t*, void**, size_t*) {
  return ctx->fail(B200_ERR_UNSUPPORTED, "batched ensemble engine not available", __FILE__, __LINE__);
}
