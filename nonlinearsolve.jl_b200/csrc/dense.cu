// dense.cu — dense Jacobian fallback (kernels K7, K8): analytic fill of the column-major n x n
// Jacobian and a blocked right-looking LU with partial pivoting (LAPACK getrf/getrs semantics).
//
// Reference path replaced: JacobianCache call lib/NonlinearSolveBase/src/jacobian.jl:237-258 (DI.jacobian! with dense
// AutoForwardDiff: exact derivatives) and the LinearSolve LU reached through linear_solve.jl:100-117 /
// NonlinearSolveBaseLinearSolveExt.jl:16-32, 102-111 (copyto!(A, J); lu!; ldiv!).
//
// Structure of getrf: for each panel of NB columns — (1) panel factorisation on CUDA cores (pivot search by block
// arg-max, row swap, scale, rank-1 update inside the panel), (2) the panel's row interchanges applied to the columns left
// and right of it, (3) TRSM  U12 = L11^{-1} A12 with L11 staged in shared memory, (4) the trailing update
// A22 -= L21 U12 — the only GEMM-shaped part — on the FP64 tensor cores (mma.sync.m8n8k4.f64 = DMMA; tcgen05 has no
// FP64 kind), operands staged through shared memory.
#include "common.cuh"
#include <math.h>
#include <algorithm>

namespace {
constexpr int NB = 32;         // panel width
constexpr int DT = 256;

// ------------------------------------------------------------------ Jacobian fill
__global__ void __launch_bounds__(DT) bruss_dense_fill_kernel(int dim, int N, double a, double A, const double* __restrict__ u,
                                                               double* __restrict__ J, int64_t ld) {
  const int64_t N2 = (int64_t)N * N, NC = (dim == 2) ? N2 : N2 * N, n = 2 * NC;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // column
  if (c >= n) return;
  const int s = (int)(c / NC);
  const int64_t cell = c - (int64_t)s * NC;
  const int i = (int)(cell % N), j = (int)((cell / N) % N), k = (dim == 3) ? (int)(cell / N2) : 0;
  const int ip = (i + 1 == N) ? 0 : i + 1, im = (i == 0) ? N - 1 : i - 1;
  const int jp = (j + 1 == N) ? 0 : j + 1, jm = (j == 0) ? N - 1 : j - 1;
  const int kp = (k + 1 == N) ? 0 : k + 1, km = (k == 0) ? N - 1 : k - 1;
  const double uc = u[cell], vc = u[cell + NC];
  const double diag_lap = (dim == 2) ? -4.0 * a : -6.0 * a;
  double dself, dcross;
  if (s == 0) { dself = 2.0 * uc * vc - (A + 1.0); dcross = A - 2.0 * uc * vc; }
  else { dself = -(uc * uc); dcross = uc * uc; }
  double* col = J + c * ld;
  const int64_t off = (int64_t)s * NC, offx = (int64_t)(1 - s) * NC;
  // neighbours first (+=, they may coincide for tiny N), column was zeroed by the caller
  col[im + (int64_t)N * j + N2 * k + off] += a;
  col[ip + (int64_t)N * j + N2 * k + off] += a;
  col[i + (int64_t)N * jm + N2 * k + off] += a;
  col[i + (int64_t)N * jp + N2 * k + off] += a;
  if (dim == 3) {
    col[i + (int64_t)N * j + N2 * km + off] += a;
    col[i + (int64_t)N * j + N2 * kp + off] += a;
  }
  col[cell + off] += diag_lap + dself;
  col[cell + offx] += dcross;
}
__global__ void __launch_bounds__(DT) small_dense_fill_kernel(int kind, int64_t n, const double* __restrict__ u, double* __restrict__ J, int64_t ld) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double* col = J + c * ld;
  if (kind == 0) {
    col[c] = 2.0 * u[c];
  } else {  // F_i = u_i + 0.1 u_i (2 u_i - u_{i-1} - u_{i+1}) - p_i
    double t = 2.0 * u[c];
    if (c > 0) t -= u[c - 1];
    if (c + 1 < n) t -= u[c + 1];
    col[c] = 1.0 + 0.1 * (t + 2.0 * u[c]);
    if (c > 0) col[c - 1] = -0.1 * u[c - 1];
    if (c + 1 < n) col[c + 1] = -0.1 * u[c + 1];
  }
}
__global__ void __launch_bounds__(DT) unit_vector_kernel(int64_t n, int64_t j, double* __restrict__ e) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) e[i] = (i == j) ? 1.0 : 0.0;
}

// ------------------------------------------------------------------ LU panel kernels (single CTA each)
// pivot search + swap inside the panel + scale + rank-1 update of the remaining panel columns, for panel column jj
__global__ void __launch_bounds__(1024) panel_column_kernel(int64_t n, double* __restrict__ A, int64_t ld, int64_t k0, int kb, int jj,
                                                             int64_t* __restrict__ ipiv, int* __restrict__ info) {
  __shared__ double smax[32];
  __shared__ int64_t sidx[32];
  __shared__ int64_t piv_s;
  __shared__ double pivval_s;
  const int64_t k = k0 + jj;
  double* colk = A + k * ld;
  // arg-max |A[i,k]|, i >= k ; ties -> smallest index (LAPACK idamax)
  double best = -1.0;
  int64_t bidx = k;
  for (int64_t i = k + threadIdx.x; i < n; i += blockDim.x) {
    const double v = fabs(colk[i]);
    if (v > best || (v != v && best == best)) { best = v; bidx = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int64_t oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { smax[wid] = best; sidx[wid] = bidx; }
  __syncthreads();
  if (wid == 0) {
    best = (lane < nw) ? smax[lane] : -2.0;
    bidx = (lane < nw) ? sidx[lane] : k;
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    if (lane == 0) {
      piv_s = bidx;
      ipiv[k] = bidx + 1;
      pivval_s = colk[bidx];
      if (colk[bidx] == 0.0 && *info == 0) *info = (int)(k + 1);
    }
  }
  __syncthreads();
  const int64_t piv = piv_s;
  const double pivval = pivval_s;
  // swap rows k and piv inside the panel
  if (piv != k) {
    for (int c = threadIdx.x; c < kb; c += blockDim.x) {
      double* col = A + (k0 + c) * ld;
      const double t = col[k];
      col[k] = col[piv];
      col[piv] = t;
    }
  }
  __syncthreads();
  if (pivval == 0.0) return;
  const double inv = 1.0 / pivval;
  // scale + rank-1 update of panel columns jj+1..kb-1
  for (int64_t i = k + 1 + threadIdx.x; i < n; i += blockDim.x) {
    const double l = colk[i] * inv;
    colk[i] = l;
    for (int c = jj + 1; c < kb; ++c) {
      double* col = A + (k0 + c) * ld;
      col[i] = fma(-l, col[k], col[i]);
    }
  }
}

// apply the panel's interchanges to every column outside the panel (thread per column)
__global__ void __launch_bounds__(DT) swap_rows_kernel(int64_t n, double* __restrict__ A, int64_t ld, int64_t k0, int kb,
                                                        const int64_t* __restrict__ ipiv) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n - kb) return;
  if (c >= k0) c += kb;  // skip the panel's own columns
  double* col = A + c * ld;
  for (int q = 0; q < kb; ++q) {
    const int64_t r = k0 + q, p = ipiv[r] - 1;
    if (p != r) {
      const double t = col[r];
      col[r] = col[p];
      col[p] = t;
    }
  }
}

// U12 = L11^{-1} A12 : thread per column of A12, L11 (unit lower, kb x kb) in shared memory
__global__ void __launch_bounds__(DT) trsm_kernel(int64_t n, double* __restrict__ A, int64_t ld, int64_t k0, int kb) {
  __shared__ double L[NB][NB + 1];
  for (int t = threadIdx.x; t < kb * kb; t += blockDim.x) {
    const int r = t % kb, c = t / kb;
    L[r][c] = A[(k0 + c) * ld + k0 + r];
  }
  __syncthreads();
  const int64_t c = k0 + kb + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double* col = A + c * ld + k0;
  double x[NB];
#pragma unroll
  for (int r = 0; r < NB; ++r) x[r] = (r < kb) ? col[r] : 0.0;
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    if (r < kb) {
      double s = x[r];
#pragma unroll
      for (int q = 0; q < NB; ++q)
        if (q < r) s = fma(-L[r][q], x[q], s);
      x[r] = s;
    }
  }
#pragma unroll
  for (int r = 0; r < NB; ++r)
    if (r < kb) col[r] = x[r];
}

// ------------------------------------------------------------------ trailing update on the FP64 tensor cores
// C[m x nn] -= Lp[m x kb] * U[kb x nn], all column-major inside A.  CTA tile 64 x 64, 8 warps laid out 4 (m) x 2 (n),
// each warp owns a 16 x 32 sub-tile = 2 x 4 DMMA m8n8k4 accumulator fragments; the K = kb <= 32 slab of both operands
// is staged once in shared memory (padded against bank conflicts).
constexpr int GT_M = 64, GT_N = 64;
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(DT) trailing_dmma_kernel(int64_t n, double* __restrict__ A, int64_t ld, int64_t k0, int kb) {
  __shared__ double Ls[NB][GT_M + 4];  // Ls[k][i] = L21[i, k]
  __shared__ double Us[NB][GT_N + 4];  // Us[k][j] = U12[k, j]
  const int64_t k1 = k0 + kb;
  const int64_t i0 = k1 + (int64_t)blockIdx.x * GT_M;
  const int64_t j0 = k1 + (int64_t)blockIdx.y * GT_N;
  for (int t = threadIdx.x; t < NB * GT_M; t += DT) {
    const int i = t % GT_M, k = t / GT_M;
    Ls[k][i] = (k < kb && i0 + i < n) ? A[(k0 + k) * ld + i0 + i] : 0.0;
  }
  for (int t = threadIdx.x; t < NB * GT_N; t += DT) {
    const int k = t % NB, j = t / NB;
    Us[k][j] = (k < kb && j0 + j < n) ? A[(j0 + j) * ld + k0 + k] : 0.0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
  const int g = lane >> 2, t4 = lane & 3;  // groupID, thread-in-group
  double acc[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
#pragma unroll
  for (int kk = 0; kk < NB; kk += 4) {
    double af[2], bf[4];
#pragma unroll
    for (int a = 0; a < 2; ++a) af[a] = Ls[kk + t4][wm + a * 8 + g];      // A fragment: row g, col t4 (row-major m8 x k4)
#pragma unroll
    for (int b = 0; b < 4; ++b) bf[b] = Us[kk + t4][wn + b * 8 + g];      // B fragment: row t4, col g (col-major k4 x n8)
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) dmma_m8n8k4(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
  }
  // C fragment: row g, cols 2*t4, 2*t4+1
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t i = i0 + wm + a * 8 + g;
      const int64_t j = j0 + wn + b * 8 + 2 * t4;
      if (i < n) {
        if (j < n) A[j * ld + i] -= acc[a][b][0];
        if (j + 1 < n) A[(j + 1) * ld + i] -= acc[a][b][1];
      }
    }
}

// ------------------------------------------------------------------ triangular solves (nrhs small)
__global__ void __launch_bounds__(DT) apply_pivots_kernel(int64_t n, int64_t nrhs, const int64_t* __restrict__ ipiv, double* __restrict__ B, int64_t ldb) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrhs) return;
  double* b = B + r * ldb;
  for (int64_t k = 0; k < n; ++k) {
    const int64_t p = ipiv[k] - 1;
    if (p != k) { const double t = b[k]; b[k] = b[p]; b[p] = t; }
  }
}
// one block step of the forward (unit lower) or backward (upper) substitution: solve the NB x NB diagonal block in a
// single warp, then subtract its contribution from the rest of the right-hand side (all CTAs).
__global__ void __launch_bounds__(DT) trisolve_diag_kernel(int lower, int64_t n, const double* __restrict__ A, int64_t ld, int64_t k0, int kb,
                                                            double* __restrict__ b) {
  __shared__ double T[NB][NB + 1];
  __shared__ double x[NB];
  for (int t = threadIdx.x; t < kb * kb; t += blockDim.x) {
    const int r = t % kb, c = t / kb;
    T[r][c] = A[(k0 + c) * ld + k0 + r];
  }
  for (int t = threadIdx.x; t < kb; t += blockDim.x) x[t] = b[k0 + t];
  __syncthreads();
  if (threadIdx.x == 0) {
    if (lower) {
      for (int r = 0; r < kb; ++r) { double s = x[r]; for (int q = 0; q < r; ++q) s -= T[r][q] * x[q]; x[r] = s; }
    } else {
      for (int r = kb - 1; r >= 0; --r) { double s = x[r]; for (int q = r + 1; q < kb; ++q) s -= T[r][q] * x[q]; x[r] = s / T[r][r]; }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < kb; t += blockDim.x) b[k0 + t] = x[t];
}
__global__ void __launch_bounds__(DT) trisolve_update_kernel(int lower, int64_t n, const double* __restrict__ A, int64_t ld, int64_t k0, int kb,
                                                              double* __restrict__ b) {
  __shared__ double x[NB];
  for (int t = threadIdx.x; t < kb; t += blockDim.x) x[t] = b[k0 + t];
  __syncthreads();
  const int64_t lo = lower ? k0 + kb : 0, hi = lower ? n : k0;
  const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  double s = b[i];
  for (int q = 0; q < kb; ++q) s = fma(-A[(k0 + q) * ld + i], x[q], s);
  b[i] = s;
}
}  // namespace

extern "C" {
int32_t b200_dense_jac_fill(b200_problem* p, const double* u, double* J, int64_t ld) {
  b200_ctx* ctx = p->ctx;
  const int64_t n = p->n;
  B200_REQUIRE(ctx, ld >= n, "dense_jac_fill: ld < n");
  if (p->kind != B200_PROB_CALLBACK) CUDA_TRY(ctx, cudaMemsetAsync(J, 0, sizeof(double) * ld * n, ctx->stream));
  const int grid = (int)((n + DT - 1) / DT);
  switch (p->kind) {
    case B200_PROB_BRUSS2D:
    case B200_PROB_BRUSS3D:
      LAUNCH(ctx, bruss_dense_fill_kernel, grid, DT, 0, p->kind == B200_PROB_BRUSS2D ? 2 : 3, p->N, p->a, p->A, u, J, ld);
      break;
    case B200_PROB_QUADRATIC:
    case B200_PROB_TRIDIAG_QUAD:
      LAUNCH(ctx, small_dense_fill_kernel, grid, DT, 0, p->kind == B200_PROB_QUADRATIC ? 0 : 1, n, u, J, ld);
      break;
    case B200_PROB_CALLBACK: {  // n operator applications on unit vectors (what chunk-1 forward AD would do)
      double* e = nullptr;
      CUDA_TRY(ctx, cudaMalloc(&e, sizeof(double) * n));
      for (int64_t j = 0; j < n; ++j) {
        LAUNCH(ctx, unit_vector_kernel, grid, DT, 0, n, j, e);
        int32_t s = b200_jvp(p, u, e, J + j * ld);
        if (s != B200_OK) { cudaFree(e); return s; }
      }
      cudaStreamSynchronize(ctx->stream);
      cudaFree(e);
    } break;
    default: return ctx->fail(B200_ERR_INVALID, "unknown problem kind", __FILE__, __LINE__);
  }
  CHECK_LAUNCH(ctx);
  return B200_OK;
}

int32_t b200_getrf(b200_ctx* ctx, int64_t n, double* A, int64_t ld, int64_t* ipiv, int32_t* info_host) {
  B200_REQUIRE(ctx, n > 0 && ld >= n, "getrf: bad dimensions");
  int* d_info = reinterpret_cast<int*>(ctx->d_scalars + 16);
  CUDA_TRY(ctx, cudaMemsetAsync(d_info, 0, sizeof(int), ctx->stream));
  for (int64_t k0 = 0; k0 < n; k0 += NB) {
    const int kb = (int)std::min<int64_t>(NB, n - k0);
    const int pthreads = (n - k0 >= 8192) ? 1024 : (n - k0 >= 1024 ? 512 : 128);
    for (int jj = 0; jj < kb; ++jj) LAUNCH(ctx, panel_column_kernel, 1, pthreads, 0, n, A, ld, k0, kb, jj, ipiv, d_info);
    if (n - kb > 0) LAUNCH(ctx, swap_rows_kernel, (int)((n - kb + DT - 1) / DT), DT, 0, n, A, ld, k0, kb, (const int64_t*)ipiv);
    const int64_t rest = n - (k0 + kb);
    if (rest > 0) {
      LAUNCH(ctx, trsm_kernel, (int)((rest + DT - 1) / DT), DT, 0, n, A, ld, k0, kb);
      dim3 grid((unsigned)((rest + GT_M - 1) / GT_M), (unsigned)((rest + GT_N - 1) / GT_N));
      LAUNCH(ctx, trailing_dmma_kernel, grid, DT, 0, n, A, ld, k0, kb);
    }
  }
  CHECK_LAUNCH(ctx);
  if (info_host) {
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalars + 16, d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *info_host = *reinterpret_cast<int*>(ctx->h_scalars + 16);
  }
  return B200_OK;
}

int32_t b200_getrs(b200_ctx* ctx, int64_t n, int64_t nrhs, const double* A, int64_t ld, const int64_t* ipiv, double* B, int64_t ldb) {
  B200_REQUIRE(ctx, n > 0 && nrhs > 0 && ld >= n && ldb >= n, "getrs: bad dimensions");
  LAUNCH(ctx, apply_pivots_kernel, (int)((nrhs + DT - 1) / DT), DT, 0, n, nrhs, ipiv, B, ldb);
  for (int64_t r = 0; r < nrhs; ++r) {
    double* b = B + r * ldb;
    for (int64_t k0 = 0; k0 < n; k0 += NB) {  // L y = P b
      const int kb = (int)std::min<int64_t>(NB, n - k0);
      LAUNCH(ctx, trisolve_diag_kernel, 1, 64, 0, 1, n, A, ld, k0, kb, b);
      const int64_t rest = n - (k0 + kb);
      if (rest > 0) LAUNCH(ctx, trisolve_update_kernel, (int)((rest + DT - 1) / DT), DT, 0, 1, n, A, ld, k0, kb, b);
    }
    const int64_t nblk = (n + NB - 1) / NB;
    for (int64_t blk = nblk - 1; blk >= 0; --blk) {  // U x = y
      const int64_t k0 = blk * NB;
      const int kb = (int)std::min<int64_t>(NB, n - k0);
      LAUNCH(ctx, trisolve_diag_kernel, 1, 64, 0, 0, n, A, ld, k0, kb, b);
      if (k0 > 0) LAUNCH(ctx, trisolve_update_kernel, (int)((k0 + DT - 1) / DT), DT, 0, 0, n, A, ld, k0, kb, b);
    }
  }
  CHECK_LAUNCH(ctx);
  return B200_OK;
}
}  // extern "C"
