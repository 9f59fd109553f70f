// sparse.cu — sparse Jacobian fallback (kernels K3, K9): structural pattern (CSC), greedy column
// colouring, compressed Jacobian evaluation by `ncolors` seeded JVP sweeps + decompression into nzval, SpMV / SpMV'.
//
// Reference path replaced: construct_concrete_adtype lib/NonlinearSolveBase/src/jacobian.jl:286-353 (AutoSparse with a
// KnownJacobianSparsityDetector / TracerSparsityDetector and the colouring algorithm selected at
// ext/NonlinearSolveBaseSparseMatrixColoringsExt.jl:13-28), J allocation jacobian.jl:119-125, refresh DI.jacobian!
// jacobian.jl:244-247.  Index arrays are bit-exact with the CPU oracle (sorted rows per column; colours 1-based).
// The pattern and the colouring are set-up work done once on the host; the per-step work (seeded JVPs, scatter, SpMV)
// runs on the device.
#include "common.cuh"
#include <algorithm>
#include <vector>

namespace {
constexpr int ST = 256;

// structural entries of column c, sorted and de-duplicated (what SparseConnectivityTracer's global detector reports
// for the residual expressions: every syntactic dependency, independent of the value of u)
int column_rows(const b200_problem* p, int64_t c, int64_t* rows) {
  int cnt = 0;
  if (p->kind == B200_PROB_BRUSS2D || p->kind == B200_PROB_BRUSS3D) {
    const int N = p->N;
    const int dim = (p->kind == B200_PROB_BRUSS2D) ? 2 : 3;
    const int64_t N2 = (int64_t)N * N, NC = (dim == 2) ? N2 : N2 * N;
    const int s = (int)(c / NC);
    const int64_t cell = c % NC;
    const int i = (int)(cell % N), j = (int)((cell / N) % N), k = (dim == 3) ? (int)(cell / N2) : 0;
    const int ip = (i + 1) % N, im = (i + N - 1) % N, jp = (j + 1) % N, jm = (j + N - 1) % N, kp = (k + 1) % N, km = (k + N - 1) % N;
    const int64_t off = (int64_t)s * NC, offx = (int64_t)(1 - s) * NC;
    rows[cnt++] = cell + off;
    rows[cnt++] = im + (int64_t)N * j + N2 * k + off;
    rows[cnt++] = ip + (int64_t)N * j + N2 * k + off;
    rows[cnt++] = i + (int64_t)N * jm + N2 * k + off;
    rows[cnt++] = i + (int64_t)N * jp + N2 * k + off;
    if (dim == 3) {
      rows[cnt++] = i + (int64_t)N * j + N2 * km + off;
      rows[cnt++] = i + (int64_t)N * j + N2 * kp + off;
    }
    rows[cnt++] = cell + offx;
  } else if (p->kind == B200_PROB_QUADRATIC) {
    rows[cnt++] = c;
  } else if (p->kind == B200_PROB_TRIDIAG_QUAD) {
    if (c > 0) rows[cnt++] = c - 1;
    rows[cnt++] = c;
    if (c + 1 < p->n) rows[cnt++] = c + 1;
  }
  std::sort(rows, rows + cnt);
  return (int)(std::unique(rows, rows + cnt) - rows);
}

__global__ void __launch_bounds__(ST) seed_kernel(int64_t n, const int32_t* __restrict__ colors, int32_t color, double* __restrict__ seed) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) seed[c] = (colors[c] == color) ? 1.0 : 0.0;
}
// decompression: for every column of this colour copy the compressed column's entries into nzval
__global__ void __launch_bounds__(ST) scatter_kernel(int64_t n, const int32_t* __restrict__ colors, int32_t color,
                                                      const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval,
                                                      const double* __restrict__ comp, double* __restrict__ nzval) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n || colors[c] != color) return;
  for (int64_t p = colptr[c]; p < colptr[c + 1]; ++p) nzval[p] = comp[rowval[p]];
}
// y = J x, row gather through the CSR view of the CSC pattern (deterministic, no atomics)
__global__ void __launch_bounds__(ST) spmv_rows_kernel(int64_t n, const int64_t* __restrict__ rowptr, const int64_t* __restrict__ csr_col,
                                                        const int64_t* __restrict__ csr_map, const double* __restrict__ nzval,
                                                        const double* __restrict__ x, double* __restrict__ y) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double s = 0.0;
  for (int64_t q = rowptr[r]; q < rowptr[r + 1]; ++q) s = fma(nzval[csr_map[q]], x[csr_col[q]], s);
  y[r] = s;
}
// y = J' x, column gather (CSC natural order)
__global__ void __launch_bounds__(ST) spmv_cols_kernel(int64_t n, const int64_t* __restrict__ colptr, const int64_t* __restrict__ rowval,
                                                        const double* __restrict__ nzval, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double s = 0.0;
  for (int64_t p = colptr[c]; p < colptr[c + 1]; ++p) s = fma(nzval[p], x[rowval[p]], s);
  y[c] = s;
}
}  // namespace

struct b200_sparse_jac {
  b200_ctx* ctx;
  b200_problem* prob;
  int64_t n, nnz;
  int64_t ncolors;
  int64_t *d_colptr, *d_rowval;           // 0-based
  int64_t *d_rowptr, *d_csr_col, *d_csr_map;
  int32_t* d_colors;                      // 1-based colours
  double *seed, *comp;
};

void b200i_sparse_jac_csr(b200_sparse_jac* sj, const int64_t** rowptr, const int64_t** csr_col, const int64_t** csr_map) {
  *rowptr = sj->d_rowptr; *csr_col = sj->d_csr_col; *csr_map = sj->d_csr_map;
}

extern "C" {
int32_t b200_pattern_nnz(b200_problem* p, int64_t* nnz) {
  B200_DEVICE_GUARD(p ? p->ctx : nullptr);
  if (p->proto_colptr) { *nnz = p->proto_colptr[p->n]; return B200_OK; }  // user jac_prototype
  B200_REQUIRE(p->ctx, p->kind != B200_PROB_CALLBACK, "pattern: callback problems must bring their own jac_prototype");
  int64_t rows[16], total = 0;
  for (int64_t c = 0; c < p->n; ++c) total += column_rows(p, c, rows);
  *nnz = total;
  return B200_OK;
}
int32_t b200_pattern(b200_problem* p, int32_t base, int64_t* colptr, int64_t* rowval) {
  B200_DEVICE_GUARD(p ? p->ctx : nullptr);
  if (p->proto_colptr) {
    for (int64_t c = 0; c <= p->n; ++c) colptr[c] = p->proto_colptr[c] + base;
    for (int64_t e = 0; e < p->proto_colptr[p->n]; ++e) rowval[e] = p->proto_rowval[e] + base;
    return B200_OK;
  }
  B200_REQUIRE(p->ctx, p->kind != B200_PROB_CALLBACK, "pattern: callback problems must bring their own jac_prototype");
  int64_t rows[16], pos = 0;
  for (int64_t c = 0; c < p->n; ++c) {
    colptr[c] = pos + base;
    const int m = column_rows(p, c, rows);
    for (int x = 0; x < m; ++x) rowval[pos++] = rows[x] + base;
  }
  colptr[p->n] = pos + base;
  return B200_OK;
}

// GreedyColoringAlgorithm(LargestFirst()) / natural order, column partition: partial distance-2 colouring of the
// bipartite row/column graph, smallest admissible colour, 1-based.  Sequential by construction (each vertex depends
// on all previously coloured neighbours), executed once at set-up on the host.
int32_t b200_coloring_column(int64_t n, const int64_t* colptr, const int64_t* rowval, int32_t base, int32_t order, int64_t* colors,
                             int64_t* ncolors) {
  if (n <= 0 || !colptr || !rowval || !colors || !ncolors) return B200_ERR_INVALID;
  const int64_t nnz = colptr[n] - base;
  std::vector<int64_t> rowptr(n + 1, 0), colidx(nnz > 0 ? nnz : 1);
  for (int64_t k = 0; k < nnz; ++k) {
    const int64_t r = rowval[k] - base;
    if (r < 0 || r >= n) return B200_ERR_INVALID;
    rowptr[r + 1]++;
  }
  for (int64_t r = 0; r < n; ++r) rowptr[r + 1] += rowptr[r];
  {
    std::vector<int64_t> fill(rowptr.begin(), rowptr.end() - 1);
    for (int64_t c = 0; c < n; ++c)
      for (int64_t k = colptr[c] - base; k < colptr[c + 1] - base; ++k) colidx[fill[rowval[k] - base]++] = c;
  }
  std::vector<int64_t> perm(n);
  for (int64_t c = 0; c < n; ++c) perm[c] = c;
  if (order == B200_ORDER_LARGEST_FIRST)  // decreasing column degree, stable (ties keep natural order)
    std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) { return (colptr[a + 1] - colptr[a]) > (colptr[b + 1] - colptr[b]); });
  std::fill(colors, colors + n, 0);
  std::vector<int64_t> forbidden(64, -1);
  int64_t maxcolor = 0;
  for (int64_t t = 0; t < n; ++t) {
    const int64_t v = perm[t];
    for (int64_t k = colptr[v] - base; k < colptr[v + 1] - base; ++k) {
      const int64_t r = rowval[k] - base;
      for (int64_t q = rowptr[r]; q < rowptr[r + 1]; ++q) {
        const int64_t x = colidx[q];
        if (colors[x] != 0) forbidden[colors[x]] = v;
      }
    }
    int64_t col = 1;
    while (col < (int64_t)forbidden.size() && forbidden[col] == v) ++col;
    if (col + 1 >= (int64_t)forbidden.size()) forbidden.resize(forbidden.size() * 2, -1);
    colors[v] = col;
    maxcolor = std::max(maxcolor, col);
  }
  *ncolors = maxcolor;
  return B200_OK;
}

int32_t b200_sparse_jac_destroy(b200_sparse_jac* sj) {
  B200_DEVICE_GUARD(sj ? sj->ctx : nullptr);
  if (!sj) return B200_OK;
  cudaStreamSynchronize(sj->ctx->stream);
  cudaFree(sj->d_colptr); cudaFree(sj->d_rowval); cudaFree(sj->d_rowptr); cudaFree(sj->d_csr_col); cudaFree(sj->d_csr_map);
  cudaFree(sj->d_colors); cudaFree(sj->seed); cudaFree(sj->comp);
  delete sj;
  return B200_OK;
}

int32_t b200_sparse_jac_create(b200_problem* prob, const int64_t* colptr, const int64_t* rowval, int32_t base, const int64_t* colors,
                               int64_t ncolors, b200_sparse_jac** out) {
  B200_DEVICE_GUARD(prob ? prob->ctx : nullptr);
  b200_ctx* ctx = prob->ctx;
  const int64_t n = prob->n;
  B200_REQUIRE(ctx, colptr && rowval && colors && ncolors > 0 && out, "sparse_jac_create: bad arguments");
  const int64_t nnz = colptr[n] - base;
  b200_sparse_jac* sj = new b200_sparse_jac();
  memset(sj, 0, sizeof(*sj));
  sj->ctx = ctx; sj->prob = prob; sj->n = n; sj->nnz = nnz; sj->ncolors = ncolors;
  std::vector<int64_t> cp(n + 1), rv(nnz), rowptr(n + 1, 0), ccol(nnz), cmap(nnz);
  std::vector<int32_t> col32(n);
  for (int64_t c = 0; c <= n; ++c) cp[c] = colptr[c] - base;
  for (int64_t k = 0; k < nnz; ++k) { rv[k] = rowval[k] - base; rowptr[rv[k] + 1]++; }
  for (int64_t r = 0; r < n; ++r) rowptr[r + 1] += rowptr[r];
  {
    std::vector<int64_t> fill(rowptr.begin(), rowptr.end() - 1);
    for (int64_t c = 0; c < n; ++c)
      for (int64_t k = cp[c]; k < cp[c + 1]; ++k) { const int64_t q = fill[rv[k]]++; ccol[q] = c; cmap[q] = k; }
  }
  for (int64_t c = 0; c < n; ++c) col32[c] = (int32_t)colors[c];
  auto up = [&](void** d, const void* h, size_t bytes) -> int32_t {
    CUDA_TRY(ctx, cudaMalloc(d, bytes ? bytes : 8));
    CUDA_TRY(ctx, cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return B200_OK;
  };
  int32_t s = up((void**)&sj->d_colptr, cp.data(), sizeof(int64_t) * (n + 1));
  if (s == B200_OK) s = up((void**)&sj->d_rowval, rv.data(), sizeof(int64_t) * nnz);
  if (s == B200_OK) s = up((void**)&sj->d_rowptr, rowptr.data(), sizeof(int64_t) * (n + 1));
  if (s == B200_OK) s = up((void**)&sj->d_csr_col, ccol.data(), sizeof(int64_t) * nnz);
  if (s == B200_OK) s = up((void**)&sj->d_csr_map, cmap.data(), sizeof(int64_t) * nnz);
  if (s == B200_OK) s = up((void**)&sj->d_colors, col32.data(), sizeof(int32_t) * n);
  if (s == B200_OK && cudaMalloc(&sj->seed, sizeof(double) * n) != cudaSuccess) s = B200_ERR_NOMEM;
  if (s == B200_OK && cudaMalloc(&sj->comp, sizeof(double) * n) != cudaSuccess) s = B200_ERR_NOMEM;
  if (s == B200_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) s = B200_ERR_CUDA;  // host vectors die at scope exit
  if (s != B200_OK) { cudaGetLastError(); b200_sparse_jac_destroy(sj); return s; }
  *out = sj;
  return B200_OK;
}

// DI.jacobian! with AutoSparse(AutoForwardDiff): one exact JVP per colour + decompression (jacobian.jl:244-247)
int32_t b200_sparse_jac_fill(b200_sparse_jac* sj, const double* u, double* nzval) {
  B200_DEVICE_GUARD(sj ? sj->ctx : nullptr);
  b200_ctx* ctx = sj->ctx;
  if (sj->prob->jac_nzval_cb) B200_TRY(b200i_sync_for_callback(ctx));
  if (sj->prob->jac_nzval_cb)  // jac!(J::SparseMatrixCSC, u, p) writes nzval directly
    return sj->prob->jac_nzval_cb(sj->prob->user, u, nzval) == 0 ? B200_OK : ctx->fail(B200_ERR_CALLBACK, "jac! callback failed", __FILE__, __LINE__);
  const int grid = (int)((sj->n + ST - 1) / ST);
  for (int32_t color = 1; color <= (int32_t)sj->ncolors; ++color) {
    LAUNCH(ctx, seed_kernel, grid, ST, 0, sj->n, (const int32_t*)sj->d_colors, color, sj->seed);
    B200_TRY(b200_jvp(sj->prob, u, sj->seed, sj->comp));
    LAUNCH(ctx, scatter_kernel, grid, ST, 0, sj->n, (const int32_t*)sj->d_colors, color, (const int64_t*)sj->d_colptr,
           (const int64_t*)sj->d_rowval, (const double*)sj->comp, nzval);
  }
  CHECK_LAUNCH(ctx);
  return B200_OK;
}

int32_t b200_spmv(b200_sparse_jac* sj, const double* nzval, const double* x, double* y) {
  B200_DEVICE_GUARD(sj ? sj->ctx : nullptr);
  b200_ctx* ctx = sj->ctx;
  LAUNCH(ctx, spmv_rows_kernel, (int)((sj->n + ST - 1) / ST), ST, 0, sj->n, (const int64_t*)sj->d_rowptr, (const int64_t*)sj->d_csr_col,
         (const int64_t*)sj->d_csr_map, nzval, x, y);
  CHECK_LAUNCH(ctx);
  return B200_OK;
}
int32_t b200_spmv_t(b200_sparse_jac* sj, const double* nzval, const double* x, double* y) {
  B200_DEVICE_GUARD(sj ? sj->ctx : nullptr);
  b200_ctx* ctx = sj->ctx;
  LAUNCH(ctx, spmv_cols_kernel, (int)((sj->n + ST - 1) / ST), ST, 0, sj->n, (const int64_t*)sj->d_colptr, (const int64_t*)sj->d_rowval, nzval,
         x, y);
  CHECK_LAUNCH(ctx);
  return B200_OK;
}
int32_t b200_sparse_jac_linop(b200_sparse_jac* sj, const double* nzval, b200_linop** out) {
  B200_DEVICE_GUARD(sj ? sj->ctx : nullptr);
  b200_linop* op = new b200_linop();
  memset(op, 0, sizeof(*op));
  op->ctx = sj->ctx; op->kind = LINOP_SPARSE_JAC; op->n = sj->n; op->sj = sj; op->nzval = nzval;
  *out = op;
  return B200_OK;
}
}  // extern "C"
