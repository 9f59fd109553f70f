// ctx.cu — context, memory, and the BLAS-1 style vector operations the reference touches on a device array
// (similar/copyto!/fill!, broadcast axpy, maximum(abs,x), norm(x,2), dot, ==).
#include "common.cuh"
#include <string.h>

int32_t b200_ctx::fail(int32_t code, const char* what, const char* file, int line) {
  char buf[768];
  snprintf(buf, sizeof(buf), "[b200newton] %s (%s:%d)", what, file, line);
  last_error = buf;
  return code;
}

void b200_ctx::prof_begin(int kid, double bytes) {
  ProfPending p;
  p.kid = kid;
  p.bytes = bytes;
  for (cudaEvent_t* e : {&p.a, &p.b}) {
    if (!prof_free.empty()) { *e = prof_free.back(); prof_free.pop_back(); }
    else cudaEventCreate(e);
  }
  cudaEventRecord(p.a, stream);
  prof_pending.push_back(p);
}
void b200_ctx::prof_end() {
  cudaEventRecord(prof_pending.back().b, stream);
  if (prof_pending.size() >= 32768) prof_collect();
}
void b200_ctx::prof_collect() {
  if (prof_pending.empty()) return;
  cudaStreamSynchronize(stream);
  if (aux_stream) cudaStreamSynchronize(aux_stream);
  for (ProfPending& p : prof_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      prof[p.kid].ms += ms;
      prof[p.kid].bytes += p.bytes;
      prof[p.kid].launches += 1;
    }
    prof_free.push_back(p.a);
    prof_free.push_back(p.b);
  }
  prof_pending.clear();
}

extern "C" {

int32_t b200_version(void) { return B200_VERSION; }

int32_t b200_ctx_profile_enable(b200_ctx* ctx, int32_t on) {
  B200_DEVICE_GUARD(ctx);
  ctx->prof_collect();
  ctx->prof_on = on != 0;
  return B200_OK;
}
int32_t b200_ctx_profile_reset(b200_ctx* ctx) {
  B200_DEVICE_GUARD(ctx);
  ctx->prof_collect();
  for (auto& s : ctx->prof) s = b200_ctx::ProfSlot();
  return B200_OK;
}
int32_t b200_ctx_profile_get(b200_ctx* ctx, int32_t kid, double* ms, double* bytes, int64_t* launches) {
  B200_DEVICE_GUARD(ctx);
  if (kid < 0 || kid >= B200_KID_COUNT) return ctx->fail(B200_ERR_INVALID, "bad kernel id", __FILE__, __LINE__);
  ctx->prof_collect();
  if (ms) *ms = ctx->prof[kid].ms;
  if (bytes) *bytes = ctx->prof[kid].bytes;
  if (launches) *launches = ctx->prof[kid].launches;
  return B200_OK;
}

int32_t b200_device_count(int32_t* count) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  *count = (e == cudaSuccess) ? c : 0;
  return (e == cudaSuccess) ? B200_OK : B200_ERR_NO_DEVICE;
}

int32_t b200_ctx_create(int32_t device, void* stream, b200_ctx** out) {
  if (!out) return B200_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return B200_ERR_NO_DEVICE;  // no CPU fallback, ever
  if (device < 0 || device >= count) return B200_ERR_INVALID;
  b200_ctx* ctx = new b200_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return B200_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return B200_ERR_CUDA; }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  ctx->l2_bytes = prop.l2CacheSize;
  if (stream) {
    ctx->stream = reinterpret_cast<cudaStream_t>(stream);
    ctx->own_stream = false;
  } else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return B200_ERR_CUDA; }
    ctx->own_stream = true;
  }
  bool ok = cudaMalloc(&ctx->d_partials, sizeof(double) * B200_RED_MAX_BLOCKS * 4) == cudaSuccess &&
            cudaMalloc(&ctx->d_scalars, sizeof(double) * 64) == cudaSuccess &&
            cudaMallocHost(&ctx->h_scalars, sizeof(double) * 64) == cudaSuccess;
  if (!ok) { b200_ctx_destroy(ctx); return B200_ERR_NOMEM; }
  cudaMemsetAsync(ctx->d_scalars, 0, sizeof(double) * 64, ctx->stream);
  *out = ctx;
  return B200_OK;
}

int32_t b200_ctx_destroy(b200_ctx* ctx) {
  B200_DEVICE_GUARD(ctx);
  if (!ctx) return B200_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->d_partials) cudaFree(ctx->d_partials);
  if (ctx->d_lu_xchg) cudaFree(ctx->d_lu_xchg);
  if (ctx->d_scalars) cudaFree(ctx->d_scalars);
  if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
  if (ctx->l2_flush) cudaFree(ctx->l2_flush);
  if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); cudaEventDestroy(ctx->ev_a); cudaEventDestroy(ctx->ev_b); }
  ctx->prof_collect();
  for (cudaEvent_t e : ctx->prof_free) cudaEventDestroy(e);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return B200_OK;
}

int32_t b200_ctx_sync(b200_ctx* ctx) {
  B200_DEVICE_GUARD(ctx);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
void* b200_ctx_stream(b200_ctx* ctx) { return reinterpret_cast<void*>(ctx->stream); }
const char* b200_last_error(b200_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }
int32_t b200_ctx_kernel_launches(b200_ctx* ctx, int64_t* count) { *count = ctx->launches; return B200_OK; }
int32_t b200_ctx_sm_count(b200_ctx* ctx, int32_t* count) { *count = ctx->sm_count; return B200_OK; }

int32_t b200_malloc(b200_ctx* ctx, size_t bytes, void** dptr) {
  B200_DEVICE_GUARD(ctx);
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 8);
  if (e != cudaSuccess) { cudaGetLastError(); *dptr = nullptr; return ctx->fail(B200_ERR_NOMEM, "cudaMalloc failed", __FILE__, __LINE__); }
  return B200_OK;
}
int32_t b200_free(b200_ctx* ctx, void* dptr) {
  B200_DEVICE_GUARD(ctx);
  if (!dptr) return B200_OK;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(ctx, cudaFree(dptr));
  return B200_OK;
}
int32_t b200_host_alloc(b200_ctx* ctx, size_t bytes, void** hptr) {
  B200_DEVICE_GUARD(ctx);
  CUDA_TRY(ctx, cudaMallocHost(hptr, bytes ? bytes : 8));
  return B200_OK;
}
int32_t b200_host_free(b200_ctx* ctx, void* hptr) {
  B200_DEVICE_GUARD(ctx);
  if (hptr) CUDA_TRY(ctx, cudaFreeHost(hptr));
  return B200_OK;
}
int32_t b200_memcpy_h2d(b200_ctx* ctx, void* dst, const void* src, size_t bytes) {
  B200_DEVICE_GUARD(ctx);
  CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
int32_t b200_memcpy_d2h(b200_ctx* ctx, void* dst, const void* src, size_t bytes) {
  B200_DEVICE_GUARD(ctx);
  CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}
int32_t b200_memcpy_d2d(b200_ctx* ctx, void* dst, const void* src, size_t bytes) {
  B200_DEVICE_GUARD(ctx);
  CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return B200_OK;
}
int32_t b200_memset(b200_ctx* ctx, void* dst, int32_t byte, size_t bytes) {
  B200_DEVICE_GUARD(ctx);
  CUDA_TRY(ctx, cudaMemsetAsync(dst, byte, bytes, ctx->stream));
  return B200_OK;
}
int32_t b200_flush_l2(b200_ctx* ctx) {
  B200_DEVICE_GUARD(ctx);
  if (!ctx->l2_flush) {
    ctx->l2_flush_bytes = ctx->l2_bytes ? 2 * ctx->l2_bytes : (size_t)256 << 20;
    CUDA_TRY(ctx, cudaMalloc(&ctx->l2_flush, ctx->l2_flush_bytes));
  }
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->l2_flush, 0, ctx->l2_flush_bytes, ctx->stream));
  return B200_OK;
}
}  // extern "C"

// ---------------------------------------------------------------- elementwise kernels
namespace {
constexpr int EW_THREADS = 256;

inline int ew_grid(const b200_ctx* ctx, int64_t n, int per_thread = 4) {
  int64_t b = (n + (int64_t)EW_THREADS * per_thread - 1) / ((int64_t)EW_THREADS * per_thread);
  int64_t cap = (int64_t)ctx->sm_count * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

enum { EW_FILL = 0, EW_COPY = 1, EW_SCAL = 2, EW_AXPY = 3, EW_AXPBY = 4, EW_MUL = 5 };

template <int OP>
__global__ void __launch_bounds__(EW_THREADS) ew_kernel(int64_t n, double a, double b, const double* __restrict__ x,
                                                         const double* __restrict__ y, double* __restrict__ z) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (OP == EW_FILL) z[i] = a;
    else if (OP == EW_COPY) z[i] = x[i];
    else if (OP == EW_SCAL) z[i] = a * z[i];
    else if (OP == EW_AXPY) z[i] = fma(a, x[i], z[i]);
    else if (OP == EW_AXPBY) z[i] = a * x[i] + b * z[i];
    else if (OP == EW_MUL) z[i] = x[i] * y[i];
  }
}

// y += a x, and sum(x^2 * a^2) accumulated into *sumsq (the Newton update u += du fused with ||du||^2;
// replaces `@bb axpy!(1, du, u)` solve.jl:438 plus the stall-test norm termination_conditions.jl:306-313)
__global__ void __launch_bounds__(EW_THREADS) axpy_norm_kernel(int64_t n, double a, const double* __restrict__ x,
                                                                double* __restrict__ y, double* __restrict__ sumsq) {
  __shared__ double red[32];
  double s = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double d = a * x[i];
    y[i] += d;
    s = fma(d, d, s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(sumsq, s);
}

template <int MODE>
__device__ __forceinline__ double red_identity() {
  if (MODE == RED_MIN) return __longlong_as_double(0x7ff0000000000000LL);
  if (MODE == RED_MAX) return __longlong_as_double(0xfff0000000000000LL);
  return 0.0;
}
template <int MODE>
__device__ __forceinline__ double red_combine(double a, double b) {
  if (MODE == RED_MAXABS || MODE == RED_MAX || MODE == RED_MAXABS2) return fmax(a, b);
  if (MODE == RED_MIN) return fmin(a, b);
  return a + b;
}
template <int MODE>
__device__ __forceinline__ double red_block(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = red_combine<MODE>(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  v = (threadIdx.x < nw) ? red[threadIdx.x] : red_identity<MODE>();
  if (wid == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_combine<MODE>(v, __shfl_xor_sync(0xffffffffu, v, o));
  }
  return v;
}

template <int MODE>
__global__ void __launch_bounds__(EW_THREADS) reduce_stage1(int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                                                             double* __restrict__ partials, double a) {
  __shared__ double red[32];
  double acc = red_identity<MODE>();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double v;
    if (MODE == RED_DOT) v = x[i] * y[i];
    else if (MODE == RED_SUMSQ) v = x[i] * x[i];
    else if (MODE == RED_MAXABS) v = abs_nf(x[i]);
    else if (MODE == RED_DIFFSQ) { double d = x[i] - y[i]; v = d * d; }
    else if (MODE == RED_NEQ) v = (x[i] != y[i]) ? 1.0 : 0.0;
    else if (MODE == RED_SUMSQ2) { double d = x[i] + y[i]; v = d * d; }
    else if (MODE == RED_MAXABS2) v = abs_nf(x[i] + y[i]);
    else if (MODE == RED_RELVIOL) v = (fabs(x[i]) <= a * fabs(x[i] + y[i])) ? 0.0 : 1.0;  // NaN counts as a violation
    else if (MODE == RED_COUNT_LE) v = (fabs(x[i]) <= a) ? 1.0 : 0.0;
    else if (MODE == RED_COUNT_DIFF_LE) v = (fabs(x[i] - y[i]) <= a) ? 1.0 : 0.0;
    else v = x[i];
    acc = red_combine<MODE>(acc, v);
  }
  acc = red_block<MODE>(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
template <int MODE>
__global__ void __launch_bounds__(EW_THREADS) reduce_stage2(int nb, const double* __restrict__ partials, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = red_identity<MODE>();
  for (int i = threadIdx.x; i < nb; i += blockDim.x) acc = red_combine<MODE>(acc, partials[i]);  // fixed order: deterministic
  acc = red_block<MODE>(acc, red);
  if (threadIdx.x == 0) *out = acc;
}

template <int MODE>
int32_t reduce_dev(b200_ctx* ctx, int64_t n, const double* x, const double* y, double* d_out, double* partials, double a = 0.0) {
  int64_t nb64 = (n + (int64_t)EW_THREADS * 8 - 1) / ((int64_t)EW_THREADS * 8);
  int nb = (int)(nb64 < 1 ? 1 : (nb64 > B200_RED_MAX_BLOCKS ? B200_RED_MAX_BLOCKS : nb64));
  LAUNCH(ctx, (reduce_stage1<MODE>), nb, EW_THREADS, 0, n, x, y, partials, a);
  LAUNCH(ctx, (reduce_stage2<MODE>), 1, EW_THREADS, 0, nb, partials, d_out);
  CHECK_LAUNCH(ctx);
  return B200_OK;
}
}  // namespace

int32_t b200i_reduce_sum_dev(b200_ctx* ctx, int64_t n, const double* x, const double* y, int mode, double* d_out, double a) {
  switch (mode) {
    case RED_SUMSQ2: return reduce_dev<RED_SUMSQ2>(ctx, n, x, y, d_out, ctx->d_partials);
    case RED_MAXABS2: return reduce_dev<RED_MAXABS2>(ctx, n, x, y, d_out, ctx->d_partials);
    case RED_RELVIOL: return reduce_dev<RED_RELVIOL>(ctx, n, x, y, d_out, ctx->d_partials, a);
    case RED_COUNT_LE: return reduce_dev<RED_COUNT_LE>(ctx, n, x, y, d_out, ctx->d_partials, a);
    case RED_COUNT_DIFF_LE: return reduce_dev<RED_COUNT_DIFF_LE>(ctx, n, x, y, d_out, ctx->d_partials, a);
    case RED_DOT: return reduce_dev<RED_DOT>(ctx, n, x, y, d_out, ctx->d_partials);
    case RED_SUMSQ: return reduce_dev<RED_SUMSQ>(ctx, n, x, y, d_out, ctx->d_partials);
    case RED_MAXABS: return reduce_dev<RED_MAXABS>(ctx, n, x, y, d_out, ctx->d_partials);
    case RED_DIFFSQ: return reduce_dev<RED_DIFFSQ>(ctx, n, x, y, d_out, ctx->d_partials);
    case RED_MIN: return reduce_dev<RED_MIN>(ctx, n, x, y, d_out, ctx->d_partials);
    case RED_MAX: return reduce_dev<RED_MAX>(ctx, n, x, y, d_out, ctx->d_partials + B200_RED_MAX_BLOCKS);
    case RED_NEQ: return reduce_dev<RED_NEQ>(ctx, n, x, y, d_out, ctx->d_partials);
  }
  return ctx->fail(B200_ERR_INVALID, "bad reduction mode", __FILE__, __LINE__);
}

int32_t b200i_fetch_scalars(b200_ctx* ctx, int count) {
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return B200_OK;
}

int32_t b200i_axpy_norm(b200_ctx* ctx, int64_t n, double a, const double* x, double* y, double* d_sumsq) {
  LAUNCH(ctx, axpy_norm_kernel, ew_grid(ctx, n), EW_THREADS, 0, n, a, x, y, d_sumsq);
  CHECK_LAUNCH(ctx);
  return B200_OK;
}

extern "C" {
int32_t b200_fill(b200_ctx* ctx, int64_t n, double a, double* x) {
  B200_DEVICE_GUARD(ctx);
  LAUNCH(ctx, (ew_kernel<EW_FILL>), ew_grid(ctx, n), EW_THREADS, 0, n, a, 0.0, nullptr, nullptr, x);
  CHECK_LAUNCH(ctx);
  return B200_OK;
}
int32_t b200_copy(b200_ctx* ctx, int64_t n, const double* x, double* y) {
  B200_DEVICE_GUARD(ctx);
  LAUNCH(ctx, (ew_kernel<EW_COPY>), ew_grid(ctx, n), EW_THREADS, 0, n, 0.0, 0.0, x, nullptr, y);
  CHECK_LAUNCH(ctx);
  return B200_OK;
}
int32_t b200_scal(b200_ctx* ctx, int64_t n, double a, double* x) {
  B200_DEVICE_GUARD(ctx);
  LAUNCH(ctx, (ew_kernel<EW_SCAL>), ew_grid(ctx, n), EW_THREADS, 0, n, a, 0.0, nullptr, nullptr, x);
  CHECK_LAUNCH(ctx);
  return B200_OK;
}
int32_t b200_axpy(b200_ctx* ctx, int64_t n, double a, const double* x, double* y) {
  B200_DEVICE_GUARD(ctx);
  LAUNCH(ctx, (ew_kernel<EW_AXPY>), ew_grid(ctx, n), EW_THREADS, 0, n, a, 0.0, x, nullptr, y);
  CHECK_LAUNCH(ctx);
  return B200_OK;
}
int32_t b200_axpby(b200_ctx* ctx, int64_t n, double a, const double* x, double b, double* y) {
  B200_DEVICE_GUARD(ctx);
  LAUNCH(ctx, (ew_kernel<EW_AXPBY>), ew_grid(ctx, n), EW_THREADS, 0, n, a, b, x, nullptr, y);
  CHECK_LAUNCH(ctx);
  return B200_OK;
}
int32_t b200_mul(b200_ctx* ctx, int64_t n, const double* x, const double* y, double* z) {
  B200_DEVICE_GUARD(ctx);
  LAUNCH(ctx, (ew_kernel<EW_MUL>), ew_grid(ctx, n), EW_THREADS, 0, n, 0.0, 0.0, x, y, z);
  CHECK_LAUNCH(ctx);
  return B200_OK;
}
int32_t b200_dot(b200_ctx* ctx, int64_t n, const double* x, const double* y, double* out_host) {
  B200_DEVICE_GUARD(ctx);
  B200_TRY(b200i_reduce_sum_dev(ctx, n, x, y, RED_DOT, ctx->d_scalars));
  B200_TRY(b200i_fetch_scalars(ctx, 1));
  *out_host = ctx->h_scalars[0];
  return B200_OK;
}
int32_t b200_nrm2(b200_ctx* ctx, int64_t n, const double* x, double* out_host) {
  B200_DEVICE_GUARD(ctx);
  B200_TRY(b200i_reduce_sum_dev(ctx, n, x, nullptr, RED_SUMSQ, ctx->d_scalars));
  B200_TRY(b200i_fetch_scalars(ctx, 1));
  *out_host = sqrt(ctx->h_scalars[0]);
  return B200_OK;
}
int32_t b200_norminf(b200_ctx* ctx, int64_t n, const double* x, double* out_host) {
  B200_DEVICE_GUARD(ctx);
  B200_TRY(b200i_reduce_sum_dev(ctx, n, x, nullptr, RED_MAXABS, ctx->d_scalars));
  B200_TRY(b200i_fetch_scalars(ctx, 1));
  *out_host = ctx->h_scalars[0];
  return B200_OK;
}
int32_t b200_diffnrm2(b200_ctx* ctx, int64_t n, const double* x, const double* y, double* out_host) {
  B200_DEVICE_GUARD(ctx);
  B200_TRY(b200i_reduce_sum_dev(ctx, n, x, y, RED_DIFFSQ, ctx->d_scalars));
  B200_TRY(b200i_fetch_scalars(ctx, 1));
  *out_host = sqrt(ctx->h_scalars[0]);
  return B200_OK;
}
int32_t b200_extrema(b200_ctx* ctx, int64_t n, const double* x, double* min_host, double* max_host) {
  B200_DEVICE_GUARD(ctx);
  B200_TRY(b200i_reduce_sum_dev(ctx, n, x, nullptr, RED_MIN, ctx->d_scalars));
  B200_TRY(b200i_reduce_sum_dev(ctx, n, x, nullptr, RED_MAX, ctx->d_scalars + 1));
  B200_TRY(b200i_fetch_scalars(ctx, 2));
  *min_host = ctx->h_scalars[0];
  *max_host = ctx->h_scalars[1];
  return B200_OK;
}
int32_t b200_equal(b200_ctx* ctx, int64_t n, const double* x, const double* y, int32_t* equal_host) {
  B200_DEVICE_GUARD(ctx);
  B200_TRY(b200i_reduce_sum_dev(ctx, n, x, y, RED_NEQ, ctx->d_scalars));
  B200_TRY(b200i_fetch_scalars(ctx, 1));
  *equal_host = (ctx->h_scalars[0] == 0.0) ? 1 : 0;
  return B200_OK;
}
}  // extern "C"
