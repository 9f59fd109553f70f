// Micro-benchmark (exploration only): latency of an all-CTA exchange of one double through L2, several protocols.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
constexpr int MAXG = 160;
__device__ __forceinline__ double warp_sum(double v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }

// mode 0: pull LL (padded slots, every CTA reads all), 1: push LL (private dense receive buffers), 2: atomic counter barrier + dense gather,
// 3: cg grid.sync + dense gather, 4: red.add.f64 accumulator + counter (non-deterministic)
__global__ void bench(int mode, int iters, unsigned long long* slots, unsigned* counter, double* accum, double* partial, long long* out) {
  cg::grid_group grid = cg::this_grid();
  const int G = gridDim.x, b = blockIdx.x, lane = threadIdx.x & 31;
  __shared__ double hs;
  double acc = 0.0;
  long long t0 = clock64();
  for (int t = 0; t < iters; ++t) {
    const double d = 1.0 + b * 1e-3 + t * 1e-6;
    const unsigned epoch = t + 1;
    if (mode == 0 || mode == 1) {
      if (threadIdx.x < 32) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(d);
        const unsigned long long w0 = (bits & 0xffffffffull) | ((unsigned long long)epoch << 32), w1 = (bits >> 32) | ((unsigned long long)epoch << 32);
        unsigned long long* base = slots + (size_t)(t & 1) * 2 * MAXG * MAXG;
        if (mode == 0) { if (lane == 0) asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(base + 32 * (size_t)b), "l"(w0), "l"(w1) : "memory"); }
        else for (int q = 0; q < 5; ++q) { int dest = lane + 32 * q; if (dest < G) asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(base + 2 * ((size_t)dest * MAXG + b)), "l"(w0), "l"(w1) : "memory"); }
        unsigned long long a0[5], a1[5];
        bool ok;
        do {
          ok = true;
          for (int q = 0; q < 5; ++q) { int s = lane + 32 * q; if (s < G) { const unsigned long long* p = (mode == 0) ? base + 32 * (size_t)s : base + 2 * ((size_t)b * MAXG + s);
              asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a0[q]), "=l"(a1[q]) : "l"(p) : "memory"); } }
          for (int q = 0; q < 5; ++q) { int s = lane + 32 * q; if (s < G) ok = ok && (unsigned)(a0[q] >> 32) == epoch && (unsigned)(a1[q] >> 32) == epoch; }
        } while (!__all_sync(0xffffffffu, ok));
        double s = 0.0;
        for (int q = 0; q < 5; ++q) { int sl = lane + 32 * q; if (sl < G) s += __longlong_as_double((long long)((a0[q] & 0xffffffffull) | (a1[q] << 32))); }
        s = warp_sum(s);
        if (lane == 0) hs = s;
      }
      __syncthreads();
    } else if (mode == 2 || mode == 3) {
      double* part = partial + (size_t)(t & 1) * MAXG;
      if (threadIdx.x == 0) part[b] = d;
      if (mode == 3) grid.sync();
      else {
        __syncthreads();
        if (threadIdx.x == 0) { __threadfence(); atomicAdd(counter, 1u); unsigned v; do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < (unsigned)G * (t + 1)); }
        __syncthreads();
      }
      if (threadIdx.x < 32) { double s = 0.0; for (int q = lane; q < G; q += 32) s += __ldcg(part + q); s = warp_sum(s); if (lane == 0) hs = s; }
      __syncthreads();
    } else {
      double* a = accum + (t & 3);
      if (threadIdx.x == 0) {
        atomicAdd(a, d); __threadfence(); atomicAdd(counter, 1u);
        unsigned v; do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < (unsigned)G * (t + 1));
        hs = __ldcg(a);
        if (b == 0) accum[(t + 2) & 3] = 0.0;  // recycle a slot two steps ahead
      }
      __syncthreads();
    }
    acc += hs;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[b] = t1 - t0; if (b == 0) out[MAXG] = (long long)acc; }
}
int main() {
  int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* slots; unsigned* counter; double *accum, *partial; long long* out;
  cudaMalloc(&slots, sizeof(unsigned long long) * 4 * MAXG * MAXG); cudaMalloc(&counter, 4); cudaMalloc(&accum, 64); cudaMalloc(&partial, sizeof(double) * 2 * MAXG); cudaMalloc(&out, 8 * (MAXG + 1));
  const char* names[] = {"pull-LL padded", "push-LL private", "atomic counter + gather", "cg grid.sync + gather", "red.add accumulator"};
  for (int threads : {32, 512})
    for (int mode = 0; mode < 5; ++mode) {
      cudaMemset(slots, 0, sizeof(unsigned long long) * 4 * MAXG * MAXG); cudaMemset(counter, 0, 4); cudaMemset(accum, 0, 64);
      int iters = 20000; int G = sm;
      void* args[] = {&mode, &iters, &slots, &counter, &accum, &partial, &out};
      cudaError_t e = cudaLaunchCooperativeKernel((const void*)bench, dim3(G), dim3(threads), args, 0, 0);
      cudaDeviceSynchronize();
      long long h[MAXG + 1]; cudaMemcpy(h, out, 8 * (MAXG + 1), cudaMemcpyDeviceToHost);
      printf("threads=%3d %-26s: %8.0f cycles/exchange (err=%s)\n", threads, names[mode], (double)h[0] / iters, cudaGetErrorString(e == cudaSuccess ? cudaGetLastError() : e));
    }
  return 0;
}
