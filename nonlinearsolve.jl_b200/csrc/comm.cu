// comm.cu — the collective step of the ensemble path (SURVEY.md §8b ABI list: b200_nccl_init / b200_ens_allgather; §8e).
//
// The ensemble shards by contiguous trajectory blocks with NO collective on the data path; the only exchange is after the
// solve: one all-gather of the solutions and one all-reduce of the status counters.  These entry points put that step
// behind the C ABI so that a host without torch.distributed (the Julia glue, tests/abi_c) can run BASELINE config 5 on
// 8 GPUs: NCCL over NVLink / NVSwitch, enqueued on the context's stream, so it orders after the solve kernel without a
// host synchronisation.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the system copy or the one a host process already loaded — e.g.
// torch's bundled build): libb200newton.so itself links nothing but cudart, and a process that never calls these entry
// points never loads NCCL.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>

struct b200_comm {
  b200_ctx* ctx;
  ncclComm_t comm;
  int32_t nranks, rank;
  double* d_stats;  // 8 doubles of device scratch for the status all-reduce
  double* h_stats;  // pinned mirror
};

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void nccl_load() {
  const char* names[] = {getenv("B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm || !*nm) continue;
    g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) { g_nccl.why = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return; }
#define NCCL_SYM(field, name)                                                                            \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.handle, name));                   \
  if (!g_nccl.field) { g_nccl.why = std::string("libnccl lacks ") + name; g_nccl.handle = nullptr; return; }
  NCCL_SYM(GetVersion, "ncclGetVersion")
  NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  NCCL_SYM(CommInitRank, "ncclCommInitRank")
  NCCL_SYM(CommInitAll, "ncclCommInitAll")
  NCCL_SYM(CommDestroy, "ncclCommDestroy")
  NCCL_SYM(AllGather, "ncclAllGather")
  NCCL_SYM(AllReduce, "ncclAllReduce")
  NCCL_SYM(GroupStart, "ncclGroupStart")
  NCCL_SYM(GroupEnd, "ncclGroupEnd")
  NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
}
bool nccl_ready() {
  std::call_once(g_nccl_once, nccl_load);
  return g_nccl.handle != nullptr;
}
#define NCCL_TRY(ctx, expr)                                                                              \
  do {                                                                                                   \
    ncclResult_t r__ = (expr);                                                                           \
    if (r__ != ncclSuccess) {                                                                            \
      char buf__[384];                                                                                   \
      snprintf(buf__, sizeof(buf__), "%s -> %s", #expr, g_nccl.GetErrorString(r__));                     \
      return (ctx)->fail(B200_ERR_CUDA, buf__, __FILE__, __LINE__);                                      \
    }                                                                                                    \
  } while (0)

int32_t comm_finish(b200_ctx* ctx, ncclComm_t c, int32_t nranks, int32_t rank, b200_comm** out) {
  b200_comm* cm = new b200_comm();
  cm->ctx = ctx; cm->comm = c; cm->nranks = nranks; cm->rank = rank; cm->d_stats = nullptr; cm->h_stats = nullptr;
  if (cudaMalloc(&cm->d_stats, 16 * sizeof(double)) != cudaSuccess || cudaMallocHost(&cm->h_stats, 16 * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    g_nccl.CommDestroy(c);
    delete cm;
    return ctx->fail(B200_ERR_NOMEM, "nccl_init: scratch allocation failed", __FILE__, __LINE__);
  }
  *out = cm;
  return B200_OK;
}
}  // namespace

extern "C" {
int32_t b200_nccl_version(int32_t* version) {
  if (!nccl_ready()) return B200_ERR_UNSUPPORTED;
  int v = 0;
  if (g_nccl.GetVersion(&v) != ncclSuccess) return B200_ERR_CUDA;
  *version = v;
  return B200_OK;
}

int32_t b200_nccl_unique_id(void* id128_host) {
  if (!id128_host) return B200_ERR_INVALID;
  if (!nccl_ready()) return B200_ERR_UNSUPPORTED;
  static_assert(sizeof(ncclUniqueId) == B200_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return B200_ERR_CUDA;
  memcpy(id128_host, &id, sizeof(id));
  return B200_OK;
}

int32_t b200_nccl_init(b200_ctx* ctx, int32_t nranks, int32_t rank, const void* unique_id128_host, b200_comm** comm) {
  B200_DEVICE_GUARD(ctx);
  B200_REQUIRE(ctx, nranks >= 1 && rank >= 0 && rank < nranks && unique_id128_host && comm, "nccl_init: bad arguments");
  if (!nccl_ready()) return ctx->fail(B200_ERR_UNSUPPORTED, g_nccl.why.c_str(), __FILE__, __LINE__);
  ncclUniqueId id;
  memcpy(&id, unique_id128_host, sizeof(id));
  ncclComm_t c;
  NCCL_TRY(ctx, g_nccl.CommInitRank(&c, nranks, id, rank));
  return comm_finish(ctx, c, nranks, rank, comm);
}

int32_t b200_nccl_init_all(b200_ctx* const* ctxs, int32_t ndev, b200_comm** comms) {
  if (!ctxs || ndev < 1 || !comms || !ctxs[0]) return B200_ERR_INVALID;
  b200_ctx* ctx0 = ctxs[0];
  if (!nccl_ready()) return ctx0->fail(B200_ERR_UNSUPPORTED, g_nccl.why.c_str(), __FILE__, __LINE__);
  std::vector<int> devs(ndev);
  std::vector<ncclComm_t> cs(ndev);
  for (int i = 0; i < ndev; ++i) devs[i] = ctxs[i]->device;
  NCCL_TRY(ctx0, g_nccl.CommInitAll(cs.data(), ndev, devs.data()));
  for (int i = 0; i < ndev; ++i) {
    B200_DEVICE_GUARD(ctxs[i]);
    B200_TRY(comm_finish(ctxs[i], cs[i], ndev, i, &comms[i]));
  }
  return B200_OK;
}

int32_t b200_nccl_destroy(b200_comm* cm) {
  if (!cm) return B200_OK;
  B200_DEVICE_GUARD(cm->ctx);
  cudaStreamSynchronize(cm->ctx->stream);
  g_nccl.CommDestroy(cm->comm);
  cudaFree(cm->d_stats);
  cudaFreeHost(cm->h_stats);
  delete cm;
  return B200_OK;
}

int32_t b200_nccl_group_start(void) { return nccl_ready() && g_nccl.GroupStart() == ncclSuccess ? B200_OK : B200_ERR_CUDA; }
int32_t b200_nccl_group_end(void) { return nccl_ready() && g_nccl.GroupEnd() == ncclSuccess ? B200_OK : B200_ERR_CUDA; }

// u_all[r * count_local + i] = u_local of rank r: the ordered collection of the trajectory blocks (EnsembleSolution.u)
int32_t b200_ens_allgather(b200_comm* cm, const double* u_local_dev, int64_t count_local, double* u_all_dev) {
  B200_DEVICE_GUARD(cm ? cm->ctx : nullptr);
  if (!cm) return B200_ERR_INVALID;
  b200_ctx* ctx = cm->ctx;
  B200_REQUIRE(ctx, u_local_dev && u_all_dev && count_local > 0, "ens_allgather: bad arguments");
  NCCL_TRY(ctx, g_nccl.AllGather(u_local_dev, u_all_dev, (size_t)count_local, ncclDouble, cm->comm, ctx->stream));
  return B200_OK;
}

// Enqueue (asynchronously, on the ctx stream) the reduction of one rank's counters: sums of nprob / nsuccess / total_nsteps /
// total_njvp, maxima of max_nsteps / worst_resid_inf.  b200_ens_allreduce_stats_finish() synchronises and writes the struct.
int32_t b200_ens_allreduce_stats_begin(b200_comm* cm, const b200_ens_result* local_host) {
  B200_DEVICE_GUARD(cm ? cm->ctx : nullptr);
  if (!cm || !local_host) return B200_ERR_INVALID;
  b200_ctx* ctx = cm->ctx;
  double* h = cm->h_stats;
  h[0] = local_host->nprob; h[1] = local_host->nsuccess; h[2] = (double)local_host->total_nsteps; h[3] = (double)local_host->total_njvp;
  h[4] = local_host->max_nsteps;
  h[5] = (local_host->worst_resid_inf == local_host->worst_resid_inf) ? local_host->worst_resid_inf : INFINITY;  // NaN must win the max
  CUDA_TRY(ctx, cudaMemcpyAsync(cm->d_stats, h, 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_TRY(ctx, g_nccl.AllReduce(cm->d_stats, cm->d_stats + 8, 4, ncclDouble, ncclSum, cm->comm, ctx->stream));
  NCCL_TRY(ctx, g_nccl.AllReduce(cm->d_stats + 4, cm->d_stats + 12, 2, ncclDouble, ncclMax, cm->comm, ctx->stream));
  return B200_OK;
}
int32_t b200_ens_allreduce_stats_finish(b200_comm* cm, b200_ens_result* global_host) {
  B200_DEVICE_GUARD(cm ? cm->ctx : nullptr);
  if (!cm || !global_host) return B200_ERR_INVALID;
  b200_ctx* ctx = cm->ctx;
  double* h = cm->h_stats + 8;
  CUDA_TRY(ctx, cudaMemcpyAsync(h, cm->d_stats + 8, 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  memset(global_host, 0, sizeof(*global_host));
  global_host->nprob = (int32_t)h[0]; global_host->nsuccess = (int32_t)h[1];
  global_host->total_nsteps = (int64_t)h[2]; global_host->total_njvp = (int64_t)h[3];
  global_host->max_nsteps = (int32_t)h[4]; global_host->worst_resid_inf = h[5];
  return B200_OK;
}
int32_t b200_ens_allreduce_stats(b200_comm* cm, const b200_ens_result* local_host, b200_ens_result* global_host) {
  B200_TRY(b200_ens_allreduce_stats_begin(cm, local_host));
  return b200_ens_allreduce_stats_finish(cm, global_host);
}
}  // extern "C"
