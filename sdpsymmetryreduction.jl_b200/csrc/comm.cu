// comm.cu -- multi-GPU plumbing: one context per process / GPU, NCCL loaded with dlopen
// so that single-GPU users need no NCCL at all.  The reference has no distributed code;
// this is new surface.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "sdpsr_internal.cuh"

namespace {

struct ncclUniqueIdLike {
  char internal[128];
};
typedef void* ncclComm_t;
typedef int ncclResult_t;

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueIdLike*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueIdLike, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& api() {
  static NcclApi a;
  if (a.lib) return a;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nme : names) {
    a.lib = dlopen(nme, RTLD_NOW | RTLD_GLOBAL);
    if (a.lib) break;
  }
  if (!a.lib) return a;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.lib, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.lib, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
  a.AllGather = (decltype(a.AllGather))dlsym(a.lib, "ncclAllGather");
  a.AllReduce = (decltype(a.AllReduce))dlsym(a.lib, "ncclAllReduce");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
  a.Broadcast = (decltype(a.Broadcast))dlsym(a.lib, "ncclBroadcast");
  a.GroupStart = (decltype(a.GroupStart))dlsym(a.lib, "ncclGroupStart");
  a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.lib, "ncclGroupEnd");
  a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.AllReduce && a.Broadcast &&
         a.GroupStart && a.GroupEnd;
  return a;
}

}  // namespace

constexpr int NCCL_UINT8 = 1, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8, NCCL_MAX = 2;

#define NCCL_TRY(call)                                                                               \
  do {                                                                                               \
    ncclResult_t _r = (call);                                                                        \
    if (_r != 0)                                                                                     \
      return ctx->fail(SDPSR_E_NCCL, std::string(#call) + ": " +                                     \
                                         (api().GetErrorString ? api().GetErrorString(_r) : "?"));  \
  } while (0)

// Tile-column tn (columns [tn*tile_cols, ...)) of the column-major matrix C is owned by rank
// tn % nranks; it is one contiguous slab.  One grouped launch broadcasts every slab from its owner.
int sdpsr_comm_exchange_tilecols(sdpsr_ctx* ctx, double* C, int64_t ldc, int64_t ncols, int tile_cols,
                                 int ntilecols) {
  if (ctx->nranks <= 1) return SDPSR_OK;
  Timed tm(ctx, SDPSR_K_MISC, (double)ldc * (double)ncols * 8.0);
  NCCL_TRY(api().GroupStart());
  for (int tn = 0; tn < ntilecols; ++tn) {
    const int64_t c0 = (int64_t)tn * tile_cols;
    const int64_t w = std::min<int64_t>(tile_cols, ncols - c0);
    double* slab = C + ldc * c0;
    NCCL_TRY(api().Broadcast(slab, slab, (size_t)(ldc * w), NCCL_FLOAT64, tn % ctx->nranks, (ncclComm_t)ctx->nccl,
                             ctx->stream));
  }
  NCCL_TRY(api().GroupEnd());
  return SDPSR_OK;
}

int sdpsr_comm_bcast(sdpsr_ctx* ctx, void* buf, size_t bytes, int root) {
  if (ctx->nranks <= 1) return SDPSR_OK;
  NCCL_TRY(api().Broadcast(buf, buf, bytes, NCCL_UINT8, root, (ncclComm_t)ctx->nccl, ctx->stream));
  return SDPSR_OK;
}

int sdpsr_comm_allreduce_max_u64(sdpsr_ctx* ctx, unsigned long long* buf, size_t count) {
  if (ctx->nranks <= 1) return SDPSR_OK;
  NCCL_TRY(api().AllReduce(buf, buf, count, NCCL_UINT64, NCCL_MAX, (ncclComm_t)ctx->nccl, ctx->stream));
  return SDPSR_OK;
}

int sdpsr_comm_barrier(sdpsr_ctx* ctx) {
  if (ctx->nranks <= 1) return SDPSR_OK;
  NCCL_TRY(api().AllReduce(ctx->d_barrier, ctx->d_barrier, 1, /*ncclInt32*/ 2, /*ncclSum*/ 0, (ncclComm_t)ctx->nccl,
                           ctx->stream));
  return SDPSR_OK;
}

// device table of the G peer copies of output buffer C, or NULL when C is not a registered buffer
double* const* sdpsr_comm_peer_table(sdpsr_ctx* ctx, const double* C) {
  if (!ctx->peer_ok) return nullptr;
  const double* mine[3] = {ctx->X2, ctx->T, ctx->W};
  for (int b = 0; b < 3; ++b)
    if (C == mine[b]) return ctx->d_peer + b * sdpsr_ctx::MAX_RANKS;
  return nullptr;
}

// Map every rank's X2 / T / W into this process (CUDA IPC over NVLink peer access).
static int setup_peer_buffers(sdpsr_ctx* ctx) {
  const int G = ctx->nranks;
  SDPSR_REQUIRE(G <= sdpsr_ctx::MAX_RANKS, SDPSR_E_UNSUPPORTED, "at most 16 ranks");
  SDPSR_TRY(sdpsr_ensure_matrix(ctx, &ctx->T));
  SDPSR_TRY(sdpsr_ensure_matrix(ctx, &ctx->W));
  SDPSR_CUDA(cudaMalloc(&ctx->d_barrier, sizeof(int)));
  SDPSR_CUDA(cudaMemsetAsync(ctx->d_barrier, 0, sizeof(int), ctx->stream));
  double* mine[3] = {ctx->X2, ctx->T, ctx->W};
  cudaIpcMemHandle_t hs[3];
  for (int b = 0; b < 3; ++b) SDPSR_CUDA(cudaIpcGetMemHandle(&hs[b], mine[b]));
  unsigned char* d_h = nullptr;
  const size_t hb = sizeof(cudaIpcMemHandle_t) * 3;
  SDPSR_CUDA(cudaMalloc(&d_h, hb * G));
  SDPSR_CUDA(cudaMemcpyAsync(d_h + hb * ctx->rank, hs, hb, cudaMemcpyHostToDevice, ctx->stream));
  NCCL_TRY(api().AllGather(d_h + hb * ctx->rank, d_h, hb, NCCL_UINT8, (ncclComm_t)ctx->nccl, ctx->stream));
  std::vector<cudaIpcMemHandle_t> all((size_t)3 * G);
  SDPSR_CUDA(cudaMemcpyAsync(all.data(), d_h, hb * G, cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_h);
  bool ok = true;
  for (int r = 0; r < G && ok; ++r)
    for (int b = 0; b < 3; ++b) {
      if (r == ctx->rank) {
        ctx->peer_ptr[b][r] = mine[b];
        continue;
      }
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[(size_t)3 * r + b], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = false;
        break;
      }
      ctx->peer_ptr[b][r] = reinterpret_cast<double*>(p);
    }
  if (ok) {
    SDPSR_CUDA(cudaMalloc(&ctx->d_peer, sizeof(double*) * 3 * sdpsr_ctx::MAX_RANKS));
    SDPSR_CUDA(cudaMemcpyAsync(ctx->d_peer, ctx->peer_ptr, sizeof(double*) * 3 * sdpsr_ctx::MAX_RANKS,
                               cudaMemcpyHostToDevice, ctx->stream));
    SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  // every rank must take the same path: agree on success
  int flag = ok ? 1 : 0, *d_f = ctx->d_barrier;
  SDPSR_CUDA(cudaMemcpyAsync(d_f, &flag, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  NCCL_TRY(api().AllReduce(d_f, d_f, 1, /*ncclInt32*/ 2, /*ncclMin*/ 3, (ncclComm_t)ctx->nccl, ctx->stream));
  SDPSR_CUDA(cudaMemcpyAsync(&flag, d_f, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  SDPSR_CUDA(cudaMemsetAsync(ctx->d_barrier, 0, sizeof(int), ctx->stream));
  ctx->peer_ok = flag == 1 && !(ctx->flags & SDPSR_F_NCCL_EXCHANGE);
  return SDPSR_OK;
}

void sdpsr_comm_free(sdpsr_ctx* ctx) {
  // Remote stores only happen inside sharded GEMMs, which end with a cross-rank barrier, so none is in
  // flight here.  Order: unmap the peers' buffers, rendezvous, and only then let the caller free ours
  // (freeing exported memory that is still mapped elsewhere is undefined).  Every rank must destroy its
  // context (collective, like the communicator itself).
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (int b = 0; b < 3; ++b)
    for (int r = 0; r < sdpsr_ctx::MAX_RANKS; ++r) {
      if (ctx->peer_ptr[b][r] && r != ctx->rank) cudaIpcCloseMemHandle(ctx->peer_ptr[b][r]);
      ctx->peer_ptr[b][r] = nullptr;
    }
  if (ctx->nccl && api().ok && ctx->d_barrier) {
    api().AllReduce(ctx->d_barrier, ctx->d_barrier, 1, 2, 0, (ncclComm_t)ctx->nccl, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
  }
  cudaFree(ctx->d_peer);
  cudaFree(ctx->d_barrier);
  ctx->d_peer = nullptr;
  ctx->d_barrier = nullptr;
  ctx->peer_ok = false;
  if (ctx->nccl && api().ok) api().CommDestroy((ncclComm_t)ctx->nccl);
  cudaFree(ctx->d_tiles);
  ctx->d_tiles = nullptr;
  ctx->tile_alloc = 0;
  ctx->nccl = nullptr;
  ctx->nranks = 1;
  ctx->rank = 0;
}

extern "C" int sdpsr_comm_unique_id(void* id128) {
  if (!id128) return SDPSR_E_INVALID;
  if (!api().ok) return SDPSR_E_NCCL;
  ncclUniqueIdLike id;
  if (api().GetUniqueId(&id) != 0) return SDPSR_E_NCCL;
  std::memcpy(id128, &id, 128);
  return SDPSR_OK;
}

extern "C" int sdpsr_comm_init(sdpsr_ctx* ctx, int nranks, int rank, const void* id128) {
  if (!ctx) return SDPSR_E_INVALID;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return ctx->fail(SDPSR_E_CUDA, "cudaSetDevice failed");
  SDPSR_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks && id128, SDPSR_E_INVALID, "bad communicator arguments");
  SDPSR_REQUIRE(api().ok, SDPSR_E_NCCL, "libnccl.so.2 could not be loaded");
  sdpsr_comm_free(ctx);
  ncclUniqueIdLike id;
  std::memcpy(&id, id128, 128);
  ncclComm_t comm = nullptr;
  const ncclResult_t r = api().CommInitRank(&comm, nranks, id, rank);
  SDPSR_REQUIRE(r == 0, SDPSR_E_NCCL,
                std::string("ncclCommInitRank failed: ") + (api().GetErrorString ? api().GetErrorString(r) : "?"));
  ctx->nccl = comm;
  ctx->nranks = nranks;
  ctx->rank = rank;
  if (nranks > 1) SDPSR_TRY(setup_peer_buffers(ctx));
  return SDPSR_OK;
}

extern "C" int sdpsr_comm_info(sdpsr_ctx* ctx, int* nranks, int* rank) {
  if (!ctx) return SDPSR_E_INVALID;
  if (nranks) *nranks = ctx->nranks;
  if (rank) *rank = ctx->rank;
  return SDPSR_OK;
}
