// comm.cu -- multi-GPU plumbing.  The reference has no distributed code; this is new surface
// (SURVEY.md 8e).  Two transports behind one internal interface:
//   * NCCL   one context per PROCESS / GPU (torchrun); libnccl is loaded with dlopen so that single-GPU
//            users need no NCCL at all; big buffers move through CUDA-IPC peer mappings over NVLink;
//   * local  G contexts of ONE process, one host thread per rank, on the same or on different devices
//            ("the sharded path must also run with G ranks mapped onto one device").
//            Collectives are pointer exchanges through a shared group object plus cudaMemcpy; the
//            barrier is a host barrier after a stream synchronise.  No kernel ever waits for another
//            rank's kernel, so G ranks can share one GPU.
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
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

static NcclApi load_nccl_api() {
  NcclApi a;
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

NcclApi& api() {          // thread-safe one-time initialisation
  static NcclApi a = load_nccl_api();
  return a;
}

// ---- the in-process transport ------------------------------------------------------------------
struct LocalGroup {
  int nranks = 0;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  uint64_t generation = 0;
  int attached = 0;
  bool broken = false;                         // a rank failed: every later collective errors out
  void* posted[sdpsr_ctx::MAX_RANKS] = {};     // pointer exchange slots (valid between two barriers)
  sdpsr_ctx* ctxs[sdpsr_ctx::MAX_RANKS] = {};
};

// false when a rank did not show up within the timeout (10 minutes, SDPSR_LOCAL_BARRIER_TIMEOUT_S; it failed or its thread died): the caller
// reports an error instead of hanging the process
bool local_host_barrier(LocalGroup* g) {
  std::unique_lock<std::mutex> lk(g->mu);
  if (g->broken) return false;
  const uint64_t gen = g->generation;
  if (++g->arrived == g->nranks) {
    g->arrived = 0;
    ++g->generation;
    g->cv.notify_all();
    return true;
  }
  static const int timeout_s = [] {
    const char* e = getenv("SDPSR_LOCAL_BARRIER_TIMEOUT_S");
    const int v = e ? atoi(e) : 0;
    return v > 0 ? v : 600;
  }();
  if (!g->cv.wait_for(lk, std::chrono::seconds(timeout_s), [&] { return g->generation != gen || g->broken; }) || g->broken) {
    g->broken = true;
    g->cv.notify_all();
    return false;
  }
  return true;
}

LocalGroup* local_of(sdpsr_ctx* ctx) { return reinterpret_cast<LocalGroup*>(ctx->local_group); }

// stream-drain + host barrier: every rank's previously enqueued work is complete afterwards
int local_barrier(sdpsr_ctx* ctx) {
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!local_host_barrier(local_of(ctx)))
    return ctx->fail(SDPSR_E_NCCL, "local communicator: a rank did not reach the barrier (it failed or left)");
  return SDPSR_OK;
}

// copy `bytes` from rank r's device pointer into mine (same or another device of this process)
int local_copy_from(sdpsr_ctx* ctx, void* dst, int r, const void* src, size_t bytes) {
  LocalGroup* g = local_of(ctx);
  const int sdev = g->ctxs[r]->device;
  if (sdev == ctx->device)
    SDPSR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  else
    SDPSR_CUDA(cudaMemcpyPeerAsync(dst, ctx->device, src, sdev, bytes, ctx->stream));
  return SDPSR_OK;
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
                                 int ntilecols, bool snake) {
  if (ctx->nranks <= 1) return SDPSR_OK;
  Timed tm(ctx, SDPSR_K_MISC, (double)ldc * (double)ncols * 8.0);
  if (ctx->local_group) {
    LocalGroup* g = local_of(ctx);
    g->posted[ctx->rank] = C;
    SDPSR_TRY(local_barrier(ctx));                 // every owner's slabs are complete and posted
    for (int tn = 0; tn < ntilecols; ++tn) {
      const int owner = snake ? sdpsr_tilecol_owner_snake(tn, ctx->nranks) : tn % ctx->nranks;
      if (owner == ctx->rank) continue;
      const int64_t c0 = (int64_t)tn * tile_cols;
      const int64_t w = std::min<int64_t>(tile_cols, ncols - c0);
      SDPSR_TRY(local_copy_from(ctx, C + ldc * c0, owner, reinterpret_cast<double*>(g->posted[owner]) + ldc * c0,
                                (size_t)(ldc * w) * sizeof(double)));
    }
    return local_barrier(ctx);                     // nobody overwrites a slab that is still being read
  }
  NCCL_TRY(api().GroupStart());
  for (int tn = 0; tn < ntilecols; ++tn) {
    const int64_t c0 = (int64_t)tn * tile_cols;
    const int64_t w = std::min<int64_t>(tile_cols, ncols - c0);
    double* slab = C + ldc * c0;
    NCCL_TRY(api().Broadcast(slab, slab, (size_t)(ldc * w), NCCL_FLOAT64, snake ? sdpsr_tilecol_owner_snake(tn, ctx->nranks) : tn % ctx->nranks, (ncclComm_t)ctx->nccl,
                             ctx->stream));
  }
  NCCL_TRY(api().GroupEnd());
  return SDPSR_OK;
}

int sdpsr_comm_bcast(sdpsr_ctx* ctx, void* buf, size_t bytes, int root) {
  if (ctx->nranks <= 1) return SDPSR_OK;
  if (ctx->local_group) {
    LocalGroup* g = local_of(ctx);
    g->posted[ctx->rank] = buf;
    SDPSR_TRY(local_barrier(ctx));
    if (ctx->rank != root) SDPSR_TRY(local_copy_from(ctx, buf, root, g->posted[root], bytes));
    return local_barrier(ctx);
  }
  NCCL_TRY(api().Broadcast(buf, buf, bytes, NCCL_UINT8, root, (ncclComm_t)ctx->nccl, ctx->stream));
  return SDPSR_OK;
}

// recv holds nranks blocks of `bytes`; block `rank` is this rank's contribution (already in place)
int sdpsr_comm_allgather(sdpsr_ctx* ctx, void* recv, size_t bytes) {
  if (ctx->nranks <= 1) return SDPSR_OK;
  unsigned char* base = reinterpret_cast<unsigned char*>(recv);
  if (ctx->local_group) {
    LocalGroup* g = local_of(ctx);
    g->posted[ctx->rank] = recv;
    SDPSR_TRY(local_barrier(ctx));
    for (int r = 0; r < ctx->nranks; ++r)
      if (r != ctx->rank)
        SDPSR_TRY(local_copy_from(ctx, base + bytes * r, r, reinterpret_cast<unsigned char*>(g->posted[r]) + bytes * r, bytes));
    return local_barrier(ctx);
  }
  NCCL_TRY(api().AllGather(base + bytes * ctx->rank, base, bytes, NCCL_UINT8, (ncclComm_t)ctx->nccl, ctx->stream));
  return SDPSR_OK;
}

// In-place all-gather of unequal blocks: rank r owns bytes [offset[r], offset[r] + bytes[r]) of `base`
int sdpsr_comm_allgatherv(sdpsr_ctx* ctx, void* base, const size_t* offset, const size_t* bytes) {
  if (ctx->nranks <= 1) return SDPSR_OK;
  unsigned char* b = reinterpret_cast<unsigned char*>(base);
  if (ctx->local_group) {
    LocalGroup* g = local_of(ctx);
    g->posted[ctx->rank] = base;
    SDPSR_TRY(local_barrier(ctx));
    for (int r = 0; r < ctx->nranks; ++r)
      if (r != ctx->rank && bytes[r])
        SDPSR_TRY(local_copy_from(ctx, b + offset[r], r, reinterpret_cast<unsigned char*>(g->posted[r]) + offset[r], bytes[r]));
    return local_barrier(ctx);
  }
  NCCL_TRY(api().GroupStart());
  for (int r = 0; r < ctx->nranks; ++r)
    if (bytes[r])
      NCCL_TRY(api().Broadcast(b + offset[r], b + offset[r], bytes[r], NCCL_UINT8, r, (ncclComm_t)ctx->nccl, ctx->stream));
  NCCL_TRY(api().GroupEnd());
  return SDPSR_OK;
}

// Every rank passes a host flag; all receive the minimum (used to agree on a code path: a branch that
// contains collectives must be taken by every rank or by none).  Blocking.
int sdpsr_comm_agree_min(sdpsr_ctx* ctx, int* flag) {
  if (ctx->nranks <= 1) return SDPSR_OK;
  int* d_all = nullptr;
  SDPSR_TRY(sdpsr_scratch_t(ctx, 28, (size_t)sdpsr_ctx::MAX_RANKS, &d_all));
  SDPSR_CUDA(cudaMemcpyAsync(d_all + ctx->rank, flag, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  SDPSR_TRY(sdpsr_comm_allgather(ctx, d_all, sizeof(int)));
  int h[sdpsr_ctx::MAX_RANKS];
  SDPSR_CUDA(cudaMemcpyAsync(h, d_all, sizeof(int) * (size_t)ctx->nranks, cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  int m = h[0];
  for (int r = 1; r < ctx->nranks; ++r) m = std::min(m, h[r]);
  *flag = m;
  return SDPSR_OK;
}

// true when another rank of the in-process transport computes on the same device: its persistent kernels then
// compete with this rank's for the SMs, so "every CTA of the launch is co-resident" does not hold
bool sdpsr_comm_shares_device(const sdpsr_ctx* ctx) {
  if (!ctx->local_group) return false;
  LocalGroup* g = (LocalGroup*)ctx->local_group;
  for (int r = 0; r < g->nranks; ++r)
    if (g->ctxs[r] && g->ctxs[r] != ctx && g->ctxs[r]->device == ctx->device) return true;
  return false;
}

int sdpsr_comm_barrier(sdpsr_ctx* ctx) {
  if (ctx->nranks <= 1) return SDPSR_OK;
  if (ctx->local_group) return local_barrier(ctx);
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

// local transport: the peers are contexts of this process, their buffers are plain device pointers
static int setup_peer_buffers_local(sdpsr_ctx* ctx) {
  LocalGroup* g = local_of(ctx);
  const int G = ctx->nranks;
  SDPSR_TRY(sdpsr_ensure_matrix(ctx, &ctx->T));
  SDPSR_TRY(sdpsr_ensure_matrix(ctx, &ctx->W));
  SDPSR_TRY(local_barrier(ctx));                   // every rank has registered and allocated
  bool ok = true;
  for (int r = 0; r < G; ++r) {
    sdpsr_ctx* o = g->ctxs[r];
    if (o->device != ctx->device) {
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, ctx->device, o->device) != cudaSuccess || !can) ok = false;
      if (ok) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
        cudaGetLastError();
      }
    }
    ctx->peer_ptr[0][r] = o->X2;
    ctx->peer_ptr[1][r] = o->T;
    ctx->peer_ptr[2][r] = o->W;
  }
  SDPSR_CUDA(cudaMalloc(&ctx->d_peer, sizeof(double*) * 3 * sdpsr_ctx::MAX_RANKS));
  SDPSR_CUDA(cudaMemcpyAsync(ctx->d_peer, ctx->peer_ptr, sizeof(double*) * 3 * sdpsr_ctx::MAX_RANKS,
                             cudaMemcpyHostToDevice, ctx->stream));
  int flag = ok ? 1 : 0;
  SDPSR_TRY(sdpsr_comm_agree_min(ctx, &flag));
  ctx->peer_ok = flag == 1 && !(ctx->flags & SDPSR_F_NCCL_EXCHANGE);
  return SDPSR_OK;
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
  if (ctx->local_group) {
    // collective: nobody frees a buffer that a peer's kernel may still be storing into
    LocalGroup* g = local_of(ctx);
    local_host_barrier(g);
    bool last;
    {
      std::lock_guard<std::mutex> lk(g->mu);
      g->ctxs[ctx->rank] = nullptr;
      last = --g->attached == 0;
    }
    if (last) delete g;
    ctx->local_group = nullptr;
    for (int b = 0; b < 3; ++b)
      for (int r = 0; r < sdpsr_ctx::MAX_RANKS; ++r) ctx->peer_ptr[b][r] = nullptr;
  }
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

extern "C" int sdpsr_comm_local_group(void** group, int nranks) {
  if (!group || nranks < 1 || nranks > sdpsr_ctx::MAX_RANKS) return SDPSR_E_INVALID;
  LocalGroup* g = new LocalGroup();
  g->nranks = nranks;
  *group = g;
  return SDPSR_OK;
}

extern "C" int sdpsr_comm_init_local(sdpsr_ctx* ctx, void* group, int rank) {
  if (!ctx) return SDPSR_E_INVALID;
  if (cudaSetDevice(ctx->device) != cudaSuccess) return ctx->fail(SDPSR_E_CUDA, "cudaSetDevice failed");
  LocalGroup* g = reinterpret_cast<LocalGroup*>(group);
  SDPSR_REQUIRE(g && rank >= 0 && rank < g->nranks, SDPSR_E_INVALID, "bad local communicator arguments");
  sdpsr_comm_free(ctx);
  {
    std::lock_guard<std::mutex> lk(g->mu);
    SDPSR_REQUIRE(g->ctxs[rank] == nullptr, SDPSR_E_INVALID, "rank already attached to this group");
    g->ctxs[rank] = ctx;
    ++g->attached;
  }
  ctx->local_group = g;
  ctx->nranks = g->nranks;
  ctx->rank = rank;
  if (g->nranks > 1) SDPSR_TRY(setup_peer_buffers_local(ctx));
  return SDPSR_OK;
}

extern "C" int sdpsr_comm_info(sdpsr_ctx* ctx, int* nranks, int* rank) {
  if (!ctx) return SDPSR_E_INVALID;
  if (nranks) *nranks = ctx->nranks;
  if (rank) *rank = ctx->rank;
  return SDPSR_OK;
}
