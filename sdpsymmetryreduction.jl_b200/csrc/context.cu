// context.cu -- lifetime, buffers, host<->device staging and the Partition part of
// the C ABI (include/sdpsr.h).  Reference sites are cited per entry point there.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include <cstdlib>

#include "sdpsr_internal.cuh"

static thread_local std::string g_create_error;

// ---------------------------------------------------------------------------
// timing
// ---------------------------------------------------------------------------
static cudaEvent_t take_event(sdpsr_ctx* c) {
  if (!c->ev_pool.empty()) {
    cudaEvent_t e = c->ev_pool.back();
    c->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

static void drain_events(sdpsr_ctx* c) {
  if (c->ev_pending.empty()) return;
  cudaStreamSynchronize(c->stream);
  for (auto& p : c->ev_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      c->t_ms[p.family] += ms;
      c->t_launch[p.family] += 1;
      c->t_work[p.family] += p.work;
    }
    c->ev_pool.push_back(p.a);
    c->ev_pool.push_back(p.b);
  }
  c->ev_pending.clear();
}

Timed::Timed(sdpsr_ctx* ctx, int family, double work_) : c(ctx), fam(family), work(work_) {
  if (c->flags & SDPSR_F_TIMING) {
    a = take_event(c);
    b = take_event(c);
    cudaEventRecord(a, c->stream);
  }
}
Timed::~Timed() {
  if (a) {
    cudaEventRecord(b, c->stream);
    c->ev_pending.push_back(EventPair{a, b, fam, work});
    if (c->ev_pending.size() > 4096) drain_events(c);
  }
}

int sdpsr_scratch(sdpsr_ctx* ctx, int slot, size_t bytes, void** out) {
  if (slot < 0 || slot >= sdpsr_ctx::SCRATCH_SLOTS) return ctx->fail(SDPSR_E_INVALID, "bad scratch slot");
  if (bytes < 256) bytes = 256;
  if (ctx->scratch_bytes[slot] < bytes) {
    cudaFree(ctx->scratch_ptr[slot]);
    ctx->scratch_ptr[slot] = nullptr;
    ctx->scratch_bytes[slot] = 0;
    const size_t want = bytes + bytes / 4;
    if (cudaMalloc(&ctx->scratch_ptr[slot], want) != cudaSuccess) {
      cudaGetLastError();
      return ctx->fail(SDPSR_E_ALLOC, "scratch allocation of " + std::to_string(want) + " bytes failed");
    }
    ctx->scratch_bytes[slot] = want;
  }
  *out = ctx->scratch_ptr[slot];
  return SDPSR_OK;
}

// ---------------------------------------------------------------------------
// small kernels for staging
// ---------------------------------------------------------------------------
namespace {

template <typename T>
__global__ void labels_in_kernel(const T* __restrict__ src, uint32_t* __restrict__ dst, int64_t n, int64_t ld,
                                 uint32_t* __restrict__ bad) {
  const int64_t j = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ld; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t v = 0u;
    if (i < n) {
      const long long s = (long long)src[i + n * j];
      if (s < 0 || s > 0xfffffffell) {
        *bad = 1u;
      } else {
        v = (uint32_t)s;
      }
    }
    dst[i + ld * j] = v;
  }
}

template <typename T>
__global__ void labels_out_kernel(const uint32_t* __restrict__ src, T* __restrict__ dst, int64_t total,
                                  uint64_t maxval, uint32_t* __restrict__ bad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t v = src[i];
    if ((uint64_t)v > maxval) *bad = 1u;
    dst[i] = (T)v;
  }
}

__global__ void count_zero_kernel(const uint32_t* __restrict__ labels, int64_t n, int64_t ld,
                                  unsigned long long* __restrict__ out) {
  const int64_t j = blockIdx.y;
  unsigned long long c = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    c += labels[i + ld * j] == 0u ? 1ull : 0ull;
  for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

}  // namespace

// ---------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------
extern "C" int sdpsr_version(void) { return SDPSR_VERSION; }

extern "C" const char* sdpsr_last_error(const sdpsr_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int sdpsr_device_count(int* count) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (count) *count = (e == cudaSuccess) ? c : 0;
  return e == cudaSuccess ? SDPSR_OK : SDPSR_E_NO_DEVICE;
}

static int create_impl(sdpsr_ctx* ctx) {
  SDPSR_CUDA(cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  SDPSR_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
  ctx->sm_count = prop.multiProcessorCount;
  SDPSR_REQUIRE(prop.major >= 10, SDPSR_E_UNSUPPORTED,
                std::string("libsdpsr_cuda is built for sm_100a only; device is ") + prop.name);
  SDPSR_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  const size_t e = ctx->elems;
  SDPSR_CUDA(cudaMalloc(&ctx->labels, e * sizeof(uint32_t)));
  SDPSR_CUDA(cudaMalloc(&ctx->labels_alt, e * sizeof(uint32_t)));
  SDPSR_CUDA(cudaMalloc(&ctx->X, e * sizeof(double)));
  SDPSR_CUDA(cudaMalloc(&ctx->X2, e * sizeof(double)));
  SDPSR_CUDA(cudaMalloc(&ctx->d_scalars, 64 * sizeof(uint64_t)));
  SDPSR_CUDA(cudaMallocHost(&ctx->h_pinned, 4096));
  SDPSR_CUDA(cudaMemsetAsync(ctx->X, 0, e * sizeof(double), ctx->stream));
  SDPSR_CUDA(cudaMemsetAsync(ctx->X2, 0, e * sizeof(double), ctx->stream));
  SDPSR_CUDA(cudaMemsetAsync(ctx->labels_alt, 0, e * sizeof(uint32_t), ctx->stream));
  return sdpsr_partition_reset(ctx);
}

extern "C" int sdpsr_create(sdpsr_ctx** out, int64_t n, int device, uint32_t flags) {
  if (!out) return SDPSR_E_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_create_error = "no CUDA device visible: libsdpsr_cuda has no CPU fallback";
    cudaGetLastError();
    return SDPSR_E_NO_DEVICE;
  }
  if (n < 1 || n > 65520 || device < 0 || device >= ndev) {
    g_create_error = "sdpsr_create: need 1 <= n <= 65520 and a valid device index";
    return SDPSR_E_INVALID;
  }
  sdpsr_ctx* ctx = new sdpsr_ctx();
  ctx->device = device;
  ctx->n = n;
  ctx->ld = round_up(n, 16);
  ctx->elems = (size_t)ctx->ld * (size_t)n;
  ctx->flags = flags;
  if (const char* sl = getenv("SDPSR_I8_SLICES")) {   // default digit count of the INT8 square (2..8)
    const int v = atoi(sl);
    if (v >= 2 && v <= 8) ctx->i8_slices = v;
  }
  if (const char* pr = getenv("SDPSR_I8_PAIR")) ctx->i8_pair = atoi(pr) != 0 ? 1 : 0;
  if (const char* sg = getenv("SDPSR_I8_SEGBLOCKS")) {   // test hook: short K segments at small N
    const int v = atoi(sg);
    if (v >= 1) ctx->i8_segblocks = v;
  }
  const int st = create_impl(ctx);
  if (st != SDPSR_OK) {
    g_create_error = ctx->err;
    sdpsr_destroy(ctx);
    return st;
  }
  *out = ctx;
  return SDPSR_OK;
}

extern "C" int sdpsr_destroy(sdpsr_ctx* ctx) {
  if (!ctx) return SDPSR_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaEventDestroy(ctx->copy_gate);
    cudaEventDestroy(ctx->staged_ev);
    cudaEventDestroy(ctx->labels_ev);
    cudaStreamDestroy(ctx->copy_stream);
  }
  drain_events(ctx);
  for (auto e : ctx->ev_pool) cudaEventDestroy(e);
  sdpsr_comm_free(ctx);
  sdpsr_blockdiag_free(ctx);
  sdpsr_krylov_free(ctx);
  sdpsr_constraints_free(ctx);
  sdpsr_table_free(ctx->tab[0]);
  sdpsr_table_free(ctx->tab[1]);
  sdpsr_table_free(ctx->tab_scratch);
  sdpsr_table_free(ctx->tab_merge);
  cudaFree(ctx->clabels);
  cudaFree(ctx->labels);
  cudaFree(ctx->labels_alt);
  cudaFree(ctx->labels_tmp);
  cudaFree(ctx->X);
  cudaFree(ctx->X2);
  cudaFree(ctx->lut);
  cudaFree(ctx->d_values);
  cudaFree(ctx->rk_mi);
  cudaFree(ctx->bitmap);
  cudaFree(ctx->bm_block);
  cudaFree(ctx->d_scalars);
  for (int i = 0; i < sdpsr_ctx::SCRATCH_SLOTS; ++i) cudaFree(ctx->scratch_ptr[i]);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return SDPSR_OK;
}

#define CTX_ENTER()                                   \
  if (!ctx) return SDPSR_E_INVALID;                   \
  if (cudaSetDevice(ctx->device) != cudaSuccess) return ctx->fail(SDPSR_E_CUDA, "cudaSetDevice failed"); \
  ++ctx->api_seq

static int finish(sdpsr_ctx* ctx) {
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  SDPSR_CUDA(cudaGetLastError());
  return SDPSR_OK;
}

// ---------------------------------------------------------------------------
// Partition state
// ---------------------------------------------------------------------------
extern "C" int sdpsr_partition_reset(sdpsr_ctx* ctx) {
  CTX_ENTER();
  SDPSR_CUDA(cudaMemsetAsync(ctx->labels, 0, ctx->elems * sizeof(uint32_t), ctx->stream));
  KeyTable& t = ctx->tab[ctx->cur];
  SDPSR_TRY(sdpsr_table_alloc(ctx, t, 64));
  SDPSR_CUDA(cudaMemsetAsync(t.slots, 0xff, (size_t)t.cap * sizeof(KeySlot), ctx->stream));
  SDPSR_CUDA(cudaMemsetAsync(t.meta, 0, 4 * sizeof(uint32_t), ctx->stream));
  SDPSR_CUDA(cudaMemsetAsync(t.rank, 0, sizeof(uint32_t), ctx->stream));
  t.count = 0;
  ctx->dim = 0;
  ctx->part_epoch += 1;
  ctx->x_is_fill = false;
  ctx->x_valid = false;
  ctx->key_decodable = false;
  ctx->sym_state = 1;          // the empty partition is transpose-invariant
  ctx->labels_full = true;     // all zero everywhere
  ctx->clabels_valid = false;
  return finish(ctx);
}

static int stage_labels(sdpsr_ctx* ctx, const void* labels, int elt_bytes) {
  SDPSR_REQUIRE(labels != nullptr, SDPSR_E_INVALID, "labels is NULL");
  SDPSR_REQUIRE(elt_bytes == 1 || elt_bytes == 2 || elt_bytes == 4 || elt_bytes == 8, SDPSR_E_INVALID,
                "elt_bytes must be 1, 2, 4 or 8");
  SDPSR_TRY(sdpsr_ensure_tmp_labels(ctx));
  const size_t nn = (size_t)ctx->n * (size_t)ctx->n;
  void* raw = ctx->X2;   // staging: X2 holds no state between calls that matter here
  SDPSR_CUDA(cudaMemcpyAsync(raw, labels, nn * elt_bytes, cudaMemcpyDefault, ctx->stream));
  uint32_t* bad = ctx->d_scalars;
  SDPSR_CUDA(cudaMemsetAsync(bad, 0, sizeof(uint32_t), ctx->stream));
  dim3 grid((unsigned)std::min<int64_t>((ctx->ld + 255) / 256, 64), (unsigned)ctx->n);
  // 8-byte inputs are read as signed (Julia Int64); negative labels are rejected
  switch (elt_bytes) {
    case 1: labels_in_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>((const uint8_t*)raw, ctx->labels_tmp, ctx->n, ctx->ld, bad); break;
    case 2: labels_in_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>((const uint16_t*)raw, ctx->labels_tmp, ctx->n, ctx->ld, bad); break;
    case 4: labels_in_kernel<uint32_t><<<grid, 256, 0, ctx->stream>>>((const uint32_t*)raw, ctx->labels_tmp, ctx->n, ctx->ld, bad); break;
    default: labels_in_kernel<int64_t><<<grid, 256, 0, ctx->stream>>>((const int64_t*)raw, ctx->labels_tmp, ctx->n, ctx->ld, bad); break;
  }
  count_launch(ctx);
  SDPSR_CUDA(cudaGetLastError());
  uint32_t* hb = reinterpret_cast<uint32_t*>(ctx->h_pinned) + 16;
  SDPSR_CUDA(cudaMemcpyAsync(hb, bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  SDPSR_REQUIRE(*hb == 0u, SDPSR_E_INVALID, "labels must be integers in 0 .. 2^32-2 (src/partitions.jl:46)");
  // X2 served as staging; it is fully overwritten by the next product or upload.
  return SDPSR_OK;
}

extern "C" int sdpsr_partition_set_labels(sdpsr_ctx* ctx, const void* labels, int elt_bytes, int64_t* dim) {
  CTX_ENTER();
  SDPSR_TRY(stage_labels(ctx, labels, elt_bytes));
  RefineSpec sp;
  sp.mode = KM_PAIR;
  sp.lab2 = ctx->labels_tmp;
  sp.ignore_labels = true;
  sp.do_round = false;
  SDPSR_TRY(sdpsr_refine_pass(ctx, sp, dim));
  ctx->x_valid = false;
  return finish(ctx);
}

extern "C" int sdpsr_refine_labels(sdpsr_ctx* ctx, const void* labels, int elt_bytes, int64_t* dim) {
  CTX_ENTER();
  SDPSR_TRY(stage_labels(ctx, labels, elt_bytes));
  RefineSpec sp;
  sp.mode = KM_PAIR;
  sp.lab2 = ctx->labels_tmp;
  sp.do_round = false;
  SDPSR_TRY(sdpsr_refine_pass(ctx, sp, dim));
  return finish(ctx);
}

extern "C" int sdpsr_partition_get_labels(sdpsr_ctx* ctx, void* labels, int elt_bytes) {
  CTX_ENTER();
  SDPSR_REQUIRE(labels != nullptr, SDPSR_E_INVALID, "labels is NULL");
  SDPSR_REQUIRE(elt_bytes == 1 || elt_bytes == 2 || elt_bytes == 4 || elt_bytes == 8, SDPSR_E_INVALID,
                "elt_bytes must be 1, 2, 4 or 8");
  SDPSR_TRY(sdpsr_ensure_tmp_labels(ctx));
  const int64_t nn = ctx->n * ctx->n;
  SDPSR_TRY(sdpsr_canonical_labels(ctx, ctx->labels_tmp));
  if (elt_bytes == 4) {
    SDPSR_CUDA(cudaMemcpyAsync(labels, ctx->labels_tmp, (size_t)nn * 4, cudaMemcpyDefault, ctx->stream));
    return finish(ctx);
  }
  // the reference throws InexactError as soon as a label exceeds typemax(T)
  const uint64_t maxval = elt_bytes == 1 ? 0xffull : elt_bytes == 2 ? 0xffffull : ~0ull;
  SDPSR_REQUIRE((uint64_t)ctx->dim <= maxval, SDPSR_E_LABEL_OVERFLOW,
                "dim(P) does not fit the requested label type (InexactError in the reference)");
  uint32_t* bad = ctx->d_scalars;
  SDPSR_CUDA(cudaMemsetAsync(bad, 0, sizeof(uint32_t), ctx->stream));
  void* raw = ctx->X2;
  const int grid = (int)std::min<int64_t>((nn + 255) / 256, (int64_t)ctx->sm_count * 8);
  switch (elt_bytes) {
    case 1: labels_out_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>(ctx->labels_tmp, (uint8_t*)raw, nn, maxval, bad); break;
    case 2: labels_out_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(ctx->labels_tmp, (uint16_t*)raw, nn, maxval, bad); break;
    default: labels_out_kernel<uint64_t><<<grid, 256, 0, ctx->stream>>>(ctx->labels_tmp, (uint64_t*)raw, nn, maxval, bad); break;
  }
  count_launch(ctx);
  SDPSR_CUDA(cudaGetLastError());
  SDPSR_CUDA(cudaMemcpyAsync(labels, raw, (size_t)nn * elt_bytes, cudaMemcpyDefault, ctx->stream));
  return finish(ctx);
}

// ---------------------------------------------------------------------------
// Transfers on the copy stream (overlap with the kernels of ctx->stream)
// ---------------------------------------------------------------------------
// The large host transfers are done by a KERNEL (zero-copy loads / stores over PCIe, a few CTAs) rather than by the
// copy engines: a DMA engine serves the streams in submission order, so a 2 GB cudaMemcpyAsync on the copy stream
// made every small cudaMemcpyAsync of the compute stream (index arrays, counters, flags: every call has some) wait
// for it -- measured: no overlap at all.  Needs page-locked (mapped) host memory; anything else takes cudaMemcpyAsync.
__global__ void __launch_bounds__(256) pcie_copy_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {          // four independent 16-byte requests per thread in flight
    const uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
    dst[i] = a;
    dst[i + stride] = b;
    dst[i + 2 * stride] = c;
    dst[i + 3 * stride] = d;
  }
  for (; i < n16; i += stride) dst[i] = src[i];
}

// true when `host` is page-locked memory the device can address directly (cudaHostAlloc / cudaHostRegister)
static bool device_can_address(const void* host, void** dev) {
  cudaPointerAttributes pa;
  if (cudaPointerGetAttributes(&pa, host) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (pa.type != cudaMemoryTypeHost || pa.devicePointer == nullptr) return false;
  *dev = pa.devicePointer;
  return true;
}

// bytes % 16 == 0 and both pointers 16-byte aligned, else cudaMemcpyAsync
static int big_copy(sdpsr_ctx* ctx, void* dst, const void* src, size_t bytes, bool to_host) {
  void* mapped = nullptr;
  const void* host = to_host ? dst : src;
  if (bytes % 16 == 0 && ((uintptr_t)dst % 16 == 0) && ((uintptr_t)src % 16 == 0) && device_can_address(host, &mapped)) {
    uint4* d = reinterpret_cast<uint4*>(to_host ? mapped : dst);
    const uint4* s = reinterpret_cast<const uint4*>(to_host ? src : mapped);
    pcie_copy_kernel<<<32, 256, 0, ctx->copy_stream>>>(d, s, bytes / 16);
    count_launch(ctx);
    SDPSR_CUDA(cudaGetLastError());
    return SDPSR_OK;
  }
  SDPSR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->copy_stream));
  return SDPSR_OK;
}

int sdpsr_copy_stream(sdpsr_ctx* ctx) {
  if (ctx->copy_stream) return SDPSR_OK;
  SDPSR_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  SDPSR_CUDA(cudaEventCreateWithFlags(&ctx->copy_gate, cudaEventDisableTiming));
  SDPSR_CUDA(cudaEventCreateWithFlags(&ctx->staged_ev, cudaEventDisableTiming));
  SDPSR_CUDA(cudaEventCreateWithFlags(&ctx->labels_ev, cudaEventDisableTiming));
  return SDPSR_OK;
}

/* Start the upload of a HOST objective matrix C into the context (the staging copy sdpsr_init_partition would do
 * first) so that it overlaps the constraint set-up: stage, sdpsr_set_constraints_*, sdpsr_init_partition(same C).
 * Any other call in between drops the staging (init_partition then copies as usual).  A no-op for device-resident C.
 * A sharded context stages its own column block only.  Returns without waiting. */
extern "C" int sdpsr_stage_objective(sdpsr_ctx* ctx, const double* C) {
  CTX_ENTER();
  SDPSR_REQUIRE(C != nullptr, SDPSR_E_INVALID, "C is NULL");
  if (ctx->staged_src) {                              // an earlier staging that was never consumed
    SDPSR_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->staged_ev, 0));
    ctx->staged_src = nullptr;
  }
  cudaPointerAttributes pa;
  const bool on_device = cudaPointerGetAttributes(&pa, C) == cudaSuccess && pa.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  if (on_device) return SDPSR_OK;
  if (ctx->nranks > 1 && ctx->ld != ctx->n) return SDPSR_OK;
  SDPSR_TRY(sdpsr_copy_stream(ctx));
  // whatever ctx->stream still does with X comes first
  SDPSR_CUDA(cudaEventRecord(ctx->copy_gate, ctx->stream));
  SDPSR_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_gate, 0));
  if (ctx->nranks > 1) {
    // sharded: this rank's column block only (init_partition all-gathers the blocks over NVLink)
    const int64_t b0 = ctx->n * ctx->rank / ctx->nranks, b1 = ctx->n * (ctx->rank + 1) / ctx->nranks;
    if (b1 > b0)
      SDPSR_TRY(big_copy(ctx, ctx->X + b0 * ctx->ld, C + b0 * ctx->n, (size_t)(b1 - b0) * (size_t)ctx->n * 8, /*to_host=*/false));
  } else if (ctx->ld != ctx->n) {
    SDPSR_CUDA(cudaMemsetAsync(ctx->X, 0, ctx->elems * 8, ctx->copy_stream));
    SDPSR_CUDA(cudaMemcpy2DAsync(ctx->X, (size_t)ctx->ld * 8, C, (size_t)ctx->n * 8, (size_t)ctx->n * 8, (size_t)ctx->n,
                                 cudaMemcpyDefault, ctx->copy_stream));
  } else {
    SDPSR_TRY(big_copy(ctx, ctx->X, C, (size_t)ctx->n * (size_t)ctx->n * 8, /*to_host=*/false));
  }
  SDPSR_CUDA(cudaEventRecord(ctx->staged_ev, ctx->copy_stream));
  ctx->staged_src = C;
  ctx->staged_seq = ctx->api_seq;
  ctx->x_valid = false;
  ctx->x_is_fill = false;
  return SDPSR_OK;
}

/* sdpsr_partition_get_labels without the wait: the canonical labels are produced on ctx->stream into a staging
 * buffer of their own and travel to `labels` (pinned host memory for a real overlap) on the copy stream while later
 * calls compute; sdpsr_partition_labels_wait returns once they have arrived (and reports SDPSR_E_LABEL_OVERFLOW if a
 * label did not fit).  `labels` must stay valid until then. */
extern "C" int sdpsr_partition_get_labels_async(sdpsr_ctx* ctx, void* labels, int elt_bytes) {
  CTX_ENTER();
  SDPSR_REQUIRE(labels != nullptr, SDPSR_E_INVALID, "labels is NULL");
  SDPSR_REQUIRE(elt_bytes == 1 || elt_bytes == 2 || elt_bytes == 4 || elt_bytes == 8, SDPSR_E_INVALID,
                "elt_bytes must be 1, 2, 4 or 8");
  const uint64_t maxval = elt_bytes == 1 ? 0xffull : elt_bytes == 2 ? 0xffffull : ~0ull;
  SDPSR_REQUIRE(elt_bytes == 4 || (uint64_t)ctx->dim <= maxval, SDPSR_E_LABEL_OVERFLOW,
                "dim(P) does not fit the requested label type (InexactError in the reference)");
  SDPSR_TRY(sdpsr_copy_stream(ctx));
  if (ctx->labels_pending) {
    SDPSR_CUDA(cudaEventSynchronize(ctx->labels_ev));
    ctx->labels_pending = false;
  }
  const int64_t nn = ctx->n * ctx->n;
  unsigned char* stage = nullptr;
  SDPSR_TRY(sdpsr_scratch_t(ctx, 44, (size_t)nn * (size_t)elt_bytes, &stage));
  uint32_t* bad = ctx->d_scalars;
  uint32_t* h_bad = reinterpret_cast<uint32_t*>(ctx->h_pinned) + 384;
  SDPSR_CUDA(cudaMemsetAsync(bad, 0, sizeof(uint32_t), ctx->stream));
  if (elt_bytes == 4) {
    SDPSR_TRY(sdpsr_canonical_labels(ctx, reinterpret_cast<uint32_t*>(stage)));
  } else {
    SDPSR_TRY(sdpsr_ensure_tmp_labels(ctx));
    SDPSR_TRY(sdpsr_canonical_labels(ctx, ctx->labels_tmp));
    const int grid = (int)std::min<int64_t>((nn + 255) / 256, (int64_t)ctx->sm_count * 8);
    switch (elt_bytes) {
      case 1: labels_out_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>(ctx->labels_tmp, (uint8_t*)stage, nn, maxval, bad); break;
      case 2: labels_out_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(ctx->labels_tmp, (uint16_t*)stage, nn, maxval, bad); break;
      default: labels_out_kernel<uint64_t><<<grid, 256, 0, ctx->stream>>>(ctx->labels_tmp, (uint64_t*)stage, nn, maxval, bad); break;
    }
    count_launch(ctx);
    SDPSR_CUDA(cudaGetLastError());
  }
  // (the flag leaves on ctx->stream: d_scalars is reused by the calls that follow)
  SDPSR_CUDA(cudaMemcpyAsync(h_bad, bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaEventRecord(ctx->copy_gate, ctx->stream));
  SDPSR_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_gate, 0));
  SDPSR_TRY(big_copy(ctx, labels, stage, (size_t)nn * elt_bytes, /*to_host=*/true));
  SDPSR_CUDA(cudaEventRecord(ctx->labels_ev, ctx->copy_stream));
  ctx->labels_pending = true;
  return SDPSR_OK;
}

extern "C" int sdpsr_partition_labels_wait(sdpsr_ctx* ctx) {
  CTX_ENTER();
  if (!ctx->labels_pending) return SDPSR_OK;
  SDPSR_CUDA(cudaEventSynchronize(ctx->labels_ev));
  ctx->labels_pending = false;
  const uint32_t* h_bad = reinterpret_cast<const uint32_t*>(ctx->h_pinned) + 384;
  SDPSR_REQUIRE(*h_bad == 0u, SDPSR_E_LABEL_OVERFLOW, "a label does not fit the requested integer width");
  return SDPSR_OK;
}

extern "C" int sdpsr_partition_dim(sdpsr_ctx* ctx, int64_t* dim) {
  CTX_ENTER();
  SDPSR_REQUIRE(dim != nullptr, SDPSR_E_INVALID, "dim is NULL");
  *dim = ctx->dim;
  return SDPSR_OK;
}

extern "C" int sdpsr_partition_zero_count(sdpsr_ctx* ctx, int64_t* count) {
  CTX_ENTER();
  SDPSR_REQUIRE(count != nullptr, SDPSR_E_INVALID, "count is NULL");
  SDPSR_TRY(sdpsr_shard_ensure_full_labels(ctx));
  unsigned long long* d = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 2);
  SDPSR_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), ctx->stream));
  dim3 grid((unsigned)std::min<int64_t>((ctx->n + 255) / 256, 64), (unsigned)ctx->n);
  count_zero_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->labels, ctx->n, ctx->ld, d);
  count_launch(ctx);
  unsigned long long* h = reinterpret_cast<unsigned long long*>(ctx->h_pinned) + 16;
  SDPSR_CUDA(cudaMemcpyAsync(h, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_TRY(finish(ctx));
  *count = (int64_t)*h;
  return SDPSR_OK;
}

// _constraints(P) (src/diagonalize.jl:42-50): for every class the ascending column-major linear indices of
// its entries, as one CSR (ptr[dim + 1], idx[N^2 - #zeros]).  This is the generic slow path of the reference's
// AbstractPartition contract, not a hot path: the canonical label matrix is brought to the host once and
// the stable counting sort runs there.
extern "C" int sdpsr_partition_constraints(sdpsr_ctx* ctx, int64_t* ptr, uint32_t* idx, int64_t idx_len, int index_base) {
  CTX_ENTER();
  SDPSR_REQUIRE(ptr != nullptr && (idx != nullptr || idx_len == 0), SDPSR_E_INVALID, "ptr / idx is NULL");
  SDPSR_REQUIRE(index_base == 0 || index_base == 1, SDPSR_E_INVALID, "index_base must be 0 or 1");
  SDPSR_TRY(sdpsr_ensure_tmp_labels(ctx));
  const int64_t nn = ctx->n * ctx->n, d = ctx->dim;
  SDPSR_REQUIRE(nn + index_base <= 0xffffffffll, SDPSR_E_UNSUPPORTED, "linear indices do not fit UInt32");
  SDPSR_TRY(sdpsr_canonical_labels(ctx, ctx->labels_tmp));
  std::vector<uint32_t> lab((size_t)nn);
  SDPSR_CUDA(cudaMemcpyAsync(lab.data(), ctx->labels_tmp, (size_t)nn * 4, cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_TRY(finish(ctx));
  std::vector<int64_t> cnt((size_t)d + 2, 0);
  for (int64_t i = 0; i < nn; ++i) cnt[(size_t)lab[(size_t)i] + 1] += 1;       // cnt[c + 1] = |class c|
  int64_t total = 0;
  for (int64_t c = 1; c <= d; ++c) {
    ptr[c - 1] = total;
    total += cnt[(size_t)c + 1];
  }
  ptr[d] = total;
  SDPSR_REQUIRE(idx_len == total, SDPSR_E_INVALID, "idx_len must be N^2 minus the number of zero labels");
  std::vector<int64_t> cur(ptr, ptr + d);
  for (int64_t i = 0; i < nn; ++i) {
    const uint32_t c = lab[(size_t)i];
    if (c) idx[cur[(size_t)c - 1]++] = (uint32_t)(i + index_base);
  }
  return SDPSR_OK;
}

extern "C" int sdpsr_partition_is_symmetric(sdpsr_ctx* ctx, int* is_symmetric) {
  CTX_ENTER();
  SDPSR_REQUIRE(is_symmetric != nullptr, SDPSR_E_INVALID, "is_symmetric is NULL");
  return sdpsr_symmetric_check(ctx, is_symmetric);
}

// ---------------------------------------------------------------------------
// dense matrix staging
// ---------------------------------------------------------------------------
static int matrix_ptr(sdpsr_ctx* ctx, int which, double** p) {
  switch (which) {
    case SDPSR_MAT_X: *p = ctx->X; break;
    case SDPSR_MAT_X2: *p = ctx->X2; break;
    case SDPSR_MAT_Q: *p = ctx->Q; break;
    case SDPSR_MAT_W: *p = ctx->W; break;
    default: return ctx->fail(SDPSR_E_INVALID, "unknown matrix id");
  }
  if (!*p) {
    // Q and W are allocated on first use
    double** slot = which == SDPSR_MAT_Q ? &ctx->Q : &ctx->W;
    SDPSR_CUDA(cudaMalloc(slot, ctx->elems * sizeof(double)));
    SDPSR_CUDA(cudaMemsetAsync(*slot, 0, ctx->elems * sizeof(double), ctx->stream));
    *p = *slot;
  }
  return SDPSR_OK;
}

static int upload_matrix(sdpsr_ctx* ctx, double* dst, const double* src) {
  if (ctx->ld != ctx->n) SDPSR_CUDA(cudaMemsetAsync(dst, 0, ctx->elems * sizeof(double), ctx->stream));
  SDPSR_CUDA(cudaMemcpy2DAsync(dst, (size_t)ctx->ld * 8, src, (size_t)ctx->n * 8, (size_t)ctx->n * 8, (size_t)ctx->n,
                               cudaMemcpyDefault, ctx->stream));
  return SDPSR_OK;
}

extern "C" int sdpsr_get_matrix(sdpsr_ctx* ctx, int which, double* out) {
  CTX_ENTER();
  SDPSR_REQUIRE(out != nullptr, SDPSR_E_INVALID, "out is NULL");
  double* p = nullptr;
  SDPSR_TRY(matrix_ptr(ctx, which, &p));
  if (which == SDPSR_MAT_X) {
    SDPSR_REQUIRE(ctx->x_valid, SDPSR_E_STATE, "X is not defined yet (call sdpsr_fill first)");
    if (ctx->x_is_fill) SDPSR_TRY(sdpsr_materialize_fill(ctx, ctx->X));
  }
  SDPSR_CUDA(cudaMemcpy2DAsync(out, (size_t)ctx->n * 8, p, (size_t)ctx->ld * 8, (size_t)ctx->n * 8, (size_t)ctx->n,
                               cudaMemcpyDefault, ctx->stream));
  return finish(ctx);
}

extern "C" int sdpsr_set_matrix(sdpsr_ctx* ctx, int which, const double* in) {
  CTX_ENTER();
  SDPSR_REQUIRE(in != nullptr, SDPSR_E_INVALID, "in is NULL");
  double* p = nullptr;
  SDPSR_TRY(matrix_ptr(ctx, which, &p));
  SDPSR_TRY(upload_matrix(ctx, p, in));
  if (which == SDPSR_MAT_X) {
    ctx->x_valid = true;
    ctx->x_is_fill = false;
  }
  return finish(ctx);
}

extern "C" int sdpsr_gemm(sdpsr_ctx* ctx, int a, int b, int c) {
  CTX_ENTER();
  SDPSR_REQUIRE(c != a && c != b, SDPSR_E_INVALID, "output must not alias an input");
  double *pa, *pb, *pc;
  SDPSR_TRY(matrix_ptr(ctx, a, &pa));
  SDPSR_TRY(matrix_ptr(ctx, b, &pb));
  SDPSR_TRY(matrix_ptr(ctx, c, &pc));
  SDPSR_TRY(sdpsr_gemm_f64(ctx, pa, ctx->ld, pb, ctx->ld, pc, ctx->ld, ctx->ld, ctx->n, ctx->n, false));
  return finish(ctx);
}

// ---------------------------------------------------------------------------
// refine / fill / square
// ---------------------------------------------------------------------------
extern "C" int sdpsr_refine_values(sdpsr_ctx* ctx, const double* M, double atol, int do_round, int64_t* dim) {
  CTX_ENTER();
  SDPSR_REQUIRE(M != nullptr, SDPSR_E_INVALID, "M is NULL");
  SDPSR_TRY(upload_matrix(ctx, ctx->X2, M));
  SDPSR_TRY(sdpsr_generic_refine_values(ctx, ctx->X2, atol, do_round != 0, nullptr, dim));
  return finish(ctx);
}

extern "C" int sdpsr_fill(sdpsr_ctx* ctx, const double* values, int64_t len) {
  CTX_ENTER();
  SDPSR_REQUIRE(len == ctx->dim, SDPSR_E_INVALID, "length(values) != dim(P) (src/partitions.jl:69)");
  SDPSR_REQUIRE(values != nullptr || len == 0, SDPSR_E_INVALID, "values is NULL");
  SDPSR_TRY(sdpsr_upload_values(ctx, values, len));
  SDPSR_TRY(sdpsr_build_lut(ctx, ctx->d_values, len));
  ctx->x_is_fill = true;    // X == fill(S, lut); materialised on demand
  ctx->x_valid = true;
  return finish(ctx);
}

// X2 = X * X.  method -1: automatic (INT8 tensor path for symmetric X where it is the faster one,
// else DMMA), 0: DMMA, 1: INT8 (error if X is not symmetric).  While X is still the un-materialised
// fill(S, lut) its symmetry is the partition's (cached, 4 B/entry when it has to be checked) and the INT8
// path gathers its digit slices straight from the labels; X is only written out for the DMMA path.
static int square_x(sdpsr_ctx* ctx, int method, int slices, int bits, int* was_symmetric) {
  // X symmetric (bit for bit) => X*X symmetric: compute the lower tiles only and mirror them,
  // which also makes X2 exactly symmetric (SYRK-style, half the flops)
  int sym = 0;
  if (!(ctx->flags & SDPSR_F_NO_SYRK) || method == 1) {
    if (ctx->x_is_fill)
      SDPSR_TRY(sdpsr_symmetric_check(ctx, &sym));
    else
      SDPSR_TRY(sdpsr_matrix_symmetric(ctx, ctx->X, &sym));
  }
  if (was_symmetric) *was_symmetric = sym;
  bool use_i8 = false;
  if (method == 1) {
    SDPSR_REQUIRE(sym != 0, SDPSR_E_INVALID, "the INT8 square needs a bit-for-bit symmetric X");
    use_i8 = true;
  } else if (method < 0 && sym && !(ctx->flags & (SDPSR_F_NO_I8 | SDPSR_F_NO_SYRK)) && ctx->n <= 65536) {
    use_i8 = (ctx->flags & SDPSR_F_FORCE_I8) || ctx->n >= 2048;
  }
  if (use_i8) {
    int done = 0;
    SDPSR_TRY(sdpsr_square_i8(ctx, ctx->x_is_fill ? nullptr : ctx->X, ctx->X2, slices, bits, /*shard=*/true,
                              /*force_range=*/method == 1, &done));
    if (done) return SDPSR_OK;
    SDPSR_REQUIRE(method != 1, SDPSR_E_UNSUPPORTED, "the INT8 square does not handle Inf/NaN or extreme exponents");
  }
  if (ctx->x_is_fill) SDPSR_TRY(sdpsr_materialize_fill(ctx, ctx->X));     // S is unchanged: X stays a valid fill
  SDPSR_TRY(sdpsr_gemm_f64(ctx, ctx->X, ctx->ld, ctx->X, ctx->ld, ctx->X2, ctx->ld, ctx->ld, ctx->n, ctx->n,
                           sym != 0 && !(ctx->flags & SDPSR_F_NO_SYRK), /*shard=*/true));
  return SDPSR_OK;
}

extern "C" int sdpsr_square_round_refine(sdpsr_ctx* ctx, double atol, int64_t* dim) {
  CTX_ENTER();
  SDPSR_REQUIRE(ctx->x_valid, SDPSR_E_STATE, "X is not defined yet (call sdpsr_fill first)");
  int sym = 0;
  if (sdpsr_shard_active(ctx)) {        // the refine below reads this rank's column block of X2 only
    ctx->mirror_col0 = ctx->n * ctx->rank / ctx->nranks;
    ctx->mirror_col1 = ctx->n * (ctx->rank + 1) / ctx->nranks;
  }
  const int sq_status = square_x(ctx, /*method=*/-1, ctx->i8_slices, 0, &sym);
  ctx->mirror_col0 = 0;
  ctx->mirror_col1 = -1;
  SDPSR_TRY(sq_status);
  // X symmetric => X2 symmetric bit for bit (mirrored lower tiles): the refined partition stays symmetric
  SDPSR_TRY(sdpsr_generic_refine_values(ctx, ctx->X2, atol, true, nullptr, dim,
                                        /*keeps_symmetry=*/sym != 0 && !(ctx->flags & SDPSR_F_NO_SYRK)));
  return finish(ctx);
}

extern "C" int sdpsr_square(sdpsr_ctx* ctx, int method, int slices) {
  CTX_ENTER();
  SDPSR_REQUIRE(ctx->x_valid, SDPSR_E_STATE, "X is not defined yet (call sdpsr_fill or sdpsr_set_matrix first)");
  SDPSR_REQUIRE(method >= 0 && method <= 3, SDPSR_E_INVALID,
                "method must be 0 (DMMA), 1 (INT8), 2 (INT8, 7-bit digits) or 3 (INT8, 8-bit digits)");
  SDPSR_TRY(square_x(ctx, method ? 1 : 0, slices > 0 ? slices : ctx->i8_slices, method == 2 ? 7 : method == 3 ? 8 : 0,
                     nullptr));
  return finish(ctx);
}

extern "C" int sdpsr_set_square_slices(sdpsr_ctx* ctx, int slices) {
  CTX_ENTER();
  SDPSR_REQUIRE(slices == 0 || (slices >= 2 && slices <= 8), SDPSR_E_INVALID, "slices must be 0 (default) or in [2, 8]");
  ctx->i8_slices = slices;
  return SDPSR_OK;
}

extern "C" int sdpsr_product_round_refine(sdpsr_ctx* ctx, const double* rx, const double* ry, int64_t len,
                                          double atol, int64_t* dim) {
  CTX_ENTER();
  SDPSR_REQUIRE(len == ctx->dim, SDPSR_E_INVALID, "length(values) != dim(P) (src/partitions.jl:69)");
  double* Y = nullptr;
  SDPSR_TRY(matrix_ptr(ctx, SDPSR_MAT_W, &Y));
  SDPSR_TRY(sdpsr_upload_values(ctx, ry, len));
  SDPSR_TRY(sdpsr_build_lut(ctx, ctx->d_values, len));
  SDPSR_TRY(sdpsr_materialize_fill(ctx, Y));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));   // d_values is reused below
  SDPSR_TRY(sdpsr_upload_values(ctx, rx, len));
  SDPSR_TRY(sdpsr_build_lut(ctx, ctx->d_values, len));
  SDPSR_TRY(sdpsr_materialize_fill(ctx, ctx->X));
  ctx->x_valid = true;
  SDPSR_TRY(sdpsr_gemm_f64(ctx, ctx->X, ctx->ld, Y, ctx->ld, ctx->X2, ctx->ld, ctx->ld, ctx->n, ctx->n, false,
                           /*shard=*/true));
  SDPSR_TRY(sdpsr_generic_refine_values(ctx, ctx->X2, atol, true, nullptr, dim));
  return finish(ctx);
}

// ---------------------------------------------------------------------------
// timing / counters
// ---------------------------------------------------------------------------
extern "C" int sdpsr_timing_reset(sdpsr_ctx* ctx) {
  CTX_ENTER();
  drain_events(ctx);
  for (int i = 0; i < SDPSR_K_COUNT; ++i) {
    ctx->t_ms[i] = 0;
    ctx->t_launch[i] = 0;
    ctx->t_work[i] = 0;
  }
  return SDPSR_OK;
}

extern "C" int sdpsr_timing_get(sdpsr_ctx* ctx, int family, double* total_ms, int64_t* launches, double* work) {
  CTX_ENTER();
  SDPSR_REQUIRE(family >= 0 && family < SDPSR_K_COUNT, SDPSR_E_INVALID, "unknown kernel family");
  drain_events(ctx);
  if (total_ms) *total_ms = ctx->t_ms[family];
  if (launches) *launches = ctx->t_launch[family];
  if (work) *work = ctx->t_work[family];
  return SDPSR_OK;
}

extern "C" int sdpsr_set_stream(sdpsr_ctx* ctx, void* cuda_stream) {
  CTX_ENTER();
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  drain_events(ctx);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->own_stream = false;
  ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  if (ctx->solver) {           // cuSOLVER handle is bound lazily; rebind on next sdpsr_eig
    sdpsr_blockdiag_rebind(ctx);
  }
  return SDPSR_OK;
}

extern "C" int sdpsr_launch_count(sdpsr_ctx* ctx, int64_t* launches) {
  CTX_ENTER();
  if (launches) *launches = ctx->launches;
  return SDPSR_OK;
}
