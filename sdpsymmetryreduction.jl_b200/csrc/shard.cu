// shard.cu -- the partition itself sharded over the ranks of a communicator.
//
// Shards `Partition{T}(M)` + `refine!` + `__sort_unique!` (src/partitions.jl:24-35,44-66): with G ranks the
// label matrix is cut into G contiguous COLUMN BLOCKS (column-major storage: one contiguous range of the
// padded linear index per rank).  A refine pass then is
//   1. local   every rank runs the streaming pass (refine.cu) over ITS block only: N^2/G entries, a
//              rank-local key table  key -> (slot, smallest linear index inside the block);
//   2. merge   the <= dim local (key, first index) pairs of every rank are all-gathered (C2: a few KB) and
//              merged identically on every rank: min of the first indices per key, then the reference's
//              canonical numbering = rank of the first indices (first occurrence in column-major order);
//   3. relabel the block is rewritten from local slot ids to those canonical ids (8 B/entry on N^2/G).
// In this mode the labels ARE the canonical labels and tab[cur] is an identity table (slot i <-> class
// i+1) that every rank holds bit for bit; nothing observable depends on the rank count.
//   4. gather  (C1, lazy) only a consumer that needs other ranks' columns triggers the all-gather of the
//              blocks -- as 1- or 2-byte labels while dim allows it (the INT8 square gathers its digit
//              slices straight from those), expanded to u32 locally when a u32 consumer asks.
// The closure loop needs one such gather per iteration (before the square); the projection pass and both
// refine passes work on the local block alone.
#include <algorithm>
#include <cstring>

#include "sdpsr_internal.cuh"

namespace {

__device__ __forceinline__ uint64_t mix64s(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

struct __align__(16) PackedKey {
  uint64_t key;
  uint32_t minidx;
  uint32_t pad;
};

// buf[0] = (count, -) header, buf[1 + i] = (key, first index) of the i-th occupied slot of the local table (i < cap)
__global__ void pack_kernel(const uint32_t* __restrict__ occ, const KeySlot* __restrict__ slots, uint32_t count, uint32_t cap,
                            PackedKey* __restrict__ buf) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    PackedKey h;
    h.key = count;
    h.minidx = 0u;
    h.pad = 0u;
    buf[0] = h;
  }
  if (i >= count || i >= cap) return;
  const uint32_t s = occ[i];
  PackedKey p;
  p.key = slots[s].key;
  p.minidx = slots[s].minidx;
  p.pad = 0u;
  buf[1 + i] = p;
}

// insert every gathered pair into the merge table: key -> min over the ranks of the first index
__global__ void merge_insert_kernel(const PackedKey* __restrict__ buf, const uint32_t* __restrict__ counts, uint32_t maxc,
                                    int nranks, KeySlot* __restrict__ mslots,
                                    uint32_t* __restrict__ mocc, uint32_t* __restrict__ mmeta, uint32_t mask,
                                    uint32_t limit) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint64_t)maxc * nranks) return;
  const uint32_t r = (uint32_t)(t / maxc), i = (uint32_t)(t % maxc);
  if (i >= counts[r]) return;
  const PackedKey p = buf[(uint64_t)r * (maxc + 1) + 1 + i];       // (every rank's record starts with its header)
  uint32_t s = (uint32_t)(mix64s(p.key) >> 20) & mask;
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    const unsigned long long old =
        atomicCAS(reinterpret_cast<unsigned long long*>(&mslots[s].key), (unsigned long long)KEY_EMPTY, (unsigned long long)p.key);
    if (old == KEY_EMPTY) {
      const uint32_t pos = atomicAdd(mmeta, 1u);
      if (pos < limit)
        mocc[pos] = s;
      else
        mmeta[1] = 1u;
    }
    if (old == KEY_EMPTY || old == p.key) {
      atomicMin(&mslots[s].minidx, p.minidx);
      return;
    }
    s = (s + 1) & mask;
  }
  mmeta[1] = 1u;
}

__device__ __forceinline__ uint32_t merged_find(const KeySlot* __restrict__ mslots, uint32_t mask, uint64_t key) {
  uint32_t s = (uint32_t)(mix64s(key) >> 20) & mask;
  for (uint32_t probe = 0; probe <= mask; ++probe) {
    if (mslots[s].key == key) return s;
    s = (s + 1) & mask;
  }
  return 0u;   // unreachable: every local key was inserted
}

// l2g[local slot + 1] = canonical label of the local key
__global__ void local_to_global_kernel(const uint32_t* __restrict__ occ, const KeySlot* __restrict__ slots, uint32_t count,
                                       const KeySlot* __restrict__ mslots, const uint32_t* __restrict__ mrank,
                                       uint32_t mask, uint32_t* __restrict__ l2g) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) l2g[0] = 0u;
  if (i >= count) return;
  const uint32_t s = occ[i];
  l2g[s + 1] = mrank[merged_find(mslots, mask, slots[s].key) + 1];
}

// the identity table every rank ends up with: slot r-1 holds the key / first index of canonical class r
__global__ void identity_table_kernel(const uint32_t* __restrict__ mocc, const KeySlot* __restrict__ mslots,
                                      const uint32_t* __restrict__ mrank, uint32_t count, KeySlot* __restrict__ slots,
                                      uint32_t* __restrict__ occ, uint32_t* __restrict__ rank, uint32_t* __restrict__ meta) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    rank[0] = 0u;
    meta[0] = count;
    meta[1] = 0u;
  }
  if (i >= count) return;
  const uint32_t ms = mocc[i];
  const uint32_t r = mrank[ms + 1];          // 1 .. count
  slots[r - 1].key = mslots[ms].key;
  slots[r - 1].minidx = mslots[ms].minidx;
  occ[r - 1] = r - 1;
  rank[r] = r;
}

__global__ void relabel_block_kernel(uint32_t* __restrict__ lab, const uint32_t* __restrict__ l2g, uint64_t total) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
  for (uint64_t b = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; b < total; b += stride) {
    uint4 l = *reinterpret_cast<const uint4*>(lab + b);
    l.x = __ldg(l2g + l.x);
    l.y = __ldg(l2g + l.y);
    l.z = __ldg(l2g + l.z);
    l.w = __ldg(l2g + l.w);
    *reinterpret_cast<uint4*>(lab + b) = l;
  }
}

template <typename T>
__global__ void compact_kernel(const uint32_t* __restrict__ lab, T* __restrict__ out, uint64_t total) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
  for (uint64_t b = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; b < total; b += stride) {
    const uint4 l = __ldcs(reinterpret_cast<const uint4*>(lab + b));
    out[b] = (T)l.x;
    out[b + 1] = (T)l.y;
    out[b + 2] = (T)l.z;
    out[b + 3] = (T)l.w;
  }
}

template <typename T>
__global__ void expand_kernel(const T* __restrict__ in, uint32_t* __restrict__ lab, uint64_t total) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
  for (uint64_t b = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; b < total; b += stride)
    __stcs(reinterpret_cast<uint4*>(lab + b), make_uint4((uint32_t)in[b], (uint32_t)in[b + 1], (uint32_t)in[b + 2], (uint32_t)in[b + 3]));
}

int grid_for(sdpsr_ctx* ctx, uint64_t total) {
  return (int)std::max<uint64_t>(1, std::min<uint64_t>((total / 4 + 255) / 256, (uint64_t)ctx->sm_count * 8));
}

}  // namespace

// column block of rank r: columns [n*r/G, n*(r+1)/G)  ->  padded linear indices [begin, end)
void sdpsr_shard_block(const sdpsr_ctx* ctx, int r, uint64_t* begin, uint64_t* end) {
  const int G = ctx->nranks;
  const uint64_t c0 = (uint64_t)ctx->n * (uint64_t)r / (uint64_t)G, c1 = (uint64_t)ctx->n * (uint64_t)(r + 1) / (uint64_t)G;
  *begin = c0 * (uint64_t)ctx->ld;
  *end = c1 * (uint64_t)ctx->ld;
}

bool sdpsr_shard_active(const sdpsr_ctx* ctx) { return ctx->nranks > 1 && !(ctx->flags & SDPSR_F_REPLICATED_REFINE); }

// Steps 2 and 3 of a sharded refine pass.  `tloc` is the rank-local table of the pass that just ran over
// this rank's block of `lab` (= ctx->labels_alt, provisional local slot ids); on return `tloc` is the
// identity table of the merged partition and the block holds canonical labels.
int sdpsr_shard_merge(sdpsr_ctx* ctx, KeyTable& tloc, uint32_t* lab, int64_t* dim) {
  const int G = ctx->nranks;
  Timed tm(ctx, SDPSR_K_RANK, 0.0);
  // ---- C2: the packed (count | key, first index ...) records of every rank, ONE all-gather -----------
  // A record holds up to PK_CAP pairs (64 KB); only a pass that finds more classes than that on some rank
  // pays a second, exactly sized exchange.
  constexpr uint32_t PK_CAP = 4096;
  uint32_t* d_cnt = nullptr;
  SDPSR_TRY(sdpsr_scratch_t(ctx, 30, (size_t)sdpsr_ctx::MAX_RANKS, &d_cnt));
  uint32_t* h_cnt = reinterpret_cast<uint32_t*>(ctx->h_pinned) + 96;        // MAX_RANKS words
  PackedKey* h_hdr = reinterpret_cast<PackedKey*>(reinterpret_cast<unsigned char*>(ctx->h_pinned) + 2048);   // MAX_RANKS headers
  PackedKey* buf = nullptr;
  uint32_t maxc = PK_CAP;
  for (int attempt = 0;; ++attempt) {
    const size_t rec = (size_t)maxc + 1;
    SDPSR_TRY(sdpsr_scratch_t(ctx, 31, rec * (size_t)G, &buf));
    pack_kernel<<<(std::max<uint32_t>(1u, std::min(tloc.count, maxc)) + 255) / 256, 256, 0, ctx->stream>>>(
        tloc.occ, tloc.slots, tloc.count, maxc, buf + rec * ctx->rank);
    count_launch(ctx);
    SDPSR_TRY(sdpsr_comm_allgather(ctx, buf, rec * sizeof(PackedKey)));
    SDPSR_CUDA(cudaMemcpy2DAsync(h_hdr, sizeof(PackedKey), buf, rec * sizeof(PackedKey), sizeof(PackedKey), (size_t)G,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
    uint32_t need = 0;
    for (int r = 0; r < G; ++r) {
      h_cnt[r] = (uint32_t)h_hdr[r].key;
      need = std::max(need, h_cnt[r]);
    }
    if (need <= maxc) break;
    SDPSR_REQUIRE(attempt == 0, SDPSR_E_CUDA, "internal: key-table exchange did not converge");
    maxc = need;                                   // identical on every rank: all retry together
  }
  uint64_t total = 0;
  for (int r = 0; r < G; ++r) total += h_cnt[r];
  SDPSR_CUDA(cudaMemcpyAsync(d_cnt, h_cnt, sizeof(uint32_t) * (size_t)G, cudaMemcpyHostToDevice, ctx->stream));
  KeyTable& tm_ = ctx->tab_merge;
  if (total == 0) {       // the all-zero matrix: the empty partition
    SDPSR_CUDA(cudaMemsetAsync(tloc.meta, 0, 4 * sizeof(uint32_t), ctx->stream));
    SDPSR_CUDA(cudaMemsetAsync(tloc.rank, 0, sizeof(uint32_t), ctx->stream));
    tloc.count = 0;
    if (dim) *dim = 0;
    return SDPSR_OK;
  }
  // ---- identical merge on every rank ---------------------------------------------------------------
  const size_t mcap = std::max<size_t>(64, next_pow2(2 * total));
  SDPSR_TRY(sdpsr_table_alloc(ctx, tm_, mcap));
  SDPSR_CUDA(cudaMemsetAsync(tm_.slots, 0xff, (size_t)tm_.cap * sizeof(KeySlot), ctx->stream));
  SDPSR_CUDA(cudaMemsetAsync(tm_.meta, 0, 4 * sizeof(uint32_t), ctx->stream));
  {
    const uint64_t nthreads = (uint64_t)maxc * G;
    merge_insert_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, ctx->stream>>>(
        buf, d_cnt, maxc, G, tm_.slots, tm_.occ, tm_.meta, tm_.cap - 1, tm_.cap);
    count_launch(ctx);
  }
  uint32_t* hm = reinterpret_cast<uint32_t*>(ctx->h_pinned);
  SDPSR_CUDA(cudaMemcpyAsync(hm, tm_.meta, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  SDPSR_REQUIRE(hm[1] == 0u, SDPSR_E_CUDA, "internal: merge table overflow");
  tm_.count = hm[0];
  SDPSR_TRY(sdpsr_rank_table(ctx, tm_));          // canonical numbering: rank of the merged first indices
  // ---- local slot ids -> canonical labels ----------------------------------------------------------
  uint32_t* l2g = nullptr;
  SDPSR_TRY(sdpsr_scratch_t(ctx, 6, (size_t)tloc.cap + 1, &l2g));
  local_to_global_kernel<<<(std::max<uint32_t>(tloc.count, 1) + 255) / 256, 256, 0, ctx->stream>>>(
      tloc.occ, tloc.slots, tloc.count, tm_.slots, tm_.rank, tm_.cap - 1, l2g);
  count_launch(ctx);
  uint64_t b0, b1;
  sdpsr_shard_block(ctx, ctx->rank, &b0, &b1);
  if (b1 > b0) {
    relabel_block_kernel<<<grid_for(ctx, b1 - b0), 256, 0, ctx->stream>>>(lab + b0, l2g, b1 - b0);
    count_launch(ctx);
  }
  // ---- tloc becomes the identity table of the merged partition -------------------------------------
  const uint32_t dimg = tm_.count;
  const size_t need = std::max<size_t>(tloc.cap, next_pow2(2 * (uint64_t)dimg));
  // (l2g was built from tloc by kernels earlier on this stream; a reallocation synchronises by itself)
  SDPSR_TRY(sdpsr_table_alloc(ctx, tloc, need));
  SDPSR_CUDA(cudaMemsetAsync(tloc.slots, 0xff, (size_t)tloc.cap * sizeof(KeySlot), ctx->stream));
  identity_table_kernel<<<(dimg + 255) / 256, 256, 0, ctx->stream>>>(tm_.occ, tm_.slots, tm_.rank, dimg, tloc.slots,
                                                                      tloc.occ, tloc.rank, tloc.meta);
  count_launch(ctx);
  SDPSR_CUDA(cudaGetLastError());
  tloc.count = dimg;
  if (dim) *dim = dimg;
  return SDPSR_OK;
}

static int allgather_blocks(sdpsr_ctx* ctx, void* base, size_t elt_bytes) {
  size_t off[sdpsr_ctx::MAX_RANKS], len[sdpsr_ctx::MAX_RANKS];
  for (int r = 0; r < ctx->nranks; ++r) {
    uint64_t b0, b1;
    sdpsr_shard_block(ctx, r, &b0, &b1);
    off[r] = (size_t)b0 * elt_bytes;
    len[r] = (size_t)(b1 - b0) * elt_bytes;
  }
  return sdpsr_comm_allgatherv(ctx, base, off, len);
}

// C1: every rank's block of the label matrix as 1- / 2-byte canonical labels (4 bytes: the u32 labels
// themselves).  Collective; a no-op while the gathered copy is current.
int sdpsr_shard_gather_compact(sdpsr_ctx* ctx) {
  if (!sdpsr_shard_active(ctx) || ctx->labels_full || ctx->clabels_valid) return SDPSR_OK;
  const int w = ctx->dim < 256 ? 1 : ctx->dim < 65536 ? 2 : 4;
  uint64_t b0, b1;
  sdpsr_shard_block(ctx, ctx->rank, &b0, &b1);
  Timed tm(ctx, SDPSR_K_MISC, (double)ctx->elems * w);
  if (w == 4) {
    SDPSR_TRY(allgather_blocks(ctx, ctx->labels, 4));
    ctx->labels_full = true;
    ctx->clabel_width = 4;
    return SDPSR_OK;
  }
  if (!ctx->clabels) SDPSR_CUDA(cudaMalloc(&ctx->clabels, ctx->elems * 2));
  if (b1 > b0) {
    if (w == 1)
      compact_kernel<uint8_t><<<grid_for(ctx, b1 - b0), 256, 0, ctx->stream>>>(ctx->labels + b0, (uint8_t*)ctx->clabels + b0, b1 - b0);
    else
      compact_kernel<uint16_t><<<grid_for(ctx, b1 - b0), 256, 0, ctx->stream>>>(ctx->labels + b0, (uint16_t*)ctx->clabels + b0, b1 - b0);
    count_launch(ctx);
  }
  SDPSR_TRY(allgather_blocks(ctx, ctx->clabels, (size_t)w));
  ctx->clabel_width = w;
  ctx->clabels_valid = true;
  return SDPSR_OK;
}

// u32 labels of ALL columns on this rank (collective).  Consumers that read other ranks' columns through
// ctx->labels call this first; it costs one compact gather plus a local expansion.
int sdpsr_shard_ensure_full_labels(sdpsr_ctx* ctx) {
  if (!sdpsr_shard_active(ctx) || ctx->labels_full) return SDPSR_OK;
  SDPSR_TRY(sdpsr_shard_gather_compact(ctx));
  if (ctx->labels_full) return SDPSR_OK;
  for (int r = 0; r < ctx->nranks; ++r) {
    if (r == ctx->rank) continue;
    uint64_t b0, b1;
    sdpsr_shard_block(ctx, r, &b0, &b1);
    if (b1 == b0) continue;
    if (ctx->clabel_width == 1)
      expand_kernel<uint8_t><<<grid_for(ctx, b1 - b0), 256, 0, ctx->stream>>>((const uint8_t*)ctx->clabels + b0, ctx->labels + b0, b1 - b0);
    else
      expand_kernel<uint16_t><<<grid_for(ctx, b1 - b0), 256, 0, ctx->stream>>>((const uint16_t*)ctx->clabels + b0, ctx->labels + b0, b1 - b0);
    count_launch(ctx);
  }
  SDPSR_CUDA(cudaGetLastError());
  ctx->labels_full = true;
  return SDPSR_OK;
}
