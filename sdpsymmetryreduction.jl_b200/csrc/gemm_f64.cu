// gemm_f64.cu -- FP64 tensor-core GEMM for sm_100a:  C = A * B, column-major.
//
// Replaces OpenBLAS dgemm behind `mul!(X2, X, X)` (src/partitions.jl:172),
// `mul!(XY, X, Y)` (:212) and the products of src/eigen_decomposition.jl:70,203.
//
// Blackwell has no tcgen05 kind for f64; the FP64 tensor path is the warp-level
// `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4).  Structure:
//   * one producer warp feeds a STAGES-deep ring of shared-memory tiles with TMA
//     (cp.async.bulk.tensor.2d, FLOAT64 tensor maps, SWIZZLE_128B) signalled through
//     full/empty mbarriers;
//   * 8 consumer warps own a 64 x 32 sub-tile each of the 128 x 128 CTA tile and
//     issue DMMA from registers loaded with conflict-free LDS.64.
// Operand layouts (both are what column-major storage gives for free):
//   A tile: BM/16 boxes of [16 k rows][16 m] doubles (m contiguous, 128-byte rows)
//   B tile: [128 n rows][16 k] doubles (k contiguous, 128-byte rows)
// The 128-byte swizzle XORs the 16-byte chunk index with (row & 7).  Because the m
// and n positions of an MMA fragment may be any fixed permutation of the tile's rows
// and columns, the lanes read rows {0,1,8,9,2,3,10,11} (+4) of a box and columns
// {0,2,4,6,1,3,5,7} of an 8-column group; with that choice every half-warp LDS.64
// touches 16 distinct bank pairs (derivation in DESIGN.md "GEMM").
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include "sdpsr_internal.cuh"

namespace {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 16;
constexpr int STAGES = 4;
constexpr int A_BOX_BYTES = 16 * BK * 8;               // 2 KB
constexpr int A_STAGE_BYTES = BM * BK * 8;             // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 8;             // 16 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int RASTER_GROUP = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// lower==1: only tiles with tile_m >= tile_n are launched (C symmetric); the strictly lower
// tiles are mirrored into the upper triangle by mirror_kernel afterwards.
template <bool PEERS>
__global__ void __launch_bounds__(THREADS, 1)
gemm_f64_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                double* __restrict__ C, int64_t ldc, int M, int Nc, int K, int tiles_m, int tiles_n, int lower,
                const int2* __restrict__ tile_list, int accum, double* const* __restrict__ peers, int npeers) {
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;       // SWIZZLE_128B atoms are 1 KB
  const uint32_t bar_full = base + STAGES * STAGE_BYTES;
  const uint32_t bar_empty = bar_full + STAGES * 8;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates -----------------------------------------------------
  int tm, tn;
  if (tile_list) {               // multi-GPU: this rank's share of the tiles
    const int2 t = tile_list[blockIdx.x];
    tm = t.x;
    tn = t.y;
  } else if (!lower) {
    const int pid = blockIdx.x;
    const int per_group = RASTER_GROUP * tiles_m;
    const int g = pid / per_group;
    const int first_n = g * RASTER_GROUP;
    const int gw = min(RASTER_GROUP, tiles_n - first_n);
    const int r = pid - g * per_group;
    tn = first_n + r % gw;
    tm = r / gw;
  } else {
    // pid enumerates the lower triangle row by row: pid = tm*(tm+1)/2 + tn
    const int pid = blockIdx.x;
    int t = (int)((sqrtf(8.0f * (float)pid + 1.0f) - 1.0f) * 0.5f);
    while ((t + 1) * (t + 2) / 2 <= pid) ++t;
    while (t * (t + 1) / 2 > pid) --t;
    tm = t;
    tn = pid - t * (t + 1) / 2;
  }
  const int m0 = tm * BM;
  const int n0 = tn * BN;
  const int KB = (K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == CONSUMER_WARPS) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        const uint32_t full = bar_full + 8 * s;
        mbar_expect_tx(full, STAGE_BYTES);
        const uint32_t sa = base + s * STAGE_BYTES;
        const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
        for (int bx = 0; bx < BM / 16; ++bx) tma_load_2d(sa + bx * A_BOX_BYTES, &tmA, full, m0 + 16 * bx, kb * BK);
        tma_load_2d(sb, &tmB, full, kb * BK, n0);
      }
    }
    return;
  }

  // ===================== DMMA consumers =====================
  const int wm = warp & 1;       // 64-row half of the CTA tile
  const int wn = warp >> 1;      // 32-column quarter
  const int ms = lane >> 2;      // fragment row slot
  const int kk = lane & 3;       // fragment k slot
  // permuted positions inside a 16-row box / 8-column group
  const int mrow = (ms & 1) + ((ms >> 1) & 1) * 8 + (ms >> 2) * 2;   // + 4*h
  const int ncol = (ms & 3) * 2 + (ms >> 2);

  // A fragment byte offset inside a box for (h, ks): k = 4*ks + kk
  //   off = k*128 + (((m>>1) ^ (k&7)) << 4) + (m&1)*8
  // B fragment byte offset for (group g, ks): n = wn*32 + g*8 + ncol, chunk = 2*ks + (kk>>1)
  //   off = n*128 + ((chunk ^ (n&7)) << 4) + (kk&1)*8
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int kb = 0; kb < KB; ++kb) {
    const int s = kb % STAGES;
    const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
    mbar_wait(bar_full + 8 * s, ph);
    const uint32_t sa = base + s * STAGE_BYTES + (wm * 4) * A_BOX_BYTES;
    const uint32_t sb = base + s * STAGE_BYTES + A_STAGE_BYTES;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      const int k = 4 * ks + kk;
      double a[8], b[4];
#pragma unroll
      for (int bx = 0; bx < 4; ++bx)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int m = mrow + 4 * h;
          a[2 * bx + h] = lds_f64(sa + bx * A_BOX_BYTES + k * 128 + ((((m >> 1) ^ (k & 7))) << 4) + (m & 1) * 8);
        }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int n = wn * 32 + g * 8 + ncol;
        const int chunk = 2 * ks + (kk >> 1);
        b[g] = lds_f64(sb + n * 128 + ((chunk ^ (n & 7)) << 4) + (kk & 1) * 8);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * s);
  }

  // ---- epilogue: fragments -> global (each warp store covers whole 32-byte sectors) ----
  // C fragment: row slot ms, column slots 2*kk and 2*kk+1
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + wm * 64 + (i >> 1) * 16 + mrow + 4 * (i & 1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int nslot = 2 * kk + c;
        const int n = n0 + wn * 32 + j * 8 + (nslot & 3) * 2 + (nslot >> 2);
        if (m < M && n < Nc) {
          const int64_t off = (int64_t)m + ldc * (int64_t)n;
          if (PEERS) {
            // fused exchange: the tile goes straight into every rank's copy of C (own copy included)
            // over NVLink peer mappings, while the other CTAs are still computing
#pragma unroll 1
            for (int r = 0; r < npeers; ++r) peers[r][off] = acc[i][j][c];
          } else {
            double* dst = C + off;
            // accum: 0 -> C = A*B, +1 -> C += A*B, -1 -> C -= A*B (complex products from real GEMMs)
            *dst = accum == 0 ? acc[i][j][c] : (accum > 0 ? *dst + acc[i][j][c] : *dst - acc[i][j][c]);
          }
        }
      }
    }
  }
}

// C[j, i] = C[i, j] for i > j, tile by tile through shared memory (coalesced both ways)
__global__ void __launch_bounds__(256) mirror_lower_kernel(double* __restrict__ C, int64_t ldc, int n, int bi0) {
  __shared__ double t[32][33];
  const int bi = blockIdx.x + bi0, bj = blockIdx.y;
  if (bi < bj) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int i = bi * 32 + tx, j = bj * 32 + r;
    t[r][tx] = (i < n && j < n) ? C[(int64_t)i + ldc * j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    // write C[j', i'] with j' = bj*32 + tx (row), i' = bi*32 + r (column)
    const int jr = bj * 32 + tx, ic = bi * 32 + r;
    if (jr < n && ic < n && ic > jr) C[(int64_t)jr + ldc * ic] = t[tx][r];
  }
}

typedef CUresult (*PFN_tmapEncode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_tmapEncode get_encode() {
  static PFN_tmapEncode fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncode>(p);
  }
  return fn;
}

// 2-D FLOAT64 map over a column-major matrix: dim0 = rows (contiguous), dim1 = columns
int make_map(sdpsr_ctx* ctx, CUtensorMap* map, const double* ptr, int64_t rows, int64_t cols, int64_t ld,
             uint32_t box_rows, uint32_t box_cols) {
  PFN_tmapEncode enc = get_encode();
  SDPSR_REQUIRE(enc != nullptr, SDPSR_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {box_rows, box_cols};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SDPSR_REQUIRE(r == CUDA_SUCCESS, SDPSR_E_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  return SDPSR_OK;
}

}  // namespace

// Tiles of the (possibly lower-triangular) product owned by `rank` of `nranks`: tile-COLUMNS are dealt
// round-robin (tn % nranks == rank), which balances the triangle and keeps every owned slab contiguous
// in column-major storage.  Mirrored in sdpsymmetryreduction.jl_b200/sharding.py (CPU-tested).
static void owned_tiles(int tiles_m, int tiles_n, bool lower, int nranks, int rank, std::vector<int2>& out) {
  out.clear();
  for (int tn = rank; tn < tiles_n; tn += nranks)
    for (int tm = lower ? tn : 0; tm < tiles_m; ++tm) out.push_back(make_int2(tm, tn));
}

int sdpsr_tile_deal_i8(int64_t n, int nranks, int rank, int pair, std::vector<int2>& out);

/* The tile deal of the sharded products as the kernels use it, for `rank` of `nranks` (host-only: works without a
 * device).  kind 0: FP64 GEMM, all 128 x 128 tiles; 1: FP64 GEMM, lower triangle; 2: INT8 square (128 x 256 tiles);
 * 3: INT8 square on CTA pairs (256 x 256 tiles).  out receives (tm, tn) pairs, *count their number.            */
extern "C" int sdpsr_debug_tile_deal(int kind, int64_t n, int nranks, int rank, int32_t* out, int64_t cap, int64_t* count) {
  if (n < 1 || nranks < 1 || rank < 0 || rank >= nranks || kind < 0 || kind > 3 || !count) return SDPSR_E_INVALID;
  std::vector<int2> tiles;
  if (kind <= 1) {
    const int t = (int)((n + BM - 1) / BM);
    owned_tiles(t, t, kind == 1, nranks, rank, tiles);
  } else {
    sdpsr_tile_deal_i8(n, nranks, rank, kind == 3, tiles);
  }
  *count = (int64_t)tiles.size();
  if (out) {
    if (cap < (int64_t)tiles.size()) return SDPSR_E_INVALID;
    for (size_t i = 0; i < tiles.size(); ++i) {
      out[2 * i] = tiles[i].x;
      out[2 * i + 1] = tiles[i].y;
    }
  }
  return SDPSR_OK;
}

int sdpsr_gemm_f64(sdpsr_ctx* ctx, const double* A, int64_t lda, const double* B, int64_t ldb, double* C,
                   int64_t ldc, int64_t M, int64_t Nc, int64_t K, bool symmetric_out, bool shard, int accum) {
  SDPSR_REQUIRE(M > 0 && Nc > 0 && K > 0, SDPSR_E_INVALID, "empty GEMM");
  SDPSR_REQUIRE(lda % 2 == 0 && ldb % 2 == 0, SDPSR_E_INVALID, "leading dimensions must be even (16-byte TMA strides)");
  SDPSR_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0), SDPSR_E_INVALID, "operands must be 16-byte aligned");
  static bool attr_set_dev[64] = {false};      // the attribute is per device
  bool& attr_set = attr_set_dev[ctx->device & 63];
  if (!attr_set) {
    SDPSR_CUDA(cudaFuncSetAttribute(gemm_f64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    SDPSR_CUDA(cudaFuncSetAttribute(gemm_f64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap tmA, tmB;
  // A: M x K, boxes of 16 rows (m) x BK columns (k)
  SDPSR_TRY(make_map(ctx, &tmA, A, M, K, lda, 16, BK));
  // B: K x Nc, boxes of BK rows (k) x BN columns (n)
  SDPSR_TRY(make_map(ctx, &tmB, B, K, Nc, ldb, BK, BN));
  const int tiles_m = (int)((M + BM - 1) / BM);
  const int tiles_n = (int)((Nc + BN - 1) / BN);
  const bool lower = symmetric_out && accum == 0 && M >= Nc && tiles_m == tiles_n;
  int64_t ntiles = lower ? (int64_t)tiles_m * (tiles_m + 1) / 2 : (int64_t)tiles_m * tiles_n;
  const bool sharded = shard && ctx->nranks > 1;
  const int2* d_tiles = nullptr;
  if (sharded) {
    std::vector<int2> tiles;
    owned_tiles(tiles_m, tiles_n, lower, ctx->nranks, ctx->rank, tiles);
    ntiles = (int64_t)tiles.size();
    if (ctx->tile_alloc < tiles.size()) {
      cudaFree(ctx->d_tiles);
      ctx->d_tiles = nullptr;
      SDPSR_CUDA(cudaMalloc(&ctx->d_tiles, std::max<size_t>(tiles.size(), 1024) * sizeof(int2)));
      ctx->tile_alloc = std::max<size_t>(tiles.size(), 1024);
    }
    if (ntiles) SDPSR_CUDA(cudaMemcpyAsync(ctx->d_tiles, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));     // `tiles` dies at scope end
    d_tiles = reinterpret_cast<const int2*>(ctx->d_tiles);
  }
  double* const* peers = (sharded && accum == 0) ? sdpsr_comm_peer_table(ctx, C) : nullptr;
  if (peers) {
    // nobody may still be reading the previous contents of C on any rank when remote stores begin
    SDPSR_TRY(sdpsr_comm_barrier(ctx));
  }
  {
    // work = flops issued: the lower-triangle launch covers ntiles full 128x128 tiles
    Timed tm(ctx, SDPSR_K_GEMM, (lower || sharded) ? 2.0 * (double)ntiles * BM * BN * (double)K
                                                   : 2.0 * (double)M * (double)Nc * (double)K);
    if (ntiles) {
      if (peers)
        gemm_f64_kernel<true><<<(unsigned)ntiles, THREADS, SMEM_BYTES, ctx->stream>>>(
            tmA, tmB, C, ldc, (int)M, (int)Nc, (int)K, tiles_m, tiles_n, lower ? 1 : 0, d_tiles, accum, peers, ctx->nranks);
      else
        gemm_f64_kernel<false><<<(unsigned)ntiles, THREADS, SMEM_BYTES, ctx->stream>>>(
            tmA, tmB, C, ldc, (int)M, (int)Nc, (int)K, tiles_m, tiles_n, lower ? 1 : 0, d_tiles, accum, nullptr, 1);
      count_launch(ctx);
    }
  }
  SDPSR_CUDA(cudaGetLastError());
  if (sharded) {
    if (peers) {
      // the tiles were stored into every rank's C by the epilogues; all ranks' kernels must have
      // finished before anyone reads C
      SDPSR_TRY(sdpsr_comm_barrier(ctx));
    } else {
      // every rank receives the tile-columns it does not own (grouped NCCL broadcasts over NVLink)
      SDPSR_TRY(sdpsr_comm_exchange_tilecols(ctx, C, ldc, Nc, BN, tiles_n));
    }
  }
  if (lower) {
    if (C == ctx->X2) SDPSR_TRY(sdpsr_mirror_lower(ctx, C, ldc, Nc, ctx->mirror_col0, ctx->mirror_col1));
    else SDPSR_TRY(sdpsr_mirror_lower(ctx, C, ldc, Nc));
  }
  return SDPSR_OK;
}

// C[j, i] = C[i, j] for i > j, restricted to the destination columns [col_begin, col_end) (col_end < 0: all)
int sdpsr_mirror_lower(sdpsr_ctx* ctx, double* C, int64_t ldc, int64_t n, int64_t col_begin, int64_t col_end) {
  if (col_end < 0) col_end = n;
  if (col_end <= col_begin) return SDPSR_OK;
  const int bi0 = (int)(col_begin / 32), bi1 = (int)((col_end + 31) / 32);
  Timed tm(ctx, SDPSR_K_MISC, (double)n * (double)(col_end - col_begin) * 8.0);
  dim3 g((unsigned)(bi1 - bi0), (unsigned)((n + 31) / 32));
  mirror_lower_kernel<<<g, 256, 0, ctx->stream>>>(C, ldc, (int)n, bi0);
  count_launch(ctx);
  SDPSR_CUDA(cudaGetLastError());
  return SDPSR_OK;
}
