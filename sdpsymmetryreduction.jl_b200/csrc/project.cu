// project.cu -- the constraint matrix A, the projection onto its row space and the
// initial partition.
//
// Replaces `qr(A')` + project_colspace! (src/partitions.jl:124-126, src/utils.jl:59-69),
// the projection step of the loop (src/partitions.jl:160-164) and the two initial
// elements CL, X0 (src/partitions.jl:129-146, Krylov.craig at :137).
//
// proj(v) = A' (A A')^-1 A v.  Two facts make it cheap on the device:
//   * (A' c)[idx] depends only on the *column pattern* A[:, idx]; entries are tagged once
//     with a pattern id, and t[p] = sum_k A[k,p] c[k] (ascending k, separate multiply and
//     add -- the order of Julia's sparse mul!) is evaluated once per pattern on the host;
//   * A v needs only the stored non-zeros: a deterministic chunked gather-dot.
// The m x m Gram system is solved on the host (m <= a few hundred) with a rank-revealing pivoted
// Cholesky in extended precision: dependent constraint rows are dropped, as a rank-revealing QR would.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "sdpsr_internal.cuh"

namespace {

constexpr int CHUNK = 8192;   // non-zeros per reduction chunk

__device__ __forceinline__ uint64_t mix64d(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

// h[col] += mix(row, value bits): an order-independent 64-bit signature of column `col` of A
__global__ void __launch_bounds__(256) pattern_hash_kernel(const uint32_t* __restrict__ col,
                                                           const double* __restrict__ val,
                                                           const uint32_t* __restrict__ chunk_row,
                                                           const uint32_t* __restrict__ chunk_beg,
                                                           const uint32_t* __restrict__ chunk_end,
                                                           unsigned long long* __restrict__ h) {
  const uint32_t c = blockIdx.x;
  const uint64_t row = chunk_row[c];
  for (uint32_t i = chunk_beg[c] + threadIdx.x; i < chunk_end[c]; i += blockDim.x) {
    const uint64_t vb = (uint64_t)__double_as_longlong(val[i]);
    atomicAdd(h + col[i], (unsigned long long)mix64d((row + 1) * 0x9e3779b97f4a7c15ull ^ mix64d(vb)));
  }
}

__global__ void relabel_kernel(uint32_t* __restrict__ ids, const uint32_t* __restrict__ rank, uint64_t total) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
    ids[i] = rank[ids[i]];
}

// cnt[p] = number of matrix entries with pattern id p (padding rows excluded); persistent CTAs over
// columns with shared-memory bins, lanes of equal id aggregated first
constexpr int PC_BINS = 8192;
__global__ void __launch_bounds__(256) pattern_count_kernel(const uint32_t* __restrict__ pid, int64_t n, int64_t ld,
                                                            unsigned long long* __restrict__ cnt, int64_t nbins) {
  __shared__ uint32_t bins[PC_BINS];
  const bool smem = nbins <= PC_BINS;
  if (smem) {
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) bins[i] = 0u;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int64_t nround = (n + blockDim.x - 1) / blockDim.x * blockDim.x;     // warp-uniform trip count
  for (int64_t j = blockIdx.x; j < n; j += gridDim.x)
    for (int64_t i = threadIdx.x; i < nround; i += blockDim.x) {
      const bool valid = i < n;
      const uint32_t p = valid ? pid[i + ld * j] : 0xffffffffu;
      const unsigned mask = __match_any_sync(0xffffffffu, p);
      if (valid && (int)(__ffs(mask) - 1) == lane) {
        if (smem) {
          const uint32_t old = atomicAdd(&bins[p], (uint32_t)__popc(mask));
          if (old > 0xf0000000u) atomicAdd(cnt + p, (unsigned long long)atomicExch(&bins[p], 0u));
        } else {
          atomicAdd(cnt + p, (unsigned long long)__popc(mask));
        }
      }
    }
  if (smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += blockDim.x)
      if (bins[i]) atomicAdd(cnt + i, (unsigned long long)bins[i]);
  }
}

// partial[c] = sum over the chunk of val * x[col]; x = xs[col] or lut[labels[col]]
__global__ void __launch_bounds__(256) rowdot_kernel(const uint32_t* __restrict__ col, const double* __restrict__ val,
                                                     const uint32_t* __restrict__ chunk_beg,
                                                     const uint32_t* __restrict__ chunk_end,
                                                     const double* __restrict__ xs, const double* __restrict__ lut,
                                                     const uint32_t* __restrict__ labels,
                                                     double* __restrict__ partial) {
  __shared__ double ws[8];
  const uint32_t c = blockIdx.x;
  double s = 0.0;
  for (uint32_t i = chunk_beg[c] + threadIdx.x; i < chunk_end[c]; i += blockDim.x) {
    const uint32_t cc = col[i];
    const double x = xs ? xs[cc] : lut[labels[cc]];
    s += val[i] * x;
  }
  for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    partial[c] = t;
  }
}

// CSR ingestion on the device: the caller's column indices (int32 / int64, 0- or 1-based, unpadded linear index)
// become padded u32 indices; range, strict ascent inside a row and explicit zeros are flagged
// (flags[0]: out of range, flags[1]: not ascending, flags[2]: explicit zero present).
template <typename IT>
__global__ void __launch_bounds__(256) csr_ingest_kernel(const IT* __restrict__ colin, const double* __restrict__ val,
                                                         const uint32_t* __restrict__ chunk_row,
                                                         const uint32_t* __restrict__ chunk_beg,
                                                         const uint32_t* __restrict__ chunk_end,
                                                         const int64_t* __restrict__ rowptr, int64_t base, int64_t n,
                                                         int64_t ld, uint32_t* __restrict__ colout,
                                                         uint32_t* __restrict__ flags) {
  const uint32_t c = blockIdx.x;
  const int64_t rstart = rowptr[chunk_row[c]];
  for (uint32_t i = chunk_beg[c] + threadIdx.x; i < chunk_end[c]; i += blockDim.x) {
    const int64_t idx = (int64_t)colin[i] - base;
    if (idx < 0 || idx >= n * n) {
      flags[0] = 1u;
      colout[i] = 0u;
      continue;
    }
    if ((int64_t)i > rstart && (int64_t)colin[i - 1] - base >= idx) flags[1] = 1u;
    if (val[i] == 0.0) flags[2] = 1u;
    colout[i] = (uint32_t)((idx % n) + ld * (idx / n));
  }
}

// out[p * m + k] = A[k, rep[p]] (0 when the entry is not stored): binary search of the padded index in row k
__global__ void pattern_rows_kernel(const uint32_t* __restrict__ col, const double* __restrict__ val,
                                    const int64_t* __restrict__ rowptr, int64_t m, const uint32_t* __restrict__ rep,
                                    int64_t npat, double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npat * m) return;
  const int64_t p = t / m, k = t % m;
  const uint32_t want = rep[p];
  int64_t lo = rowptr[k], hi = rowptr[k + 1];
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (col[mid] < want) lo = mid + 1;
    else hi = mid;
  }
  out[t] = (lo < rowptr[k + 1] && col[lo] == want) ? val[lo] : 0.0;
}

// out[b * m + k] = first stored entry of row k whose padded index is >= bound[b]
__global__ void row_bounds_kernel(const uint32_t* __restrict__ col, const int64_t* __restrict__ rowptr, int64_t m,
                                  const unsigned long long* __restrict__ bound, int nb, int64_t* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)nb * m) return;
  const int64_t b = t / m, k = t % m;
  const unsigned long long want = bound[b];
  int64_t lo = rowptr[k], hi = rowptr[k + 1];
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((unsigned long long)col[mid] < want) lo = mid + 1;
    else hi = mid;
  }
  out[t] = lo;
}

__device__ __forceinline__ double snap_round(double c, double pw, int do_snap, double atol, double scale) {
  if (do_snap) c = __ddiv_rn(rint(__dmul_rn(c, pw)), pw);      // numpy.round(c, decimals)
  if (fabs(c) < atol) return 0.0;                                // _clamp_round!, src/utils.jl:34-53
  int n;
  const double x = frexp(c, &n);
  const long long q = __double2ll_rz(__dmul_rn(scale, x));
  return ldexp(__ddiv_rn((double)q, scale), n);
}

// mode 0: out = round(snap(src - t[pid]))   (CL before symmetrisation)
// mode 1: out = t[pid]                      (x0 = A' coef0)
// mode 2: out = round(snap(t[pid]))         (X0 = proj(sym(x0)))
__global__ void __launch_bounds__(256) init_elem_kernel(int mode, const double* __restrict__ src,
                                                        const uint32_t* __restrict__ pid,
                                                        const double* __restrict__ tpat, double* __restrict__ out,
                                                        uint64_t total, double pw, int do_snap, double atol,
                                                        double scale) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const double t = tpat[pid[i]];
    double v;
    if (mode == 0)
      v = snap_round(__dsub_rn(src[i], t), pw, do_snap, atol, scale);
    else if (mode == 1)
      v = t;
    else
      v = snap_round(t, pw, do_snap, atol, scale);
    out[i] = v;
  }
}

// out[i,j] = (in[i,j] + in[j,i]) / 2   (src/utils.jl:71-81), 32x32 tiles through shared memory
__global__ void __launch_bounds__(256) symmetrize_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                         int64_t n, int64_t ld) {
  __shared__ double t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t bi = blockIdx.x, bj = blockIdx.y;
  for (int r = ty; r < 32; r += 8) {   // transposed tile: rows bj*32.., columns bi*32..
    const int64_t i = bj * 32 + tx, j = bi * 32 + r;
    t[r][tx] = (i < n && j < n) ? in[i + ld * j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = bi * 32 + tx, j = bj * 32 + r;
    if (i < n && j < n) out[i + ld * j] = (in[i + ld * j] + t[tx][r]) / 2;
  }
}

// out[i,j] = (f(i,j) + f(j,i)) / 2 for the columns j in [col0, col1), f = the elementwise step of init_elem_kernel
// (mode 0: round(snap(src - t[pid])), mode 1: t[pid]) evaluated on the fly at both positions: the un-symmetrised
// element is never written, and a rank of a sharded run produces its own column block only.  `out` must not
// alias `src`.  32 x 32 tiles, the transposed operand goes through shared memory.
__global__ void __launch_bounds__(256) init_sym_kernel(int mode, const double* __restrict__ src,
                                                       const uint32_t* __restrict__ pid, const double* __restrict__ tpat,
                                                       double* __restrict__ out, int64_t n, int64_t ld, int64_t col0,
                                                       int64_t col1, double pw, int do_snap, double atol, double scale) {
  __shared__ double t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t bi = blockIdx.x;                       // row tile
  const int64_t jb = col0 / 32 * 32 + (int64_t)blockIdx.y * 32;   // first column of the column tile
  auto f = [&](int64_t i, int64_t j) -> double {
    const double tp = tpat[pid[i + ld * j]];
    return mode == 0 ? snap_round(__dsub_rn(src[i + ld * j], tp), pw, do_snap, atol, scale) : tp;
  };
  for (int r = ty; r < 32; r += 8) {   // transposed tile: rows jb.., columns bi*32..
    const int64_t i = jb + tx, j = bi * 32 + r;
    t[r][tx] = (i < n && j < n) ? f(i, j) : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = bi * 32 + tx, j = jb + r;
    if (i < n && j >= col0 && j < col1) out[i + ld * j] = (f(i, j) + t[tx][r]) / 2;
  }
}

__global__ void __launch_bounds__(256) symcheck_kernel(const uint32_t* __restrict__ lab, int64_t n, int64_t ld,
                                                       uint32_t* __restrict__ bad) {
  __shared__ uint32_t t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t bi = blockIdx.x, bj = blockIdx.y;
  if (bi < bj) return;
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = bj * 32 + tx, j = bi * 32 + r;
    t[r][tx] = (i < n && j < n) ? lab[i + ld * j] : 0u;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = bi * 32 + tx, j = bj * 32 + r;
    if (i < n && j < n && lab[i + ld * j] != t[tx][r]) *bad = 1u;
  }
}

// Reduced-SDP assembly (README.md:57-60, test/sd_problems.jl:32-37): sums over the classes of S.
constexpr int RB_MAX = 4096;   // classes binned in shared memory per CTA

// newA[k + m*(c-1)] += val for every stored entry of row k whose class is c
__global__ void __launch_bounds__(256) reduce_rows_kernel(const uint32_t* __restrict__ col, const double* __restrict__ val,
                                                          const uint32_t* __restrict__ chunk_row,
                                                          const uint32_t* __restrict__ chunk_beg,
                                                          const uint32_t* __restrict__ chunk_end,
                                                          const uint32_t* __restrict__ labels,
                                                          const uint32_t* __restrict__ rank, int64_t m, int64_t d,
                                                          double* __restrict__ newA) {
  extern __shared__ double bins[];
  const bool use_smem = d <= RB_MAX;
  if (use_smem) {
    for (int64_t i = threadIdx.x; i < d; i += blockDim.x) bins[i] = 0.0;
    __syncthreads();
  }
  const uint32_t c = blockIdx.x;
  const int64_t k = chunk_row[c];
  for (uint32_t i = chunk_beg[c] + threadIdx.x; i < chunk_end[c]; i += blockDim.x) {
    const uint32_t cls = rank[labels[col[i]]];
    if (cls == 0u) continue;
    if (use_smem)
      atomicAdd(&bins[cls - 1], val[i]);
    else
      atomicAdd(&newA[k + m * (int64_t)(cls - 1)], val[i]);
  }
  if (use_smem) {
    __syncthreads();
    for (int64_t i = threadIdx.x; i < d; i += blockDim.x)
      if (bins[i] != 0.0) atomicAdd(&newA[k + m * i], bins[i]);
  }
}

// newC[c-1] += C[idx] for every entry of class c
__global__ void __launch_bounds__(256) reduce_objective_kernel(const double* __restrict__ Cv,
                                                               const uint32_t* __restrict__ labels,
                                                               const uint32_t* __restrict__ rank, uint64_t total,
                                                               int64_t d, double* __restrict__ newC) {
  extern __shared__ double bins[];
  const bool use_smem = d <= RB_MAX;
  if (use_smem) {
    for (int64_t i = threadIdx.x; i < d; i += blockDim.x) bins[i] = 0.0;
    __syncthreads();
  }
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t cls = rank[labels[i]];
    if (cls == 0u) continue;
    const double v = Cv[i];
    if (v == 0.0) continue;
    if (use_smem)
      atomicAdd(&bins[cls - 1], v);
    else
      atomicAdd(&newC[cls - 1], v);
  }
  if (use_smem) {
    __syncthreads();
    for (int64_t i = threadIdx.x; i < d; i += blockDim.x)
      if (bins[i] != 0.0) atomicAdd(&newC[i], bins[i]);
  }
}

__global__ void __launch_bounds__(256) symcheck_f64_kernel(const double* __restrict__ x, int64_t n, int64_t ld,
                                                           uint32_t* __restrict__ bad) {
  __shared__ double t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t bi = blockIdx.x, bj = blockIdx.y;
  if (bi < bj) return;
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = bj * 32 + tx, j = bi * 32 + r;
    t[r][tx] = (i < n && j < n) ? x[i + ld * j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = bi * 32 + tx, j = bj * 32 + r;
    if (i < n && j < n &&
        __double_as_longlong(x[i + ld * j]) != __double_as_longlong(t[tx][r]))
      *bad = 1u;
  }
}

}  // namespace

// 1 iff the device matrix is bit-for-bit symmetric (decides the half-GEMM for X*X)
int sdpsr_matrix_symmetric(sdpsr_ctx* ctx, const double* x, int* is_sym) {
  uint32_t* bad = ctx->d_scalars + 9;
  SDPSR_CUDA(cudaMemsetAsync(bad, 0, sizeof(uint32_t), ctx->stream));
  const unsigned nb = (unsigned)((ctx->n + 31) / 32);
  {
    Timed tm(ctx, SDPSR_K_MISC, (double)ctx->elems * 8.0);
    symcheck_f64_kernel<<<dim3(nb, nb), 256, 0, ctx->stream>>>(x, ctx->n, ctx->ld, bad);
    count_launch(ctx);
  }
  SDPSR_CUDA(cudaGetLastError());
  uint32_t* hb = reinterpret_cast<uint32_t*>(ctx->h_pinned) + 33;
  SDPSR_CUDA(cudaMemcpyAsync(hb, bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  *is_sym = (*hb == 0u) ? 1 : 0;
  return SDPSR_OK;
}

// ---------------------------------------------------------------------------
void sdpsr_constraints_free(sdpsr_ctx* ctx) {
  ctx->cons = ConstraintSet();      // device arrays live in scratch slots 8-14 (freed with the context)
}

// 1 iff the N x N u32 array (padded layout) equals its transpose
static int u32_symmetric(sdpsr_ctx* ctx, const uint32_t* arr, int* is_sym) {
  uint32_t* bad = ctx->d_scalars + 8;
  SDPSR_CUDA(cudaMemsetAsync(bad, 0, sizeof(uint32_t), ctx->stream));
  const unsigned nb = (unsigned)((ctx->n + 31) / 32);
  {
    Timed tm(ctx, SDPSR_K_MISC, (double)ctx->elems * 4.0);
    symcheck_kernel<<<dim3(nb, nb), 256, 0, ctx->stream>>>(arr, ctx->n, ctx->ld, bad);
    count_launch(ctx);
  }
  SDPSR_CUDA(cudaGetLastError());
  uint32_t* hb = reinterpret_cast<uint32_t*>(ctx->h_pinned) + 32;
  SDPSR_CUDA(cudaMemcpyAsync(hb, bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  *is_sym = (*hb == 0u) ? 1 : 0;
  return SDPSR_OK;
}

// Is the partition transpose-invariant?  Provisional ids are a relabelling of the classes, so the test runs
// on them directly.  The answer is cached: a refine by values that are symmetric by construction keeps it
// (sdpsr_refine_pass, RefineSpec::keeps_symmetry); anything else resets it to "unknown".
int sdpsr_symmetric_check(sdpsr_ctx* ctx, int* is_sym) {
  if (ctx->sym_state < 0) {
    int s = 0;
    SDPSR_TRY(sdpsr_shard_ensure_full_labels(ctx));
    SDPSR_TRY(u32_symmetric(ctx, ctx->labels, &s));
    ctx->sym_state = s;
  }
  *is_sym = ctx->sym_state;
  return SDPSR_OK;
}

// Rank-revealing factorisation of the Gram matrix G = A A' (symmetric positive semidefinite): Cholesky
// with diagonal pivoting in extended precision.  The reference projects with `qr(A')`
// (src/partitions.jl:124, src/utils.jl:62-66), which for sparse A is SuiteSparse's rank-revealing QR; a
// constraint row that depends (numerically) on the rows kept so far shows up here as a negligible pivot
// of the Schur complement and is dropped: the projector onto the row space -- and the minimum-norm
// solution of a consistent A x = b -- do not change when such a row is left out.  Extended precision
// (64-bit mantissa) keeps the solve of the squared-condition system at the accuracy a QR of A' has in
// double: the error of the PROJECTION A' G^-1 (A v) is governed by cond(A), not cond(A)^2, once the
// solve itself is exact to working precision.
static void gram_factor(ConstraintSet& c, const std::vector<long double>& G, int m) {
  std::vector<long double> a(G);
  c.gram_chol.assign((size_t)m * m, 0.0L);
  c.gram_perm.resize((size_t)m);
  for (int i = 0; i < m; ++i) c.gram_perm[(size_t)i] = i;
  long double dmax = 0.0L;
  for (int i = 0; i < m; ++i) dmax = std::max(dmax, a[(size_t)i * m + i]);
  const long double tol = 1e-13L * dmax;
  int rank = 0;
  std::vector<long double>& L = c.gram_chol;
  for (int k = 0; k < m; ++k) {
    int p = k;
    for (int i = k + 1; i < m; ++i)
      if (a[(size_t)i * m + i] > a[(size_t)p * m + p]) p = i;
    if (!(a[(size_t)p * m + p] > tol)) break;            // the remaining rows depend on the ones kept
    if (p != k) {                                         // symmetric interchange of k and p
      for (int j = 0; j < m; ++j) std::swap(a[(size_t)k * m + j], a[(size_t)p * m + j]);
      for (int i = 0; i < m; ++i) std::swap(a[(size_t)i * m + k], a[(size_t)i * m + p]);
      for (int j = 0; j < k; ++j) std::swap(L[(size_t)k * m + j], L[(size_t)p * m + j]);
      std::swap(c.gram_perm[(size_t)k], c.gram_perm[(size_t)p]);
    }
    const long double d = std::sqrt(a[(size_t)k * m + k]);
    L[(size_t)k * m + k] = d;
    for (int i = k + 1; i < m; ++i) L[(size_t)i * m + k] = a[(size_t)i * m + k] / d;
    for (int i = k + 1; i < m; ++i)
      for (int j = k + 1; j <= i; ++j) {
        a[(size_t)i * m + j] -= L[(size_t)i * m + k] * L[(size_t)j * m + k];
        a[(size_t)j * m + i] = a[(size_t)i * m + j];
      }
    rank = k + 1;
  }
  c.gram_rank = rank;
}

// x <- G^+ x restricted to the retained rows (dropped rows get coefficient 0)
int sdpsr_solve_gram(sdpsr_ctx* ctx, std::vector<double>& x) {
  ConstraintSet& c = ctx->cons;
  const int m = (int)c.m, r = c.gram_rank;
  const std::vector<long double>& L = c.gram_chol;
  std::vector<long double> y((size_t)r);
  for (int i = 0; i < r; ++i) {
    long double s = (long double)x[(size_t)c.gram_perm[(size_t)i]];
    for (int j = 0; j < i; ++j) s -= L[(size_t)i * m + j] * y[(size_t)j];
    y[(size_t)i] = s / L[(size_t)i * m + i];
  }
  for (int i = r - 1; i >= 0; --i) {
    long double s = y[(size_t)i];
    for (int j = i + 1; j < r; ++j) s -= L[(size_t)j * m + i] * y[(size_t)j];
    y[(size_t)i] = s / L[(size_t)i * m + i];
  }
  std::fill(x.begin(), x.end(), 0.0);
  for (int i = 0; i < r; ++i) x[(size_t)c.gram_perm[(size_t)i]] = (double)y[(size_t)i];
  return SDPSR_OK;
}

int sdpsr_upload_tpat(sdpsr_ctx* ctx, const std::vector<double>& coef) {
  ConstraintSet& c = ctx->cons;
  std::vector<double> t((size_t)c.npat + 1, 0.0);
  for (int64_t p = 1; p <= c.npat; ++p) {
    double s = 0.0;
    for (int64_t e = c.pat_ptr[p]; e < c.pat_ptr[p + 1]; ++e) {
      volatile double prod = c.pat_val[e] * coef[c.pat_row[e]];   // separate multiply and add
      s = s + prod;
    }
    t[p] = s;
  }
  SDPSR_CUDA(cudaMemcpyAsync(c.d_tpat, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));   // `t` dies at return
  return SDPSR_OK;
}

// C3: A x for x = lut[labels] with the partition sharded by column blocks.  Every rank reduces the stored
// non-zeros that fall into ITS block (per constraint row a contiguous sub-range: column indices ascend),
// the per-chunk partial sums are all-gathered and added in (rank, chunk) order on every rank: identical
// coefficients everywhere, no label of another rank's block is read.
static int rowdots_sharded_setup(sdpsr_ctx* ctx) {
  ConstraintSet& c = ctx->cons;
  if (c.sh_nranks == ctx->nranks && c.sh_rank == ctx->rank) return SDPSR_OK;
  const int G = ctx->nranks;
  const int64_t n = ctx->n;
  c.sh_chunk_row.assign((size_t)G, {});
  std::vector<uint32_t> beg, end;
  // first stored entry of every row at or behind each block boundary (padded linear index c0 * ld)
  std::vector<unsigned long long> hb((size_t)G + 1);
  for (int r = 0; r <= G; ++r) hb[(size_t)r] = (unsigned long long)(n * r / G) * (unsigned long long)ctx->ld;
  unsigned long long* d_b = nullptr;
  int64_t* d_o = nullptr;
  SDPSR_TRY(sdpsr_scratch_t(ctx, 0, (size_t)G + 1, &d_b));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 1, (size_t)(G + 1) * (size_t)c.m, &d_o));
  SDPSR_CUDA(cudaMemcpyAsync(d_b, hb.data(), hb.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  row_bounds_kernel<<<(unsigned)(((int64_t)(G + 1) * c.m + 255) / 256), 256, 0, ctx->stream>>>(c.d_col, c.d_rowptr, c.m, d_b, G + 1, d_o);
  count_launch(ctx);
  std::vector<int64_t> ho((size_t)(G + 1) * (size_t)c.m);
  SDPSR_CUDA(cudaMemcpyAsync(ho.data(), d_o, ho.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < G; ++r) {
    for (int64_t k = 0; k < c.m; ++k) {
      const int64_t e0 = ho[(size_t)r * (size_t)c.m + (size_t)k], e1 = ho[(size_t)(r + 1) * (size_t)c.m + (size_t)k];
      for (int64_t x = e0; x < e1; x += CHUNK) {
        c.sh_chunk_row[(size_t)r].push_back((uint32_t)k);
        if (r == ctx->rank) {
          beg.push_back((uint32_t)x);
          end.push_back((uint32_t)std::min<int64_t>(x + CHUNK, e1));
        }
      }
    }
  }
  c.sh_maxchunks = 1;
  for (int r = 0; r < G; ++r) c.sh_maxchunks = std::max(c.sh_maxchunks, c.sh_chunk_row[(size_t)r].size());
  const size_t mine = beg.size();
  SDPSR_TRY(sdpsr_scratch_t(ctx, 15, std::max<size_t>(1, 2 * mine), &c.d_sh_beg));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 7, c.sh_maxchunks * (size_t)G, &c.d_sh_partial));
  if (mine) {
    SDPSR_CUDA(cudaMemcpyAsync(c.d_sh_beg, beg.data(), mine * 4, cudaMemcpyHostToDevice, ctx->stream));
    SDPSR_CUDA(cudaMemcpyAsync(c.d_sh_beg + mine, end.data(), mine * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  c.sh_nranks = G;
  c.sh_rank = ctx->rank;
  return SDPSR_OK;
}

static int rowdots_sharded(sdpsr_ctx* ctx, const double* x_array, const double* lut, std::vector<double>& out) {
  ConstraintSet& c = ctx->cons;
  SDPSR_TRY(rowdots_sharded_setup(ctx));
  const int G = ctx->nranks;
  const size_t mine = c.sh_chunk_row[(size_t)ctx->rank].size();
  out.assign((size_t)c.m, 0.0);
  if (mine) {
    Timed tm(ctx, SDPSR_K_PROJECT, (double)c.nnz / G * 20.0);
    rowdot_kernel<<<(unsigned)mine, 256, 0, ctx->stream>>>(c.d_col, c.d_val, c.d_sh_beg, c.d_sh_beg + mine, x_array, lut,
                                                           ctx->labels, c.d_sh_partial + c.sh_maxchunks * (size_t)ctx->rank);
    count_launch(ctx);
  }
  SDPSR_CUDA(cudaGetLastError());
  SDPSR_TRY(sdpsr_comm_allgather(ctx, c.d_sh_partial, c.sh_maxchunks * sizeof(double)));
  std::vector<double> part(c.sh_maxchunks * (size_t)G);
  SDPSR_CUDA(cudaMemcpyAsync(part.data(), c.d_sh_partial, part.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < G; ++r)
    for (size_t i = 0; i < c.sh_chunk_row[(size_t)r].size(); ++i)
      out[c.sh_chunk_row[(size_t)r][i]] += part[c.sh_maxchunks * (size_t)r + i];      // fixed order: deterministic
  return SDPSR_OK;
}

// x_own_block: x_array is current in this rank's column block only (sharded run)
int sdpsr_rowdots(sdpsr_ctx* ctx, const double* x_array, const double* lut, std::vector<double>& out, bool x_own_block) {
  ConstraintSet& c = ctx->cons;
  if (sdpsr_shard_active(ctx) && (x_array ? x_own_block : !ctx->labels_full)) return rowdots_sharded(ctx, x_array, lut, out);
  out.assign((size_t)c.m, 0.0);
  if (c.nchunks == 0) return SDPSR_OK;
  {
    Timed tm(ctx, SDPSR_K_PROJECT, (double)c.nnz * 20.0);
    rowdot_kernel<<<(unsigned)c.nchunks, 256, 0, ctx->stream>>>(c.d_col, c.d_val, c.d_chunk_beg,
                                                                c.d_chunk_beg + c.nchunks, x_array, lut, ctx->labels,
                                                                c.d_partial);
    count_launch(ctx);
  }
  SDPSR_CUDA(cudaGetLastError());
  std::vector<double> part((size_t)c.nchunks);
  SDPSR_CUDA(cudaMemcpyAsync(part.data(), c.d_partial, part.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int64_t i = 0; i < c.nchunks; ++i) out[c.chunk_row[i]] += part[i];   // fixed order: deterministic
  return SDPSR_OK;
}

// Reduction chunks of the CSR rows (host rowptr only) + the device buffers of the constraint set.
static int constraints_layout(sdpsr_ctx* ctx, std::vector<uint32_t>& cbeg, std::vector<uint32_t>& cend) {
  ConstraintSet& c = ctx->cons;
  const int64_t m = c.m;
  const int64_t nnz = c.h_rowptr[m];
  c.nnz = nnz;
  SDPSR_REQUIRE(nnz < 0xffffffffll, SDPSR_E_UNSUPPORTED, "more than 2^32-1 stored constraint entries");
  c.chunk_row.clear();
  for (int64_t k = 0; k < m; ++k)
    for (int64_t e = c.h_rowptr[k]; e < c.h_rowptr[k + 1]; e += CHUNK) {
      c.chunk_row.push_back((uint32_t)k);
      cbeg.push_back((uint32_t)e);
      cend.push_back((uint32_t)std::min<int64_t>(e + CHUNK, c.h_rowptr[k + 1]));
    }
  c.nchunks = (int64_t)cbeg.size();
  const size_t nz_alloc = std::max<size_t>((size_t)nnz, 1), ch_alloc = std::max<size_t>((size_t)c.nchunks, 1);
  SDPSR_TRY(sdpsr_scratch_t(ctx, 8, nz_alloc, &c.d_col));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 9, nz_alloc, &c.d_val));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 10, ch_alloc, &c.d_chunk_row));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 11, 2 * ch_alloc, &c.d_chunk_beg));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 12, ch_alloc, &c.d_partial));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 13, ctx->elems, &c.d_pid));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 39, (size_t)m + 1, &c.d_rowptr));
  SDPSR_CUDA(cudaMemcpyAsync(c.d_rowptr, c.h_rowptr.data(), ((size_t)m + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (c.nchunks) {
    SDPSR_CUDA(cudaMemcpyAsync(c.d_chunk_row, c.chunk_row.data(), (size_t)c.nchunks * 4, cudaMemcpyHostToDevice, ctx->stream));
    SDPSR_CUDA(cudaMemcpyAsync(c.d_chunk_beg, cbeg.data(), (size_t)c.nchunks * 4, cudaMemcpyHostToDevice, ctx->stream));
    SDPSR_CUDA(cudaMemcpyAsync(c.d_chunk_beg + c.nchunks, cend.data(), (size_t)c.nchunks * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  return SDPSR_OK;
}

// Host ingestion (dense / CSC input, or a CSR with explicit zeros): h_rowptr is set, hcol holds unpadded linear
// indices.  Validates, pads and uploads.
static int constraints_from_host(sdpsr_ctx* ctx, const std::vector<int64_t>& hcol, const std::vector<double>& hval) {
  ConstraintSet& c = ctx->cons;
  const int64_t m = c.m, n = ctx->n, ld = ctx->ld;
  std::vector<uint32_t> cbeg, cend;
  SDPSR_TRY(constraints_layout(ctx, cbeg, cend));
  const int64_t nnz = c.nnz;
  std::vector<uint32_t> col((size_t)nnz);
  for (int64_t k = 0; k < m; ++k) {
    int64_t prev = -1;
    for (int64_t e = c.h_rowptr[k]; e < c.h_rowptr[k + 1]; ++e) {
      const int64_t idx = hcol[(size_t)e];
      SDPSR_REQUIRE(idx >= 0 && idx < n * n, SDPSR_E_INVALID, "constraint column index out of range");
      SDPSR_REQUIRE(idx > prev, SDPSR_E_INVALID, "constraint rows must have strictly increasing column indices");
      prev = idx;
      col[(size_t)e] = (uint32_t)((idx % n) + ld * (idx / n));
    }
  }
  if (nnz) {
    SDPSR_CUDA(cudaMemcpyAsync(c.d_col, col.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
    SDPSR_CUDA(cudaMemcpyAsync(c.d_val, hval.data(), (size_t)nnz * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));   // the host vectors die with the caller
  return sdpsr_constraints_finalize(ctx);
}

// Device ingestion of a caller-owned CSR (the large-problem path: 46 M stored entries for K(20,5)): the raw
// arrays are uploaded as they are and converted by one kernel; no host copy, no host pass over the non-zeros.
// *fallback = 1 when the matrix holds explicit zeros (the host path filters them).
template <typename IT>
static int constraints_from_csr(sdpsr_ctx* ctx, const int64_t* rowptr, const IT* colidx, const double* vals, int index_base,
                                int* fallback) {
  ConstraintSet& c = ctx->cons;
  const int64_t m = c.m;
  *fallback = 0;
  c.h_rowptr.assign((size_t)m + 1, 0);
  for (int64_t k = 0; k < m; ++k) {
    SDPSR_REQUIRE(rowptr[k + 1] >= rowptr[k], SDPSR_E_INVALID, "rowptr must be non-decreasing");
    c.h_rowptr[(size_t)k + 1] = rowptr[k + 1] - rowptr[0];
  }
  std::vector<uint32_t> cbeg, cend;
  SDPSR_TRY(constraints_layout(ctx, cbeg, cend));
  const int64_t nnz = c.nnz;
  if (nnz) {
    IT* raw = nullptr;
    SDPSR_TRY(sdpsr_scratch_t(ctx, 40, (size_t)nnz, &raw));
    uint32_t* flags = ctx->d_scalars + 12;
    SDPSR_CUDA(cudaMemsetAsync(flags, 0, 3 * sizeof(uint32_t), ctx->stream));
    SDPSR_CUDA(cudaMemcpyAsync(raw, colidx, (size_t)nnz * sizeof(IT), cudaMemcpyDefault, ctx->stream));
    SDPSR_CUDA(cudaMemcpyAsync(c.d_val, vals, (size_t)nnz * sizeof(double), cudaMemcpyDefault, ctx->stream));
    csr_ingest_kernel<IT><<<(unsigned)c.nchunks, 256, 0, ctx->stream>>>(raw, c.d_val, c.d_chunk_row, c.d_chunk_beg,
                                                                        c.d_chunk_beg + c.nchunks, c.d_rowptr, index_base,
                                                                        ctx->n, ctx->ld, c.d_col, flags);
    count_launch(ctx);
    SDPSR_CUDA(cudaGetLastError());
    uint32_t* hf = reinterpret_cast<uint32_t*>(ctx->h_pinned) + 112;
    SDPSR_CUDA(cudaMemcpyAsync(hf, flags, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
    SDPSR_REQUIRE(hf[0] == 0u, SDPSR_E_INVALID, "constraint column index out of range");
    SDPSR_REQUIRE(hf[1] == 0u, SDPSR_E_INVALID, "constraint rows must have strictly increasing column indices");
    if (hf[2]) {
      *fallback = 1;
      return SDPSR_OK;
    }
  }
  return sdpsr_constraints_finalize(ctx);
}

// Everything derived from the device CSR (d_col / d_val / d_rowptr + the chunk lists): pattern ids, the pattern
// table, the Gram factorisation.
int sdpsr_constraints_finalize(sdpsr_ctx* ctx) {
  ConstraintSet& c = ctx->cons;
  const int64_t m = c.m, n = ctx->n, ld = ctx->ld;
  const int64_t nnz = c.nnz;
  // ---- pattern ids: hash every column, first-occurrence rank of the hashes ------------
  unsigned long long* h = reinterpret_cast<unsigned long long*>(ctx->X2);
  SDPSR_CUDA(cudaMemsetAsync(h, 0, ctx->elems * 8, ctx->stream));
  if (c.nchunks) {
    pattern_hash_kernel<<<(unsigned)c.nchunks, 256, 0, ctx->stream>>>(c.d_col, c.d_val, c.d_chunk_row, c.d_chunk_beg,
                                                                      c.d_chunk_beg + c.nchunks, h);
    count_launch(ctx);
    SDPSR_CUDA(cudaGetLastError());
  }
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));   // host vectors above die below
  KeyTable& scratch = ctx->tab_scratch;   // reused across calls, freed with the context
  RefineSpec sp;
  sp.mode = KM_RAW;
  sp.vals = ctx->X2;
  sp.do_round = false;
  sp.raw_bits = true;
  sp.ignore_labels = true;
  sp.out_override = c.d_pid;
  sp.table_override = &scratch;
  int64_t npat = 0;
  int st = sdpsr_refine_pass(ctx, sp, &npat);
  if (st != SDPSR_OK) {
    return st;
  }
  c.npat = npat;
  {
    const int grid = (int)std::min<uint64_t>((ctx->elems + 255) / 256, (uint64_t)ctx->sm_count * 16);
    relabel_kernel<<<grid, 256, 0, ctx->stream>>>(c.d_pid, scratch.rank, ctx->elems);
    count_launch(ctx);
  }
  // representative entry of every pattern (its first occurrence) and pattern sizes
  std::vector<uint32_t> occ((size_t)npat), mi((size_t)npat), rk((size_t)scratch.cap + 1);
  if (npat) {
    SDPSR_CUDA(cudaMemcpyAsync(occ.data(), scratch.occ, (size_t)npat * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SDPSR_CUDA(cudaMemcpyAsync(rk.data(), scratch.rank, rk.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  std::vector<KeySlot> allslots((size_t)scratch.cap);
  SDPSR_CUDA(cudaMemcpyAsync(allslots.data(), scratch.slots, allslots.size() * sizeof(KeySlot), cudaMemcpyDeviceToHost, ctx->stream));
  unsigned long long* dcnt = nullptr;
  SDPSR_TRY(sdpsr_scratch_t(ctx, 0, (size_t)npat + 1, &dcnt));
  SDPSR_CUDA(cudaMemsetAsync(dcnt, 0, ((size_t)npat + 1) * 8, ctx->stream));
  pattern_count_kernel<<<(unsigned)std::min<int64_t>(n, (int64_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(
      c.d_pid, n, ld, dcnt, npat + 1);
  count_launch(ctx);
  std::vector<unsigned long long> cnt((size_t)npat + 1);
  SDPSR_CUDA(cudaMemcpyAsync(cnt.data(), dcnt, cnt.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<uint32_t> rep((size_t)npat + 1, 0xffffffffu);      // padded linear index of the representative
  for (int64_t i = 0; i < npat; ++i) {
    const uint32_t slot = occ[(size_t)i];
    rep[rk[(size_t)slot + 1]] = allslots[slot].minidx;
  }
  c.pat_cnt.assign(cnt.begin(), cnt.end());
  // ---- pattern table: column of A at each representative entry (binary searches on the device) ----------
  c.pat_ptr.assign((size_t)npat + 2, 0);
  c.pat_row.clear();
  c.pat_val.clear();
  std::vector<double> prow((size_t)npat * (size_t)m);
  if (npat) {
    uint32_t* d_rep = nullptr;
    double* d_prow = nullptr;
    SDPSR_TRY(sdpsr_scratch_t(ctx, 1, (size_t)npat, &d_rep));
    SDPSR_TRY(sdpsr_scratch_t(ctx, 2, (size_t)npat * (size_t)m, &d_prow));
    for (int64_t p = 1; p <= npat; ++p) SDPSR_REQUIRE(rep[(size_t)p] != 0xffffffffu, SDPSR_E_CUDA, "internal: pattern without representative");
    SDPSR_CUDA(cudaMemcpyAsync(d_rep, rep.data() + 1, (size_t)npat * 4, cudaMemcpyHostToDevice, ctx->stream));
    pattern_rows_kernel<<<(unsigned)((npat * m + 255) / 256), 256, 0, ctx->stream>>>(c.d_col, c.d_val, c.d_rowptr, m, d_rep, npat, d_prow);
    count_launch(ctx);
    SDPSR_CUDA(cudaMemcpyAsync(prow.data(), d_prow, prow.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  for (int64_t p = 1; p <= npat; ++p) {
    c.pat_ptr[(size_t)p] = (int64_t)c.pat_row.size();
    for (int64_t k = 0; k < m; ++k) {
      const double v = prow[(size_t)(p - 1) * (size_t)m + (size_t)k];
      if (v != 0.0) {
        c.pat_row.push_back((int32_t)k);
        c.pat_val.push_back(v);
      }
    }
  }
  c.pat_ptr[(size_t)npat + 1] = (int64_t)c.pat_row.size();
  // consistency: pattern sizes must add up to nnz (catches a 64-bit signature collision)
  {
    int64_t tot = 0;
    for (int64_t p = 1; p <= npat; ++p) tot += (c.pat_ptr[p + 1] - c.pat_ptr[p]) * (int64_t)cnt[(size_t)p];
    SDPSR_REQUIRE(tot == nnz, SDPSR_E_CUDA, "internal: constraint pattern signature collision");
  }
  // ---- Gram matrix G = A A' = sum_p cnt[p] v_p v_p' and its LU ----------------------------
  {
    std::vector<long double> G((size_t)m * m, 0.0L);
    for (int64_t p = 1; p <= npat; ++p)
      for (int64_t e1 = c.pat_ptr[p]; e1 < c.pat_ptr[p + 1]; ++e1)
        for (int64_t e2 = c.pat_ptr[p]; e2 < c.pat_ptr[p + 1]; ++e2)
          G[(size_t)c.pat_row[e1] * m + c.pat_row[e2]] +=
              (long double)cnt[(size_t)p] * (long double)c.pat_val[e1] * (long double)c.pat_val[e2];
    gram_factor(c, G, (int)m);
  }
  SDPSR_REQUIRE(c.gram_rank >= 1, SDPSR_E_SINGULAR, "the constraint matrix A is zero");
  SDPSR_TRY(sdpsr_scratch_t(ctx, 14, (size_t)npat + 1, &c.d_tpat));
  SDPSR_CUDA(cudaMemsetAsync(c.d_tpat, 0, ((size_t)npat + 1) * sizeof(double), ctx->stream));
  // A' c is a symmetric matrix for every c iff the pattern ids are transpose-invariant: then the projection
  // of a symmetric element is symmetric and the refined partition stays transpose-invariant
  SDPSR_TRY(u32_symmetric(ctx, c.d_pid, &c.pid_sym));
  c.ready = true;
  return SDPSR_OK;
}

#define CTX_ENTER()                 \
  if (!ctx) return SDPSR_E_INVALID; \
  if (cudaSetDevice(ctx->device) != cudaSuccess) return ctx->fail(SDPSR_E_CUDA, "cudaSetDevice failed"); \
  ++ctx->api_seq

static int finish(sdpsr_ctx* ctx) {
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  SDPSR_CUDA(cudaGetLastError());
  return SDPSR_OK;
}

template <typename IT>
static int set_constraints_csr_impl(sdpsr_ctx* ctx, int64_t m, const int64_t* rowptr, const IT* colidx, const double* vals,
                                    int index_base) {
  SDPSR_REQUIRE(m >= 1 && rowptr && (index_base == 0 || index_base == 1), SDPSR_E_INVALID, "bad CSR arguments");
  sdpsr_constraints_free(ctx);
  ConstraintSet& c = ctx->cons;
  c.m = m;
  const int64_t nnz_in = rowptr[m] - rowptr[0];
  SDPSR_REQUIRE(nnz_in >= 0 && (nnz_in == 0 || (colidx && vals)), SDPSR_E_INVALID, "bad CSR arguments");
  int fallback = 0;
  SDPSR_TRY(constraints_from_csr<IT>(ctx, rowptr, colidx + rowptr[0], vals + rowptr[0], index_base, &fallback));
  if (!fallback) return SDPSR_OK;
  {
    cudaPointerAttributes pa;
    const bool on_device = cudaPointerGetAttributes(&pa, vals) == cudaSuccess && pa.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    SDPSR_REQUIRE(!on_device, SDPSR_E_INVALID, "a device-resident CSR must not hold explicit zeros");
  }
  // explicit zeros carry no constraint: filter them on the host
  c.h_rowptr.assign((size_t)m + 1, 0);
  std::vector<int64_t> hcol;
  std::vector<double> hval;
  hcol.reserve((size_t)nnz_in);
  hval.reserve((size_t)nnz_in);
  for (int64_t k = 0; k < m; ++k) {
    for (int64_t e = rowptr[k]; e < rowptr[k + 1]; ++e) {
      if (vals[e] == 0.0) continue;
      hcol.push_back((int64_t)colidx[e] - index_base);
      hval.push_back(vals[e]);
    }
    c.h_rowptr[(size_t)k + 1] = (int64_t)hcol.size();
  }
  return constraints_from_host(ctx, hcol, hval);
}

extern "C" int sdpsr_set_constraints_csr(sdpsr_ctx* ctx, int64_t m, const int64_t* rowptr, const int64_t* colidx,
                                         const double* vals, int index_base) {
  CTX_ENTER();
  if (ctx->staged_src) ctx->staged_seq = ctx->api_seq;      // a staged objective survives the constraint set-up
  SDPSR_TRY(set_constraints_csr_impl<int64_t>(ctx, m, rowptr, colidx, vals, index_base));
  return finish(ctx);
}

extern "C" int sdpsr_set_constraints_csr_i32(sdpsr_ctx* ctx, int64_t m, const int64_t* rowptr, const int32_t* colidx,
                                             const double* vals, int index_base) {
  CTX_ENTER();
  if (ctx->staged_src) ctx->staged_seq = ctx->api_seq;      // a staged objective survives the constraint set-up
  SDPSR_TRY(set_constraints_csr_impl<int32_t>(ctx, m, rowptr, colidx, vals, index_base));
  return finish(ctx);
}

extern "C" int sdpsr_set_constraints_dense(sdpsr_ctx* ctx, int64_t m, const double* A) {
  CTX_ENTER();
  if (ctx->staged_src) ctx->staged_seq = ctx->api_seq;      // a staged objective survives the constraint set-up
  SDPSR_REQUIRE(m >= 1 && A, SDPSR_E_INVALID, "bad dense constraint arguments");
  sdpsr_constraints_free(ctx);
  ConstraintSet& c = ctx->cons;
  c.m = m;
  const int64_t nn = ctx->n * ctx->n;
  std::vector<int64_t> cntk((size_t)m, 0);
  for (int64_t idx = 0; idx < nn; ++idx)
    for (int64_t k = 0; k < m; ++k)
      if (A[k + m * idx] != 0.0) ++cntk[(size_t)k];
  c.h_rowptr.assign((size_t)m + 1, 0);
  for (int64_t k = 0; k < m; ++k) c.h_rowptr[(size_t)k + 1] = c.h_rowptr[(size_t)k] + cntk[(size_t)k];
  std::vector<int64_t> hcol((size_t)c.h_rowptr[m]);
  std::vector<double> hval((size_t)c.h_rowptr[m]);
  std::vector<int64_t> pos(c.h_rowptr.begin(), c.h_rowptr.end() - 1);
  for (int64_t idx = 0; idx < nn; ++idx)
    for (int64_t k = 0; k < m; ++k) {
      const double v = A[k + m * idx];
      if (v != 0.0) {
        hcol[(size_t)pos[k]] = idx;
        hval[(size_t)pos[k]++] = v;
      }
    }
  SDPSR_TRY(constraints_from_host(ctx, hcol, hval));
  return finish(ctx);
}

extern "C" int sdpsr_set_constraints_csc(sdpsr_ctx* ctx, int64_t m, const int64_t* colptr, const int64_t* rowval,
                                         const double* nzval, int index_base) {
  CTX_ENTER();
  if (ctx->staged_src) ctx->staged_seq = ctx->api_seq;      // a staged objective survives the constraint set-up
  SDPSR_REQUIRE(m >= 1 && colptr && (index_base == 0 || index_base == 1), SDPSR_E_INVALID, "bad CSC arguments");
  sdpsr_constraints_free(ctx);
  ConstraintSet& c = ctx->cons;
  c.m = m;
  const int64_t nn = ctx->n * ctx->n;
  const int64_t off = colptr[0];
  std::vector<int64_t> cntk((size_t)m, 0);
  for (int64_t e = 0; e < colptr[nn] - off; ++e) {
    const int64_t k = rowval[e] - index_base;
    SDPSR_REQUIRE(k >= 0 && k < m, SDPSR_E_INVALID, "CSC row index out of range");
    if (nzval[e] != 0.0) ++cntk[(size_t)k];
  }
  c.h_rowptr.assign((size_t)m + 1, 0);
  for (int64_t k = 0; k < m; ++k) c.h_rowptr[(size_t)k + 1] = c.h_rowptr[(size_t)k] + cntk[(size_t)k];
  std::vector<int64_t> hcol((size_t)c.h_rowptr[m]);
  std::vector<double> hval((size_t)c.h_rowptr[m]);
  std::vector<int64_t> pos(c.h_rowptr.begin(), c.h_rowptr.end() - 1);
  for (int64_t idx = 0; idx < nn; ++idx)
    for (int64_t e = colptr[idx] - off; e < colptr[idx + 1] - off; ++e) {
      if (nzval[e] == 0.0) continue;
      const int64_t k = rowval[e] - index_base;
      hcol[(size_t)pos[k]] = idx;
      hval[(size_t)pos[k]++] = nzval[e];
    }
  SDPSR_TRY(constraints_from_host(ctx, hcol, hval));
  return finish(ctx);
}

extern "C" int sdpsr_constraint_patterns(sdpsr_ctx* ctx, int64_t* npatterns) {
  CTX_ENTER();
  SDPSR_REQUIRE(ctx->cons.ready && npatterns, SDPSR_E_STATE, "constraints not set");
  *npatterns = ctx->cons.npat;
  return SDPSR_OK;
}

extern "C" int sdpsr_constraint_rank(sdpsr_ctx* ctx, int64_t* rank) {
  CTX_ENTER();
  SDPSR_REQUIRE(ctx->cons.ready && rank, SDPSR_E_STATE, "constraints not set");
  *rank = ctx->cons.gram_rank;
  return SDPSR_OK;
}

// x .-= proj(x); round; S = refine!(S, Part(X))   (src/partitions.jl:160-164)
extern "C" int sdpsr_project_round_refine(sdpsr_ctx* ctx, double atol, int64_t* dim) {
  CTX_ENTER();
  ConstraintSet& c = ctx->cons;
  SDPSR_REQUIRE(c.ready, SDPSR_E_STATE, "constraints not set (sdpsr_set_constraints_*)");
  SDPSR_REQUIRE(ctx->x_valid && ctx->x_is_fill, SDPSR_E_STATE,
                "sdpsr_project_round_refine must follow sdpsr_fill (src/partitions.jl:159-161)");
  std::vector<double> coef;
  SDPSR_TRY(sdpsr_rowdots(ctx, nullptr, ctx->lut, coef));   // A x, x = fill(S, values)
  SDPSR_TRY(sdpsr_solve_gram(ctx, coef));                   // (A A')^-1 A x
  SDPSR_TRY(sdpsr_upload_tpat(ctx, coef));
  RefineSpec sp;
  sp.fillproj = true;
  sp.lut = ctx->lut;
  sp.tpat = c.d_tpat;
  sp.pid = c.d_pid;
  sp.atol = atol;
  sp.do_round = true;
  sp.keeps_symmetry = c.pid_sym == 1;
  double sc;
  long long isc;
  int qb;
  SDPSR_TRY(sdpsr_round_params(ctx, atol, &sc, &isc, &qb));
  if (12 + qb + sdpsr_label_bits(ctx) <= 64) {
    // The projected, rounded element stays as X (:163) -- but not as N^2 doubles: the keys of the pass carry
    // the rounded values, so X = fill(S_new, lut) with lut decoded from the new table (12 B/entry pass).
    sp.mode = KM_ROUND;
    SDPSR_TRY(sdpsr_refine_pass(ctx, sp, dim));
    SDPSR_TRY(sdpsr_decode_lut(ctx, atol));
    ctx->x_valid = true;
    ctx->x_is_fill = true;
    return finish(ctx);
  } else {
    sp.vals_out = ctx->X;
    // wide rounding grid: materialise X first, then the generic two-step refine
    SDPSR_TRY(sdpsr_ensure_tmp_labels(ctx));
    KeyTable& scratch = ctx->tab_scratch;   // reused across calls, freed with the context
    sp.mode = KM_RAW;
    sp.out_override = ctx->labels_tmp;
    sp.table_override = &scratch;      // (rank-local full-range pass: X must be complete on every rank)
    int st = sdpsr_refine_pass(ctx, sp, nullptr);
    if (st == SDPSR_OK) {
      RefineSpec pr;
      pr.mode = KM_PAIR;
      pr.lab2 = ctx->labels_tmp;
      pr.do_round = false;
      pr.keeps_symmetry = sp.keeps_symmetry;
      st = sdpsr_refine_pass(ctx, pr, dim);
    }
    SDPSR_TRY(st);
  }
  ctx->x_valid = true;
  ctx->x_is_fill = false;
  return finish(ctx);
}

// The initial partition of src/partitions.jl:129-146 on the device.
extern "C" int sdpsr_init_partition(sdpsr_ctx* ctx, const double* C, const double* b, double atol,
                                    int snap_decimals, int64_t* dim) {
  CTX_ENTER();
  ConstraintSet& c = ctx->cons;
  SDPSR_REQUIRE(c.ready, SDPSR_E_STATE, "constraints not set (sdpsr_set_constraints_*)");
  SDPSR_REQUIRE(C && b, SDPSR_E_INVALID, "C or b is NULL");
  SDPSR_REQUIRE(snap_decimals <= 15, SDPSR_E_INVALID, "snap_decimals must be <= 15");
  double scale;
  long long iscale;
  int qbits;
  SDPSR_TRY(sdpsr_round_params(ctx, atol, &scale, &iscale, &qbits));
  const int do_snap = snap_decimals >= 0 ? 1 : 0;
  double pw = 1.0;
  for (int i = 0; i < snap_decimals; ++i) pw *= 10.0;
  const int grid = (int)std::min<uint64_t>((ctx->elems + 255) / 256, (uint64_t)ctx->sm_count * 16);
  const unsigned nb = (unsigned)((ctx->n + 31) / 32);
  std::vector<double> coef;

  // This rank's column block (all columns in a single-rank run): the elementwise steps and both refine
  // passes touch nothing else.
  const bool shard = sdpsr_shard_active(ctx);
  int64_t col0 = 0, col1 = ctx->n;
  if (shard) {
    col0 = ctx->n * ctx->rank / ctx->nranks;
    col1 = ctx->n * (ctx->rank + 1) / ctx->nranks;
  }
  const uint64_t e0 = (uint64_t)col0 * (uint64_t)ctx->ld, e1 = (uint64_t)col1 * (uint64_t)ctx->ld;
  const dim3 sgrid(nb, (unsigned)std::max<int64_t>(1, ((col1 + 31) / 32) - col0 / 32));
  // CL = symmetrize(round(C - proj(C)))                                   (:129-134)
  // C is read in place when it already lives on this device in the engine's own layout (ld == N);
  // otherwise it is staged into X (the H2D upload of a host matrix, or the re-striding copy).  With several
  // ranks a HOST matrix is uploaded once in total: every rank copies its own column block over PCIe and the
  // blocks are exchanged over NVLink.
  const double* Csrc = ctx->X;
  {
    cudaPointerAttributes pa;
    const bool on_device = cudaPointerGetAttributes(&pa, C) == cudaSuccess && pa.type == cudaMemoryTypeDevice &&
                           pa.device == ctx->device;
    cudaGetLastError();
    int split = (!on_device && ctx->nranks > 1) ? 1 : 0;
    if (ctx->nranks > 1) SDPSR_TRY(sdpsr_comm_agree_min(ctx, &split));     // a branch with a collective
    // a staged upload of this very matrix (sdpsr_stage_objective) with nothing but constraint set-up since
    bool staged = false;
    if (ctx->staged_src) {
      SDPSR_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->staged_ev, 0));     // (also when it is dropped: X is ours again)
      staged = ctx->staged_src == C && ctx->staged_seq + 1 == ctx->api_seq && !on_device &&
               (ctx->nranks > 1) == (split != 0);
      ctx->staged_src = nullptr;
    }
    if (on_device && ctx->ld == ctx->n && ((uintptr_t)C % 16 == 0)) {
      Csrc = C;
    } else {
      if (ctx->ld != ctx->n && !staged) SDPSR_CUDA(cudaMemsetAsync(ctx->X, 0, ctx->elems * 8, ctx->stream));
      Timed tm(ctx, SDPSR_K_MISC, (double)ctx->elems * 16.0);              // staging copy (or the H2D upload)
      if (staged && !split) {
        // sdpsr_stage_objective already moved this matrix into X on the copy stream
      } else if (split) {
        const int64_t b0 = ctx->n * ctx->rank / ctx->nranks, b1 = ctx->n * (ctx->rank + 1) / ctx->nranks;
        if (b1 > b0 && !staged)          // (staged: this rank's block is already in place)
          SDPSR_CUDA(cudaMemcpy2DAsync(ctx->X + b0 * ctx->ld, (size_t)ctx->ld * 8, C + b0 * ctx->n, (size_t)ctx->n * 8,
                                       (size_t)ctx->n * 8, (size_t)(b1 - b0), cudaMemcpyDefault, ctx->stream));
        size_t off[sdpsr_ctx::MAX_RANKS], len[sdpsr_ctx::MAX_RANKS];
        for (int r = 0; r < ctx->nranks; ++r) {
          const int64_t r0 = ctx->n * r / ctx->nranks, r1 = ctx->n * (r + 1) / ctx->nranks;
          off[r] = (size_t)r0 * (size_t)ctx->ld * 8;
          len[r] = (size_t)(r1 - r0) * (size_t)ctx->ld * 8;
        }
        SDPSR_TRY(sdpsr_comm_allgatherv(ctx, ctx->X, off, len));
      } else {
        SDPSR_CUDA(cudaMemcpy2DAsync(ctx->X, (size_t)ctx->ld * 8, C, (size_t)ctx->n * 8, (size_t)ctx->n * 8, (size_t)ctx->n,
                                     cudaMemcpyDefault, ctx->stream));
      }
    }
  }
  SDPSR_TRY(sdpsr_rowdots(ctx, Csrc, nullptr, coef));
  SDPSR_TRY(sdpsr_solve_gram(ctx, coef));
  SDPSR_TRY(sdpsr_upload_tpat(ctx, coef));
  {
    Timed tm(ctx, SDPSR_K_MISC, (double)(e1 - e0) * (24.0 + 8.0));
    init_sym_kernel<<<sgrid, 256, 0, ctx->stream>>>(0, Csrc, c.d_pid, c.d_tpat, ctx->X2, ctx->n, ctx->ld, col0, col1, pw, do_snap,
                                                    atol, scale);
    count_launch(ctx);
  }
  SDPSR_CUDA(cudaGetLastError());
  SDPSR_TRY(sdpsr_partition_reset(ctx));
  {
    RefineSpec sp;                                                        // S = Part(CL)   (:145)
    sp.mode = KM_RAW;
    sp.vals = ctx->X2;
    sp.do_round = false;
    sp.ignore_labels = true;
    sp.keeps_symmetry = true;                                             // CL was symmetrised
    SDPSR_TRY(sdpsr_refine_pass(ctx, sp, nullptr));
  }
  // X0 = round(proj(symmetrize(x0))), x0 = A' (A A')^-1 b                   (:137-142)
  coef.assign(b, b + c.m);
  SDPSR_TRY(sdpsr_solve_gram(ctx, coef));
  SDPSR_TRY(sdpsr_upload_tpat(ctx, coef));
  {
    Timed tm(ctx, SDPSR_K_MISC, (double)(e1 - e0) * (8.0 + 8.0));
    init_sym_kernel<<<sgrid, 256, 0, ctx->stream>>>(1, nullptr, c.d_pid, c.d_tpat, ctx->X, ctx->n, ctx->ld, col0, col1, pw, do_snap,
                                                    atol, scale);
    count_launch(ctx);
  }
  SDPSR_TRY(sdpsr_rowdots(ctx, ctx->X, nullptr, coef, /*x_own_block=*/true));
  SDPSR_TRY(sdpsr_solve_gram(ctx, coef));
  SDPSR_TRY(sdpsr_upload_tpat(ctx, coef));
  if (e1 > e0) {
    Timed tm(ctx, SDPSR_K_MISC, (double)(e1 - e0) * 20.0);
    const int g2 = (int)std::min<uint64_t>((e1 - e0 + 255) / 256, (uint64_t)ctx->sm_count * 16);
    init_elem_kernel<<<g2, 256, 0, ctx->stream>>>(2, nullptr, c.d_pid + e0, c.d_tpat, ctx->X2 + e0, e1 - e0, pw, do_snap, atol, scale);
    count_launch(ctx);
  }
  SDPSR_CUDA(cudaGetLastError());
  SDPSR_TRY(sdpsr_generic_refine_values(ctx, ctx->X2, atol, false, nullptr, dim,
                                        /*keeps_symmetry=*/c.pid_sym == 1));          // refine!(S, Part(X0)) (:146)
  ctx->x_valid = false;
  ctx->x_is_fill = false;
  return finish(ctx);
}

// newA = A * PMat (m x dim, column-major), newC = C' * PMat (dim)   (README.md:57-60)
extern "C" int sdpsr_reduce_problem(sdpsr_ctx* ctx, const double* C, double* newA, double* newC) {
  CTX_ENTER();
  ConstraintSet& c = ctx->cons;
  SDPSR_REQUIRE(newA == nullptr || c.ready, SDPSR_E_STATE, "constraints not set (sdpsr_set_constraints_*)");
  const int64_t d = ctx->dim, m = c.m;
  if (d == 0) return SDPSR_OK;
  SDPSR_TRY(sdpsr_shard_ensure_full_labels(ctx));
  KeyTable& t = ctx->tab[ctx->cur];
  const size_t smem = d <= RB_MAX ? (size_t)d * sizeof(double) : 0;
  double* dA = nullptr;
  double* dC = nullptr;
  int st = SDPSR_OK;
  if (newA) {
    SDPSR_TRY(sdpsr_scratch_t(ctx, 0, (size_t)m * d, &dA));
    SDPSR_CUDA(cudaMemsetAsync(dA, 0, (size_t)m * d * sizeof(double), ctx->stream));
    if (c.nchunks) {
      reduce_rows_kernel<<<(unsigned)c.nchunks, 256, smem, ctx->stream>>>(c.d_col, c.d_val, c.d_chunk_row, c.d_chunk_beg,
                                                                          c.d_chunk_beg + c.nchunks, ctx->labels, t.rank,
                                                                          m, d, dA);
      count_launch(ctx);
    }
    if (cudaMemcpyAsync(newA, dA, (size_t)m * d * sizeof(double), cudaMemcpyDefault, ctx->stream) != cudaSuccess)
      st = ctx->fail(SDPSR_E_CUDA, "copy of newA failed");
  }
  if (newC && st == SDPSR_OK) {
    SDPSR_REQUIRE(C != nullptr, SDPSR_E_INVALID, "C is NULL");
    if (ctx->ld != ctx->n) cudaMemsetAsync(ctx->X2, 0, ctx->elems * 8, ctx->stream);
    cudaMemcpy2DAsync(ctx->X2, (size_t)ctx->ld * 8, C, (size_t)ctx->n * 8, (size_t)ctx->n * 8, (size_t)ctx->n,
                      cudaMemcpyDefault, ctx->stream);
    SDPSR_TRY(sdpsr_scratch_t(ctx, 1, (size_t)d, &dC));
    cudaMemsetAsync(dC, 0, (size_t)d * sizeof(double), ctx->stream);
    const int grid = (int)std::min<uint64_t>((ctx->elems + 255) / 256, (uint64_t)ctx->sm_count * 4);
    reduce_objective_kernel<<<grid, 256, smem, ctx->stream>>>(ctx->X2, ctx->labels, t.rank, ctx->elems, d, dC);
    count_launch(ctx);
    if (cudaMemcpyAsync(newC, dC, (size_t)d * sizeof(double), cudaMemcpyDefault, ctx->stream) != cudaSuccess)
      st = ctx->fail(SDPSR_E_CUDA, "copy of newC failed");
  }
  const int fin = finish(ctx);
  return st != SDPSR_OK ? st : fin;
}
