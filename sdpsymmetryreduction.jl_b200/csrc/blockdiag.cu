// blockdiag.cu -- device side of blockDiagonalize (Murota et al., Alg. 4.1).
//
// Replaces  eigen(A)                  src/eigen_decomposition.jl:246   (cuSOLVER Xsyevd -- the
//                                                                       one library call on the path)
//           Q' A Q + block_norms      src/eigen_decomposition.jl:177-204
//           irreducible_decomposition src/eigen_decomposition.jl:295-348
//           basis_image               src/diagonalize.jl:64-89
// The scalar decisions in between (eigenvalue clustering, Otsu threshold, union-find,
// consistency check) stay in the host language, exactly as the reference has them.
#include <cusolverDn.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <mutex>
#include <cstring>

#include "sdpsr_internal.cuh"

namespace {

__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                        int64_t n, int64_t ld) {
  __shared__ double t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t bi = blockIdx.x, bj = blockIdx.y;
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = bj * 32 + tx, j = bi * 32 + r;
    t[r][tx] = (i < n && j < n) ? in[i + ld * j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = bi * 32 + tx, j = bj * 32 + r;
    if (i < n && j < n) out[i + ld * j] = t[tx][r];
  }
}

// norms[i + ne*j] = max |W[a,b]| over a in E_i, b in E_j, for i <= j with equal dimensions.
// Non-negative doubles order like their bit patterns, so atomicMax on the bits is exact.
__global__ void __launch_bounds__(256) block_max_kernel(const double* __restrict__ W, const double* __restrict__ Wim,
                                                        int64_t n, int64_t ld,
                                                        const uint32_t* __restrict__ space,
                                                        const uint32_t* __restrict__ sdim, int64_t ne,
                                                        unsigned long long* __restrict__ norms) {
  const int64_t b = blockIdx.y;
  const uint32_t sj = space[b];
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const int64_t nround = (n + step - 1) / step * step;
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < nround; a += step) {
    const bool valid = a < n;
    uint32_t si = 0xffffffffu;
    unsigned long long bits = 0ull;
    if (valid) {
      si = space[a];
      if (si <= sj && sdim[si] == sdim[sj])
        bits = (unsigned long long)__double_as_longlong(
            Wim ? hypot(W[a + ld * b], Wim[a + ld * b]) : fabs(W[a + ld * b]));   // abs(::Complex) = hypot
      else
        si = 0xfffffffeu;
    }
    // lanes of one warp mostly share the eigenspace: reduce inside equal-key groups first
    const unsigned mask = __match_any_sync(0xffffffffu, si);
    unsigned long long m = bits;
    for (int o = 16; o; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
      const uint32_t osi = __shfl_xor_sync(0xffffffffu, si, o);
      if (osi == si && other > m) m = other;
    }
    // after the butterfly only lanes whose whole warp shares the key hold the group max; fall back
    // to one atomic per lane otherwise (rare: only at eigenspace boundaries)
    if (valid && si < 0xfffffffeu) {
      if (mask == 0xffffffffu) {
        if ((threadIdx.x & 31) == 0 && m) atomicMax(norms + si + ne * sj, m);
      } else if (bits) {
        atomicMax(norms + si + ne * sj, bits);
      }
    }
  }
}

// F[:, f] = Q[:, col[f]]
__global__ void gather_cols_kernel(const double* __restrict__ Q, int64_t ld, const int64_t* __restrict__ cols,
                                   double* __restrict__ F) {
  const int64_t f = blockIdx.y;
  const double* src = Q + ld * cols[f];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ld; i += (int64_t)gridDim.x * blockDim.x)
    F[i + ld * f] = src[i];
}

// one CTA per (pair, t):   out[pair*maxm + t] = dot(Q[:, qcol[pair] + t], V[:, vcol[pair]])
__global__ void __launch_bounds__(256) pair_dot_kernel(const double* __restrict__ Q, const double* __restrict__ V,
                                                       int64_t n, int64_t ld, const int64_t* __restrict__ qcol,
                                                       const int64_t* __restrict__ vcol,
                                                       const int64_t* __restrict__ mult, int64_t maxm,
                                                       double* __restrict__ out) {
  __shared__ double ws[8];
  const int64_t pair = blockIdx.y, t = blockIdx.x;
  if (t >= mult[pair]) return;
  const double* q = Q + ld * (qcol[pair] + t);
  const double* v = V + ld * vcol[pair];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += q[i] * v[i];
  for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tt = 0.0;
    for (int w = 0; w < 8; ++w) tt += ws[w];
    out[pair * maxm + t] = tt;
  }
}

// Qhat[:, dst[pair]] = Q[:, qcol[pair] .. + mult) * u[pair, :] / ||w[pair, :]||
__global__ void __launch_bounds__(256) pair_combine_kernel(const double* __restrict__ Q, int64_t n, int64_t ld,
                                                           const int64_t* __restrict__ qcol,
                                                           const int64_t* __restrict__ mult, int64_t maxm,
                                                           const double* __restrict__ u, const double* __restrict__ w,
                                                           const int64_t* __restrict__ dst, double* __restrict__ Qhat) {
  const int64_t pair = blockIdx.y;
  const int64_t mlt = mult[pair];
  double nrm2 = 0.0;
  for (int64_t t = 0; t < mlt; ++t) nrm2 += w[pair * maxm + t] * w[pair * maxm + t];
  const double nrm = sqrt(nrm2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int64_t t = 0; t < mlt; ++t) s += Q[i + ld * (qcol[pair] + t)] * (u[pair * maxm + t] / nrm);
    Qhat[i + ld * dst[pair]] = s;
  }
}

__global__ void clamp_kernel(double* __restrict__ x, uint64_t total, double atol) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
    if (fabs(x[i]) < atol) x[i] = 0.0;
}

// Qt[c + S*r] = Qhat[r + ld*c]  (row-major copy so that one matrix row is contiguous)
__global__ void qhat_rowmajor_kernel(const double* __restrict__ Qhat, int64_t n, int64_t ld, int64_t S,
                                     double* __restrict__ Qt) {
  const int64_t c = blockIdx.y;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
    Qt[c + S * r] = Qhat[r + ld * c];
}

// --- CSR-by-label (the reference's _constraints, src/diagonalize.jl:42-50) ----------------
// Persistent CTAs walk whole columns.  Class counters live in shared memory (SM_BINS classes); lanes
// holding the same class are aggregated with __match_any_sync, so the global atomics drop from one per
// warp-group per 32 entries to one per class per column (scatter) or per CTA (count).
constexpr int SM_BINS = 8192;      // class_count: 32 KB of counters
constexpr int SCAT_BINS = 2048;    // class_scatter: counters + 8-byte bases

__global__ void __launch_bounds__(256) class_count_kernel(const uint32_t* __restrict__ lab,
                                                          const uint32_t* __restrict__ rank, int64_t n, int64_t ld,
                                                          unsigned long long* __restrict__ cnt, int64_t nbins) {
  __shared__ uint32_t bins[SM_BINS];
  const bool smem = nbins <= SM_BINS;
  if (smem) {
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) bins[i] = 0u;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int64_t nround = (n + blockDim.x - 1) / blockDim.x * blockDim.x;      // warp-uniform trip count
  for (int64_t j = blockIdx.x; j < n; j += gridDim.x) {
    for (int64_t i = threadIdx.x; i < nround; i += blockDim.x) {
      const bool valid = i < n;
      uint32_t c = valid ? lab[i + ld * j] : 0xffffffffu;
      if (valid && rank) c = rank[c];
      const unsigned mask = __match_any_sync(0xffffffffu, c);
      if (valid && c < (uint32_t)nbins && (int)(__ffs(mask) - 1) == lane) {
        if (smem) {
          // a column has < 2^32 entries, but a CTA's share of the matrix may not: flush on the way
          const uint32_t old = atomicAdd(&bins[c], (uint32_t)__popc(mask));
          if (old > 0xf0000000u) {
            atomicAdd(cnt + c, (unsigned long long)atomicExch(&bins[c], 0u));
          }
        } else {
          atomicAdd(cnt + c, (unsigned long long)__popc(mask));
        }
      }
    }
  }
  if (smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += blockDim.x)
      if (bins[i]) atomicAdd(cnt + i, (unsigned long long)bins[i]);
  }
}

// rows/cols of every entry, grouped by class: cursor[c] is the next free position of class c
__global__ void __launch_bounds__(256) class_scatter_kernel(const uint32_t* __restrict__ lab,
                                                            const uint32_t* __restrict__ rank, int64_t n, int64_t ld,
                                                            unsigned long long* __restrict__ cursor,
                                                            uint32_t* __restrict__ rows, uint32_t* __restrict__ cols,
                                                            int64_t nbins) {
  __shared__ uint32_t s_cnt[SCAT_BINS];
  __shared__ unsigned long long s_base[SCAT_BINS];
  const bool smem = nbins <= SCAT_BINS;
  const int lane = threadIdx.x & 31;
  const int64_t nround = (n + blockDim.x - 1) / blockDim.x * blockDim.x;
  for (int64_t j = blockIdx.x; j < n; j += gridDim.x) {
    if (smem) {
      for (int i = threadIdx.x; i < nbins; i += blockDim.x) s_cnt[i] = 0u;
      __syncthreads();
      // phase 1: class sizes inside this column
      for (int64_t i = threadIdx.x; i < nround; i += blockDim.x) {
        const bool valid = i < n;
        const uint32_t c = valid ? rank[lab[i + ld * j]] : 0xffffffffu;
        const unsigned mask = __match_any_sync(0xffffffffu, c);
        if (valid && c != 0u && (int)(__ffs(mask) - 1) == lane) atomicAdd(&s_cnt[c], (uint32_t)__popc(mask));
      }
      __syncthreads();
      // phase 2: one global reservation per class present in the column
      for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
        const uint32_t k = s_cnt[i];
        if (k) s_base[i] = atomicAdd(cursor + i, (unsigned long long)k);
        s_cnt[i] = 0u;                                   // becomes the running offset of phase 3
      }
      __syncthreads();
    }
    // phase 3: place the entries
    for (int64_t i = threadIdx.x; i < nround; i += blockDim.x) {
      const bool valid = i < n;
      const uint32_t c = valid ? rank[lab[i + ld * j]] : 0xffffffffu;
      const unsigned mask = __match_any_sync(0xffffffffu, c);
      const int leader = __ffs(mask) - 1;
      unsigned long long basepos = 0ull;
      if (valid && c != 0u && lane == leader)
        basepos = smem ? s_base[c] + atomicAdd(&s_cnt[c], (uint32_t)__popc(mask))
                       : atomicAdd(cursor + c, (unsigned long long)__popc(mask));
      basepos = __shfl_sync(0xffffffffu, basepos, leader);
      if (valid && c != 0u) {
        const unsigned long long pos = basepos + __popc(mask & ((1u << lane) - 1u));
        rows[pos] = (uint32_t)i;
        cols[pos] = (uint32_t)j;
      }
    }
    if (smem) __syncthreads();
  }
}

constexpr int BI_PAIRS = 16;     // output elements accumulated per thread
constexpr int BI_CHUNK = 4096;   // entries per CTA

// partial[chunk][p] = sum over the chunk's entries (r,c) of Qt[r][ca[p]] * Qt[c][cb[p]]
__global__ void __launch_bounds__(256) basis_partial_kernel(const uint32_t* __restrict__ rows,
                                                            const uint32_t* __restrict__ cols,
                                                            const unsigned long long* __restrict__ chunk_beg,
                                                            const unsigned long long* __restrict__ chunk_end,
                                                            const double* __restrict__ Qt, int64_t S,
                                                            const int* __restrict__ ca, const int* __restrict__ cb,
                                                            int npairs, double* __restrict__ partial) {
  __shared__ int sa[BI_PAIRS], sb[BI_PAIRS];
  __shared__ double red[8][BI_PAIRS];
  if (threadIdx.x < BI_PAIRS) {
    sa[threadIdx.x] = threadIdx.x < npairs ? ca[threadIdx.x] : 0;
    sb[threadIdx.x] = threadIdx.x < npairs ? cb[threadIdx.x] : 0;
  }
  __syncthreads();
  double acc[BI_PAIRS];
#pragma unroll
  for (int p = 0; p < BI_PAIRS; ++p) acc[p] = 0.0;
  const unsigned long long e0 = chunk_beg[blockIdx.x], e1 = chunk_end[blockIdx.x];
  for (unsigned long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    const double* qr = Qt + S * (int64_t)rows[e];
    const double* qc = Qt + S * (int64_t)cols[e];
#pragma unroll
    for (int p = 0; p < BI_PAIRS; ++p)
      if (p < npairs) acc[p] += qr[sa[p]] * qc[sb[p]];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int p = 0; p < BI_PAIRS; ++p) {
    double v = acc[p];
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid][p] = v;
  }
  __syncthreads();
  if (threadIdx.x < npairs) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    partial[(int64_t)blockIdx.x * BI_PAIRS + threadIdx.x] = t;
  }
}


// ---------------------------------------------------------------------------------------------
// basis_image for partitions with few classes and small blocks (association schemes): ONE pass
// over the label matrix, no per-class index lists.  Every thread keeps private bins
// [class][output element] in shared memory (layout [bin][thread]: conflict-free, no atomics), so
// the accumulation order is a fixed function of the launch geometry; the host adds the per-CTA
// partials in CTA order.   bins = d * npairs <= BS_MAX_BINS.
// ---------------------------------------------------------------------------------------------
constexpr int BS_THREADS = 128;
constexpr int BS_MAX_BINS = 200;      // 200 * 128 * 8 B = 200 KB of shared memory
constexpr int BS_MAX_PAIRS = 16;

template <int NP>
__global__ void __launch_bounds__(BS_THREADS) basis_small_kernel(const uint32_t* __restrict__ lab,
                                                                 const uint32_t* __restrict__ rank, int64_t n, int64_t ld,
                                                                 const double* __restrict__ Qt, int64_t S,
                                                                 const int* __restrict__ ca, const int* __restrict__ cb,
                                                                 int d, int64_t colw, double* __restrict__ partial) {
  // CTA (x, y): rows x*128 .. +127 (one per thread, its Qt row lives in registers), columns
  // y*colw .. +colw-1; the column's Qt row is a warp-uniform (broadcast) load.
  extern __shared__ double bs_bins[];            // [d * NP][BS_THREADS]
  const int tid = threadIdx.x;
  const int nbins = d * NP;
  for (int k = 0; k < nbins; ++k) bs_bins[k * BS_THREADS + tid] = 0.0;
  int sb[NP];
  double qa[NP];
  const int64_t a = (int64_t)blockIdx.x * BS_THREADS + tid;
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    sb[p] = cb[p];
    qa[p] = a < n ? Qt[S * a + ca[p]] : 0.0;
  }
  const int64_t b0 = (int64_t)blockIdx.y * colw, b1 = min(n, b0 + colw);
  if (a < n && b0 < b1) {
    // software pipeline: the label of column b+2, the class of column b+1 and the Qt row of column b+1
    // are in flight while column b updates the bins (the update itself is shared-memory only)
    uint32_t lab2 = b0 + 1 < b1 ? lab[a + ld * (b0 + 1)] : 0u;
    uint32_t inext = rank[lab[a + ld * b0]];
    double qn[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) qn[p] = __ldg(Qt + S * b0 + sb[p]);
    for (int64_t b = b0; b < b1; ++b) {
      const uint32_t i = inext;
      double qb[NP];
#pragma unroll
      for (int p = 0; p < NP; ++p) qb[p] = qn[p];
      if (b + 1 < b1) {
        inext = rank[lab2];
#pragma unroll
        for (int p = 0; p < NP; ++p) qn[p] = __ldg(Qt + S * (b + 1) + sb[p]);
        if (b + 2 < b1) lab2 = lab[a + ld * (b + 2)];
      }
      if (i) {
        double* bin = bs_bins + (size_t)(i - 1) * NP * BS_THREADS + tid;
#pragma unroll
        for (int p = 0; p < NP; ++p) bin[p * BS_THREADS] += qa[p] * qb[p];
      }
    }
  }
  __syncthreads();
  // fixed-order reduction of the private copies (rotated start: no bank conflicts)
  const int64_t cta = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
  for (int k = tid; k < nbins; k += BS_THREADS) {
    double s = 0.0;
    for (int t = 0; t < BS_THREADS; ++t) s += bs_bins[k * BS_THREADS + ((t + tid) & (BS_THREADS - 1))];
    partial[cta * nbins + k] = s;
  }
}

template <int NP>
int launch_basis_small(sdpsr_ctx* ctx, dim3 grid, size_t smem, const uint32_t* rank, const double* qt, int64_t S,
                       const int* pairs, int d, int64_t colw, double* part) {
  SDPSR_CUDA(cudaFuncSetAttribute(basis_small_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
  basis_small_kernel<NP><<<grid, BS_THREADS, smem, ctx->stream>>>(ctx->labels, rank, ctx->n, ctx->ld, qt, S, pairs, pairs + NP,
                                                                 d, colw, part);
  return SDPSR_OK;
}

int ensure_buffer(sdpsr_ctx* ctx, double** p) {
  if (!*p) {
    SDPSR_CUDA(cudaMalloc(p, ctx->elems * sizeof(double)));
    SDPSR_CUDA(cudaMemsetAsync(*p, 0, ctx->elems * sizeof(double), ctx->stream));
  }
  return SDPSR_OK;
}

}  // namespace

int sdpsr_ensure_matrix(sdpsr_ctx* ctx, double** p) { return ensure_buffer(ctx, p); }

namespace {

// cuSOLVER is loaded with dlopen.  Order:
//   1. $SDPSR_CUSOLVER_LIB;
//   2. a libcusolver.so.11 the process has ALREADY loaded (a host framework's copy: it matches the cuBLAS /
//      cuSPARSE that framework brought, which is what its symbols bind to -- a second, newer cuSOLVER cannot be
//      loaded next to an older cuBLAS under the same SONAME);
//   3. the CUDA toolkit's copy, after its own dependencies (cuBLASLt, cuBLAS, cuSPARSE, nvJitLink) have been
//      loaded from the same directory -- a bare box has no LD_LIBRARY_PATH entry for /usr/local/cuda/lib64;
//   4. the SONAME through the default search path.
// Known limit of case 2: PyTorch 2.11+cu128 bundles cuSOLVER 11.7.3, whose Xsyevd_bufferSize rejects
// n >= 32767 (2 n^2 overflows an int); the toolkit's 11.7.5 accepts it.  The dense path reports that case.
struct SolverApi {
  void* lib = nullptr;
  std::string path;
  decltype(&cusolverDnCreate) Create = nullptr;
  decltype(&cusolverDnDestroy) Destroy = nullptr;
  decltype(&cusolverDnSetStream) SetStream = nullptr;
  decltype(&cusolverDnCreateParams) CreateParams = nullptr;
  decltype(&cusolverDnDestroyParams) DestroyParams = nullptr;
  decltype(&cusolverDnXsyevd_bufferSize) Xsyevd_bufferSize = nullptr;
  decltype(&cusolverDnXsyevd) Xsyevd = nullptr;
  decltype(&cusolverDnXgeev_bufferSize) Xgeev_bufferSize = nullptr;
  decltype(&cusolverDnXgeev) Xgeev = nullptr;
  bool ok = false;
};

static SolverApi load_solver_api() {
  SolverApi a;
  std::vector<std::string> names;
  if (const char* e = getenv("SDPSR_CUSOLVER_LIB")) names.push_back(e);
  names.push_back("@loaded");
  names.push_back("@toolkit");
  names.push_back("libcusolver.so.11");
  names.push_back("libcusolver.so");
  const char* tk = getenv("SDPSR_CUDA_LIB_DIR");
  const std::string tkdir = tk ? tk : "/usr/local/cuda/lib64";
  for (const std::string& nme : names) {
    void* h = nullptr;
    std::string shown = nme;
    if (nme == "@loaded") {
      h = dlopen("libcusolver.so.11", RTLD_NOW | RTLD_NOLOAD);
      shown = "libcusolver.so.11 (already loaded by the host process)";
    } else if (nme == "@toolkit") {
      if (dlopen("libcublas.so.12", RTLD_NOW | RTLD_NOLOAD)) continue;      // another cuBLAS is in: its cuSOLVER must follow
      for (const char* dep : {"libnvJitLink.so.12", "libcublasLt.so.12", "libcublas.so.12", "libcusparse.so.12"})
        dlopen((tkdir + "/" + dep).c_str(), RTLD_NOW | RTLD_GLOBAL);
      shown = tkdir + "/libcusolver.so.11";
      h = dlopen(shown.c_str(), RTLD_NOW | RTLD_LOCAL);
    } else {
      h = dlopen(nme.c_str(), RTLD_NOW | RTLD_LOCAL);
    }
    if (!h) continue;
    SolverApi t;
    t.lib = h;
    t.path = shown;
#define SDPSR_SYM(field, sym) t.field = reinterpret_cast<decltype(t.field)>(dlsym(h, #sym))
    SDPSR_SYM(Create, cusolverDnCreate);
    SDPSR_SYM(Destroy, cusolverDnDestroy);
    SDPSR_SYM(SetStream, cusolverDnSetStream);
    SDPSR_SYM(CreateParams, cusolverDnCreateParams);
    SDPSR_SYM(DestroyParams, cusolverDnDestroyParams);
    SDPSR_SYM(Xsyevd_bufferSize, cusolverDnXsyevd_bufferSize);
    SDPSR_SYM(Xsyevd, cusolverDnXsyevd);
    SDPSR_SYM(Xgeev_bufferSize, cusolverDnXgeev_bufferSize);
    SDPSR_SYM(Xgeev, cusolverDnXgeev);
#undef SDPSR_SYM
    t.ok = t.Create && t.Destroy && t.SetStream && t.CreateParams && t.DestroyParams && t.Xsyevd_bufferSize &&
           t.Xsyevd && t.Xgeev_bufferSize && t.Xgeev;
    if (t.ok) {
      a = t;
      break;
    }
    dlclose(h);
  }
  return a;
}

// (initialised once, by whichever thread comes first; the others block until the library is in -- in-process ranks
//  reach their first dense decomposition together, and a `tried` flag set before the load finished handed the late
//  threads an empty table)
SolverApi& sapi() {
  static SolverApi a = load_solver_api();
  return a;
}

struct Solver {
  cusolverDnHandle_t h = nullptr;
  cusolverDnParams_t params = nullptr;
  double* d_vals = nullptr;
};

}  // namespace

void sdpsr_blockdiag_free(sdpsr_ctx* ctx) {
  if (ctx->solver) {
    Solver* s = reinterpret_cast<Solver*>(ctx->solver);
    if (s->params) sapi().DestroyParams(s->params);
    if (s->h) sapi().Destroy(s->h);
    cudaFree(s->d_vals);
    delete s;
    ctx->solver = nullptr;
  }
  cudaFree(ctx->solver_work);
  cudaFree(ctx->solver_info);
  free(ctx->solver_hwork);
  cudaFree(ctx->Q);
  cudaFree(ctx->W);
  cudaFree(ctx->T);
  cudaFree(ctx->Xi);
  cudaFree(ctx->X2i);
  cudaFree(ctx->Qi);
  cudaFree(ctx->Wi);
  cudaFree(ctx->Ti);
  ctx->Xi = ctx->X2i = ctx->Qi = ctx->Wi = ctx->Ti = ctx->Qhat_i = nullptr;
  ctx->solver_work = nullptr;
  ctx->solver_info = nullptr;
  ctx->solver_hwork = nullptr;
  ctx->Q = ctx->W = ctx->T = ctx->Qhat = nullptr;
}

void sdpsr_blockdiag_rebind(sdpsr_ctx* ctx) {
  if (ctx->solver) sapi().SetStream(reinterpret_cast<Solver*>(ctx->solver)->h, ctx->stream);
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

static int fill_into(sdpsr_ctx* ctx, const double* r, int64_t len, double* dst) {
  SDPSR_REQUIRE(len == ctx->dim, SDPSR_E_INVALID, "length(values) != dim(P) (src/partitions.jl:69)");
  SDPSR_REQUIRE(r != nullptr || len == 0, SDPSR_E_INVALID, "coefficient vector is NULL");
  SDPSR_TRY(sdpsr_upload_values(ctx, r, len));
  SDPSR_TRY(sdpsr_build_lut(ctx, ctx->d_values, len));
  SDPSR_TRY(sdpsr_materialize_fill(ctx, dst));
  ctx->x_is_fill = false;   // the lut no longer belongs to X
  return SDPSR_OK;
}


// In-process ranks (one host thread each) reach their first dense decomposition at the same moment; handle creation
// loads the library's kernels and is serialised here (cusolverDnCreate is documented thread-safe, this only removes
// the concurrent first-use from the picture).
static cusolverStatus_t solver_create(cusolverDnHandle_t* h) {
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  return sapi().Create(h);
}

extern "C" int sdpsr_eig(sdpsr_ctx* ctx, const double* r1, int64_t len, double* vals) {
  CTX_ENTER();
  SDPSR_REQUIRE(vals != nullptr, SDPSR_E_INVALID, "vals is NULL");
  int sym = 0;
  SDPSR_TRY(sdpsr_symmetric_check(ctx, &sym));
  SDPSR_REQUIRE(sym, SDPSR_E_NOT_SYMMETRIC,
                "partition is not transpose-invariant: no real symmetric eigendecomposition "
                "(InvalidDecompositionField, src/eigen_decomposition.jl:247-253)");
  SDPSR_TRY(ensure_buffer(ctx, &ctx->Q));
  SDPSR_TRY(fill_into(ctx, r1, len, ctx->Q));
  if (!ctx->solver) {
    Solver* s = new Solver();
    ctx->solver = s;
    SDPSR_REQUIRE(sapi().ok, SDPSR_E_CUSOLVER, "libcusolver.so.11 could not be loaded");
    SDPSR_REQUIRE(solver_create(&s->h) == CUSOLVER_STATUS_SUCCESS, SDPSR_E_CUSOLVER, "cusolverDnCreate failed");
    SDPSR_REQUIRE(sapi().SetStream(s->h, ctx->stream) == CUSOLVER_STATUS_SUCCESS, SDPSR_E_CUSOLVER,
                  "cusolverDnSetStream failed");
    SDPSR_REQUIRE(sapi().CreateParams(&s->params) == CUSOLVER_STATUS_SUCCESS, SDPSR_E_CUSOLVER,
                  "cusolverDnCreateParams failed");
    SDPSR_CUDA(cudaMalloc(&s->d_vals, (size_t)ctx->n * sizeof(double)));
    SDPSR_CUDA(cudaMalloc(&ctx->solver_info, sizeof(int)));
  }
  Solver* s = reinterpret_cast<Solver*>(ctx->solver);
  size_t wdev = 0, whost = 0;
  if (ctx->rank == 0) {
    const cusolverStatus_t bs =
        sapi().Xsyevd_bufferSize(s->h, s->params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, ctx->n, CUDA_R_64F,
                                    ctx->Q, ctx->ld, CUDA_R_64F, s->d_vals, CUDA_R_64F, &wdev, &whost);
    SDPSR_REQUIRE(bs == CUSOLVER_STATUS_SUCCESS, SDPSR_E_CUSOLVER,
                  "cusolverDnXsyevd_bufferSize failed (status " + std::to_string((int)bs) + ", n = " +
                      std::to_string(ctx->n) + ", " + sapi().path + ")");
  }
  if (wdev > ctx->solver_work_bytes) {
    cudaFree(ctx->solver_work);
    ctx->solver_work = nullptr;
    SDPSR_CUDA(cudaMalloc(&ctx->solver_work, wdev));
    ctx->solver_work_bytes = wdev;
  }
  if (whost > ctx->solver_hwork_bytes) {
    free(ctx->solver_hwork);
    ctx->solver_hwork = malloc(whost);
    SDPSR_REQUIRE(ctx->solver_hwork != nullptr, SDPSR_E_ALLOC, "host workspace allocation failed");
    ctx->solver_hwork_bytes = whost;
  }
  cusolverStatus_t st = CUSOLVER_STATUS_SUCCESS;
  if (ctx->rank == 0) {          // multi-GPU: one rank factorises, all receive (identical Q everywhere)
    Timed tm(ctx, SDPSR_K_EIG, 0.0);
    st = sapi().Xsyevd(s->h, s->params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, ctx->n, CUDA_R_64F, ctx->Q,
                          ctx->ld, CUDA_R_64F, s->d_vals, CUDA_R_64F, ctx->solver_work, wdev, ctx->solver_hwork, whost,
                          ctx->solver_info);
  }
  SDPSR_REQUIRE(st == CUSOLVER_STATUS_SUCCESS, SDPSR_E_CUSOLVER,
                "cusolverDnXsyevd failed (status " + std::to_string((int)st) + ")");
  if (ctx->nranks > 1) {
    if (ctx->rank != 0) SDPSR_CUDA(cudaMemsetAsync(ctx->solver_info, 0, sizeof(int), ctx->stream));
    SDPSR_TRY(sdpsr_comm_bcast(ctx, ctx->Q, ctx->elems * sizeof(double), 0));
    SDPSR_TRY(sdpsr_comm_bcast(ctx, s->d_vals, (size_t)ctx->n * sizeof(double), 0));
    SDPSR_TRY(sdpsr_comm_bcast(ctx, ctx->solver_info, sizeof(int), 0));
  }
  int* hinfo = reinterpret_cast<int*>(ctx->h_pinned) + 64;
  SDPSR_CUDA(cudaMemcpyAsync(hinfo, ctx->solver_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaMemcpyAsync(vals, s->d_vals, (size_t)ctx->n * sizeof(double), cudaMemcpyDefault, ctx->stream));
  SDPSR_TRY(finish(ctx));
  SDPSR_REQUIRE(*hinfo == 0, SDPSR_E_CUSOLVER, "syevd did not converge (info = " + std::to_string(*hinfo) + ")");
  ctx->have_Q = true;
  ctx->q_complex = false;
  return SDPSR_OK;
}

extern "C" int sdpsr_block_norms(sdpsr_ctx* ctx, const double* r2, int64_t len, const int64_t* ptrs, int64_t nptr,
                                 double* norms) {
  CTX_ENTER();
  SDPSR_REQUIRE(ctx->have_Q, SDPSR_E_STATE, "sdpsr_block_norms must follow sdpsr_eig");
  SDPSR_REQUIRE(ptrs && nptr >= 2 && norms, SDPSR_E_INVALID, "bad eigenspace pointers");
  const int64_t ne = nptr - 1, n = ctx->n, ld = ctx->ld;
  SDPSR_REQUIRE(ptrs[0] == 0 && ptrs[ne] == n, SDPSR_E_INVALID, "ptrs must run from 0 to N");
  std::vector<uint32_t> space((size_t)n), sdim((size_t)ne);
  for (int64_t e = 0; e < ne; ++e) {
    SDPSR_REQUIRE(ptrs[e + 1] > ptrs[e], SDPSR_E_INVALID, "ptrs must be strictly increasing");
    sdim[(size_t)e] = (uint32_t)(ptrs[e + 1] - ptrs[e]);
    for (int64_t a = ptrs[e]; a < ptrs[e + 1]; ++a) space[(size_t)a] = (uint32_t)e;
  }
  SDPSR_TRY(ensure_buffer(ctx, &ctx->W));
  SDPSR_TRY(ensure_buffer(ctx, &ctx->T));
  // A2 = fill(S, r2) -> X ; T = A2 * Q ; X2 = Q' ; W = Q' * T            (:203)
  SDPSR_TRY(fill_into(ctx, r2, len, ctx->X));
  ctx->x_valid = false;
  SDPSR_TRY(sdpsr_gemm_f64(ctx, ctx->X, ld, ctx->Q, ld, ctx->T, ld, ld, n, n, false, /*shard=*/true));
  const unsigned nb = (unsigned)((n + 31) / 32);
  if (ld != n) SDPSR_CUDA(cudaMemsetAsync(ctx->X2, 0, ctx->elems * 8, ctx->stream));
  transpose_kernel<<<dim3(nb, nb), 256, 0, ctx->stream>>>(ctx->Q, ctx->X2, n, ld);
  count_launch(ctx);
  // W = Q' (A2 Q) is symmetric (A2 is: sdpsr_eig checked the partition): lower tiles + mirror
  SDPSR_TRY(sdpsr_gemm_f64(ctx, ctx->X2, ld, ctx->T, ld, ctx->W, ld, ld, n, n, !(ctx->flags & SDPSR_F_NO_SYRK),
                           /*shard=*/true));
  // block maxima
  uint32_t *d_space = nullptr, *d_sdim = nullptr;
  unsigned long long* d_norms = nullptr;
  SDPSR_TRY(sdpsr_scratch_t(ctx, 0, (size_t)n, &d_space));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 1, (size_t)ne, &d_sdim));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 2, (size_t)ne * ne, &d_norms));
  SDPSR_CUDA(cudaMemcpyAsync(d_space, space.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  SDPSR_CUDA(cudaMemcpyAsync(d_sdim, sdim.data(), (size_t)ne * 4, cudaMemcpyHostToDevice, ctx->stream));
  SDPSR_CUDA(cudaMemsetAsync(d_norms, 0, (size_t)ne * ne * 8, ctx->stream));
  {
    Timed tm(ctx, SDPSR_K_MISC, (double)ctx->elems * 8.0);
    block_max_kernel<<<dim3((unsigned)std::min<int64_t>((n + 255) / 256, 64), (unsigned)n), 256, 0, ctx->stream>>>(
        ctx->W, nullptr, n, ld, d_space, d_sdim, ne, d_norms);
    count_launch(ctx);
  }
  SDPSR_CUDA(cudaGetLastError());
  SDPSR_CUDA(cudaMemcpyAsync(norms, d_norms, (size_t)ne * ne * 8, cudaMemcpyDefault, ctx->stream));
  SDPSR_TRY(finish(ctx));
  // the kernel filled i <= j (bit patterns of non-negative doubles == the doubles); mirror
  for (int64_t j = 0; j < ne; ++j)
    for (int64_t i = 0; i < j; ++i) norms[j + ne * i] = norms[i + ne * j];
  return SDPSR_OK;
}

extern "C" int sdpsr_irreducible(sdpsr_ctx* ctx, const double* r3, int64_t len, const int64_t* ptrs, int64_t nptr,
                                 const int64_t* kroot, double atol, int64_t* blk_sizes, int64_t* nblk) {
  CTX_ENTER();
  SDPSR_REQUIRE(ctx->have_Q, SDPSR_E_STATE, "sdpsr_irreducible must follow sdpsr_eig");
  SDPSR_REQUIRE(ptrs && nptr >= 2 && kroot && blk_sizes && nblk, SDPSR_E_INVALID, "bad arguments");
  const int64_t ne = nptr - 1, n = ctx->n, ld = ctx->ld;
  // classes in order of their first (= root) eigenspace                     (:299-309)
  std::vector<std::vector<int64_t>> classes;
  std::vector<int64_t> class_of((size_t)ne, -1);
  for (int64_t e = 0; e < ne; ++e) {
    const int64_t r = kroot[e];
    SDPSR_REQUIRE(r >= 0 && r <= e && kroot[r] == r, SDPSR_E_INVALID,
                  "kroot[e] must be the smallest member of e's class (src/eigen_decomposition.jl:310)");
    if (r == e) {
      class_of[(size_t)e] = (int64_t)classes.size();
      classes.emplace_back();
    }
    classes[(size_t)class_of[(size_t)r]].push_back(e);
    class_of[(size_t)e] = class_of[(size_t)r];
  }
  int64_t S = 0;
  for (auto& k : classes) S += (int64_t)k.size();
  // Qhat buffer (ld x S)
  SDPSR_TRY(sdpsr_scratch_t(ctx, 16, (size_t)ld * (size_t)S, &ctx->Qhat));
  SDPSR_CUDA(cudaMemsetAsync(ctx->Qhat, 0, (size_t)ld * (size_t)S * 8, ctx->stream));
  ctx->qhat_cols = S;
  sdpsr_module_invalidate_qhat(ctx);
  ctx->blk_sizes.clear();
  // first columns, and the list of (root, member) pairs that need work
  std::vector<int64_t> first_src, first_dst;     // Qhat[:, dst] = Q[:, src]
  std::vector<int64_t> fcols;                    // eigenvector columns whose image under A3 is needed
  std::vector<int64_t> findex((size_t)ne, -1);
  struct Pair { int64_t i, j, dst; };
  std::vector<Pair> pairs;
  int64_t col = 0;
  for (auto& k : classes) {
    ctx->blk_sizes.push_back((int64_t)k.size());
    first_src.push_back(ptrs[k[0]]);
    first_dst.push_back(col);
    if (k.size() > 1) {
      for (int64_t e : k)
        if (findex[(size_t)e] < 0) {
          findex[(size_t)e] = (int64_t)fcols.size();
          fcols.push_back(ptrs[e]);
        }
      for (size_t t = 1; t < k.size(); ++t) {
        SDPSR_REQUIRE(ptrs[k[t] + 1] - ptrs[k[t]] == ptrs[k[0] + 1] - ptrs[k[0]], SDPSR_E_INVALID,
                      "isomorphic eigenspaces must have equal dimension");
        pairs.push_back(Pair{k[0], k[t], col + (int64_t)t});
      }
    }
    col += (int64_t)k.size();
  }
  *nblk = (int64_t)classes.size();
  for (size_t k = 0; k < classes.size(); ++k) blk_sizes[k] = ctx->blk_sizes[k];

  // first column of every block = first eigenvector of the root eigenspace   (:311-314)
  for (size_t f = 0; f < first_src.size(); ++f)
    SDPSR_CUDA(cudaMemcpyAsync(ctx->Qhat + ld * first_dst[f], ctx->Q + ld * first_src[f], (size_t)ld * 8,
                               cudaMemcpyDeviceToDevice, ctx->stream));

  // draw #3 is consumed whether or not it is needed, like the reference (:306)
  SDPSR_REQUIRE(len == ctx->dim, SDPSR_E_INVALID, "length(values) != dim(P) (src/partitions.jl:69)");
  if (!pairs.empty()) {
    SDPSR_TRY(ensure_buffer(ctx, &ctx->T));
    SDPSR_TRY(ensure_buffer(ctx, &ctx->W));
    SDPSR_TRY(fill_into(ctx, r3, len, ctx->X));          // A3
    ctx->x_valid = false;
    const int64_t nf = (int64_t)fcols.size();
    SDPSR_REQUIRE(nf <= n, SDPSR_E_INVALID, "internal: too many first vectors");
    int64_t* d_fcols = nullptr;
    SDPSR_TRY(sdpsr_scratch_t(ctx, 0, (size_t)nf, &d_fcols));
    SDPSR_CUDA(cudaMemcpyAsync(d_fcols, fcols.data(), (size_t)nf * 8, cudaMemcpyHostToDevice, ctx->stream));
    // F = Q[:, fcols]  (into W), V = A3 * F (into T)
    gather_cols_kernel<<<dim3((unsigned)std::min<int64_t>((ld + 255) / 256, 64), (unsigned)nf), 256, 0, ctx->stream>>>(
        ctx->Q, ld, d_fcols, ctx->W);
    count_launch(ctx);
    SDPSR_TRY(sdpsr_gemm_f64(ctx, ctx->X, ld, ctx->W, ld, ctx->T, ld, ld, nf, n, false));
    // per pair: u = Q_j' v_i, w = Q_i' v_j, column = Q_j u / ||w||          (:327-336)
    const int64_t np = (int64_t)pairs.size();
    int64_t maxm = 1;
    std::vector<int64_t> h((size_t)np * 6);
    for (int64_t p = 0; p < np; ++p) {
      const Pair& pr = pairs[(size_t)p];
      const int64_t mlt = ptrs[pr.i + 1] - ptrs[pr.i];
      maxm = std::max(maxm, mlt);
      h[(size_t)p] = ptrs[pr.j];                        // qcol for u  (Q_j)
      h[(size_t)(np + p)] = findex[(size_t)pr.i];       // vcol for u  (v_i)
      h[(size_t)(2 * np + p)] = ptrs[pr.i];             // qcol for w  (Q_i)
      h[(size_t)(3 * np + p)] = findex[(size_t)pr.j];   // vcol for w  (v_j)
      h[(size_t)(4 * np + p)] = mlt;
      h[(size_t)(5 * np + p)] = pr.dst;
    }
    int64_t* d_h = nullptr;
    double* d_uw = nullptr;
    SDPSR_TRY(sdpsr_scratch_t(ctx, 1, h.size(), &d_h));
    SDPSR_TRY(sdpsr_scratch_t(ctx, 2, (size_t)np * (size_t)maxm * 2, &d_uw));
    SDPSR_CUDA(cudaMemcpyAsync(d_h, h.data(), h.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    double* d_u = d_uw;
    double* d_w = d_uw + np * maxm;
    SDPSR_REQUIRE(np <= 65535, SDPSR_E_UNSUPPORTED, "more than 65535 isomorphic eigenspace pairs");
    pair_dot_kernel<<<dim3((unsigned)maxm, (unsigned)np), 256, 0, ctx->stream>>>(ctx->Q, ctx->T, n, ld, d_h, d_h + np,
                                                                                  d_h + 4 * np, maxm, d_u);
    pair_dot_kernel<<<dim3((unsigned)maxm, (unsigned)np), 256, 0, ctx->stream>>>(ctx->Q, ctx->T, n, ld, d_h + 2 * np,
                                                                                  d_h + 3 * np, d_h + 4 * np, maxm, d_w);
    pair_combine_kernel<<<dim3((unsigned)std::min<int64_t>((n + 255) / 256, 64), (unsigned)np), 256, 0, ctx->stream>>>(
        ctx->Q, n, ld, d_h, d_h + 4 * np, maxm, d_u, d_w, d_h + 5 * np, ctx->Qhat);
    count_launch(ctx, 3);
    SDPSR_CUDA(cudaGetLastError());
    SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  // clamptol!.(Q_hat, atol)                                                  (src/diagonalize.jl:39)
  {
    const uint64_t total = (uint64_t)ld * (uint64_t)S;
    const int grid = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)ctx->sm_count * 16);
    clamp_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->Qhat, total, atol);
    count_launch(ctx);
  }
  return finish(ctx);
}

extern "C" int sdpsr_get_qhat(sdpsr_ctx* ctx, double* qhat, int64_t len) {
  CTX_ENTER();
  SDPSR_REQUIRE(ctx->Qhat && qhat, SDPSR_E_STATE, "no Qhat (call sdpsr_irreducible first)");
  SDPSR_REQUIRE(len == ctx->n * ctx->qhat_cols, SDPSR_E_INVALID, "len must be N * sum(blk_sizes)");
  SDPSR_CUDA(cudaMemcpy2DAsync(qhat, (size_t)ctx->n * 8, ctx->Qhat, (size_t)ctx->ld * 8, (size_t)ctx->n * 8,
                               (size_t)ctx->qhat_cols, cudaMemcpyDefault, ctx->stream));
  return finish(ctx);
}

extern "C" int sdpsr_set_qhat(sdpsr_ctx* ctx, const double* qhat, const int64_t* blk_sizes, int64_t nblk) {
  CTX_ENTER();
  SDPSR_REQUIRE(qhat && blk_sizes && nblk >= 1, SDPSR_E_INVALID, "bad arguments");
  int64_t S = 0;
  ctx->blk_sizes.assign(blk_sizes, blk_sizes + nblk);
  for (int64_t k = 0; k < nblk; ++k) {
    SDPSR_REQUIRE(blk_sizes[k] >= 1, SDPSR_E_INVALID, "block sizes must be positive");
    S += blk_sizes[k];
  }
  SDPSR_TRY(sdpsr_scratch_t(ctx, 16, (size_t)ctx->ld * (size_t)S, &ctx->Qhat));
  SDPSR_CUDA(cudaMemsetAsync(ctx->Qhat, 0, (size_t)ctx->ld * (size_t)S * 8, ctx->stream));
  SDPSR_CUDA(cudaMemcpy2DAsync(ctx->Qhat, (size_t)ctx->ld * 8, qhat, (size_t)ctx->n * 8, (size_t)ctx->n * 8, (size_t)S,
                               cudaMemcpyDefault, ctx->stream));
  ctx->qhat_cols = S;
  sdpsr_module_invalidate_qhat(ctx);
  return finish(ctx);
}

// basis_image (src/diagonalize.jl:64-89)
extern "C" int sdpsr_basis_image(sdpsr_ctx* ctx, double atol, double* out, int64_t out_len) {
  CTX_ENTER();
  SDPSR_REQUIRE(ctx->Qhat && out, SDPSR_E_STATE, "no Qhat (call sdpsr_irreducible first)");
  const int64_t n = ctx->n, ld = ctx->ld, S = ctx->qhat_cols, d = ctx->dim;
  int64_t Sq = 0;
  for (int64_t s : ctx->blk_sizes) Sq += s * s;
  SDPSR_REQUIRE(out_len == d * Sq, SDPSR_E_INVALID, "out_len must be dim * sum(s_k^2)");
  if (d == 0) return SDPSR_OK;
  SDPSR_TRY(sdpsr_shard_ensure_full_labels(ctx));
  // Qhat built inside an orbit module (krylov.cu): basis_image is a counting pass
  if (sdpsr_module_basis_available(ctx)) return sdpsr_module_basis_image(ctx, atol, out, out_len);
  KeyTable& t = ctx->tab[ctx->cur];
  if (d * Sq <= BS_MAX_BINS && Sq <= BS_MAX_PAIRS && !(ctx->flags & SDPSR_F_TINY_TABLE)) {
    // ---- few classes, small blocks: one pass over the labels (basis_small_kernel) -----------
    std::vector<int> hp;                    // [pa | pb]
    std::vector<int64_t> poff;
    {
      std::vector<int> pa, pb;
      int64_t colbase = 0, off = 0;
      for (int64_t sz : ctx->blk_sizes) {
        for (int64_t b = 0; b < sz; ++b)
          for (int64_t a = 0; a < sz; ++a) {   // column-major s x s: element (a,b) at a + s*b
            pa.push_back((int)(colbase + a));
            pb.push_back((int)(colbase + b));
            poff.push_back(off + a + sz * b);
          }
        colbase += sz;
        off += sz * sz;
      }
      hp = pa;
      hp.insert(hp.end(), pb.begin(), pb.end());
    }
    const int np = (int)Sq, nbins = (int)(d * Sq);
    const size_t smem = (size_t)nbins * BS_THREADS * sizeof(double);
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(200 * 1024) / std::max<size_t>(smem, 1)));
    // grid: row blocks of 128 x column ranges, about per_sm CTAs per SM in total
    const int64_t rowblocks = (n + BS_THREADS - 1) / BS_THREADS;
    const int64_t want = std::max<int64_t>(1, ((int64_t)ctx->sm_count * per_sm + rowblocks - 1) / rowblocks);
    const int64_t colsplit = std::min<int64_t>(want, std::max<int64_t>(1, n / 64));
    const int64_t colw = (n + colsplit - 1) / colsplit;
    const dim3 grid3((unsigned)rowblocks, (unsigned)((n + colw - 1) / colw));
    const int grid = (int)(grid3.x * grid3.y);
    double *d_qt = nullptr, *d_part = nullptr;
    int* d_pairs = nullptr;
    SDPSR_TRY(sdpsr_scratch_t(ctx, 4, (size_t)n * (size_t)S, &d_qt));
    SDPSR_TRY(sdpsr_scratch_t(ctx, 3, (size_t)grid * nbins, &d_part));
    SDPSR_TRY(sdpsr_scratch_t(ctx, 5, (size_t)2 * np, &d_pairs));
    SDPSR_CUDA(cudaMemcpyAsync(d_pairs, hp.data(), hp.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<double> part((size_t)grid * nbins);
    {
      Timed tm(ctx, SDPSR_K_BASIS, (double)ctx->elems * 4.0);
      qhat_rowmajor_kernel<<<dim3((unsigned)std::min<int64_t>((n + 255) / 256, 64), (unsigned)S), 256, 0, ctx->stream>>>(
          ctx->Qhat, n, ld, S, d_qt);
      int st = SDPSR_OK;
      switch (np) {
#define BS_CASE(K) case K: st = launch_basis_small<K>(ctx, grid3, smem, t.rank, d_qt, S, d_pairs, (int)d, colw, d_part); break;
        BS_CASE(1) BS_CASE(2) BS_CASE(3) BS_CASE(4) BS_CASE(5) BS_CASE(6) BS_CASE(7) BS_CASE(8)
        BS_CASE(9) BS_CASE(10) BS_CASE(11) BS_CASE(12) BS_CASE(13) BS_CASE(14) BS_CASE(15) BS_CASE(16)
#undef BS_CASE
        default: st = ctx->fail(SDPSR_E_INVALID, "basis_small: bad pair count");
      }
      SDPSR_TRY(st);
      count_launch(ctx, 2);
    }
    SDPSR_CUDA(cudaGetLastError());
    SDPSR_CUDA(cudaMemcpyAsync(part.data(), d_part, part.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<double> result((size_t)(d * Sq), 0.0);
    for (int64_t i = 0; i < d; ++i)
      for (int p = 0; p < np; ++p) {
        double sum = 0.0;
        for (int c = 0; c < grid; ++c) sum += part[(size_t)c * nbins + (size_t)(i * np + p)];
        if (std::fabs(sum) < atol) sum = 0.0;                // clamptol!, :85
        result[(size_t)(i * Sq + poff[(size_t)p])] = sum;
      }
    SDPSR_CUDA(cudaMemcpy(out, result.data(), result.size() * 8, cudaMemcpyDefault));
    return finish(ctx);
  }
  // ---- CSR by canonical label (_constraints, :42-50) --------------------------------------
  unsigned long long* d_cnt = nullptr;
  SDPSR_TRY(sdpsr_scratch_t(ctx, 0, (size_t)d + 2, &d_cnt));
  SDPSR_CUDA(cudaMemsetAsync(d_cnt, 0, ((size_t)d + 2) * 8, ctx->stream));
  const unsigned g2 = (unsigned)std::min<int64_t>(n, (int64_t)ctx->sm_count * 8);
  {
    Timed tm(ctx, SDPSR_K_BASIS, (double)ctx->elems * 4.0);
    class_count_kernel<<<g2, 256, 0, ctx->stream>>>(ctx->labels, t.rank, n, ld, d_cnt, d + 1);
    count_launch(ctx);
  }
  std::vector<unsigned long long> cnt((size_t)d + 2);
  SDPSR_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, cnt.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<unsigned long long> start((size_t)d + 2, 0);
  for (int64_t i = 1; i <= d; ++i) start[(size_t)i + 1] = start[(size_t)i] + cnt[(size_t)i];
  const unsigned long long nent = start[(size_t)d + 1];
  SDPSR_CUDA(cudaMemcpyAsync(d_cnt, start.data(), start.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  uint32_t* d_rc = nullptr;
  SDPSR_TRY(sdpsr_scratch_t(ctx, 1, std::max<size_t>(1, (size_t)nent) * 2, &d_rc));
  uint32_t* d_rows = d_rc;
  uint32_t* d_cols = d_rc + nent;
  {
    Timed tm(ctx, SDPSR_K_BASIS, (double)ctx->elems * 12.0);
    class_scatter_kernel<<<g2, 256, 0, ctx->stream>>>(ctx->labels, t.rank, n, ld, d_cnt, d_rows, d_cols, d + 1);
    count_launch(ctx);
  }
  // chunks: class i owns chunks [cstart[i], cstart[i+1])
  std::vector<unsigned long long> cbeg, cend;
  std::vector<int64_t> cstart((size_t)d + 2, 0);
  for (int64_t i = 1; i <= d; ++i) {
    cstart[(size_t)i] = (int64_t)cbeg.size();
    for (unsigned long long e = start[(size_t)i]; e < start[(size_t)i + 1]; e += BI_CHUNK) {
      cbeg.push_back(e);
      cend.push_back(std::min<unsigned long long>(e + BI_CHUNK, start[(size_t)i + 1]));
    }
  }
  cstart[(size_t)d + 1] = (int64_t)cbeg.size();
  const int64_t nchunks = (int64_t)cbeg.size();
  unsigned long long* d_cb = nullptr;
  double* d_part = nullptr;
  double* d_qt = nullptr;
  int* d_pairs = nullptr;
  SDPSR_TRY(sdpsr_scratch_t(ctx, 2, std::max<size_t>(1, (size_t)nchunks) * 2, &d_cb));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 3, std::max<size_t>(1, (size_t)nchunks) * BI_PAIRS, &d_part));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 4, (size_t)n * (size_t)S, &d_qt));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 5, (size_t)2 * BI_PAIRS, &d_pairs));
  SDPSR_CUDA(cudaMemcpyAsync(d_cb, cbeg.data(), (size_t)nchunks * 8, cudaMemcpyHostToDevice, ctx->stream));
  SDPSR_CUDA(cudaMemcpyAsync(d_cb + nchunks, cend.data(), (size_t)nchunks * 8, cudaMemcpyHostToDevice, ctx->stream));
  qhat_rowmajor_kernel<<<dim3((unsigned)std::min<int64_t>((n + 255) / 256, 64), (unsigned)S), 256, 0, ctx->stream>>>(
      ctx->Qhat, n, ld, S, d_qt);
  count_launch(ctx);
  // ---- all output elements (k, a, b) as pairs of Qhat columns, BI_PAIRS at a time ----------
  std::vector<int> pa, pb;
  std::vector<int64_t> poff;   // offset of (k,a,b) inside one class's packed record
  {
    int64_t colbase = 0, off = 0;
    for (int64_t s : ctx->blk_sizes) {
      for (int64_t b = 0; b < s; ++b)
        for (int64_t a = 0; a < s; ++a) {   // column-major s x s: element (a,b) at a + s*b
          pa.push_back((int)(colbase + a));
          pb.push_back((int)(colbase + b));
          poff.push_back(off + a + s * b);
        }
      colbase += s;
      off += s * s;
    }
  }
  std::vector<double> part((size_t)std::max<int64_t>(1, nchunks) * BI_PAIRS);
  std::vector<double> result((size_t)(d * Sq), 0.0);
  int status = SDPSR_OK;
  for (size_t p0 = 0; p0 < pa.size() && status == SDPSR_OK; p0 += BI_PAIRS) {
    const int np = (int)std::min<size_t>(BI_PAIRS, pa.size() - p0);
    int hp[2 * BI_PAIRS] = {0};
    for (int p = 0; p < np; ++p) {
      hp[p] = pa[p0 + p];
      hp[BI_PAIRS + p] = pb[p0 + p];
    }
    cudaMemcpyAsync(d_pairs, hp, sizeof(hp), cudaMemcpyHostToDevice, ctx->stream);
    if (nchunks) {
      Timed tm(ctx, SDPSR_K_BASIS, (double)nent * (8.0 + 16.0 * np));
      basis_partial_kernel<<<(unsigned)nchunks, 256, 0, ctx->stream>>>(d_rows, d_cols, d_cb, d_cb + nchunks, d_qt, S,
                                                                      d_pairs, d_pairs + BI_PAIRS, np, d_part);
      count_launch(ctx);
    }
    cudaMemcpyAsync(part.data(), d_part, (size_t)nchunks * BI_PAIRS * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      status = ctx->fail(SDPSR_E_CUDA, "basis_image kernel failed");
      break;
    }
    for (int64_t i = 1; i <= d; ++i)
      for (int p = 0; p < np; ++p) {
        double s = 0.0;
        for (int64_t c = cstart[(size_t)i]; c < cstart[(size_t)i + 1]; ++c) s += part[(size_t)c * BI_PAIRS + p];
        if (std::fabs(s) < atol) s = 0.0;                  // clamptol!, :85
        result[(size_t)((i - 1) * Sq + poff[p0 + p])] = s;
      }
  }
  SDPSR_TRY(status);
  SDPSR_CUDA(cudaMemcpy(out, result.data(), result.size() * 8, cudaMemcpyDefault));
  return finish(ctx);
}


// cuSOLVER Xsyevd of a small dense symmetric matrix held on the device (the module path of krylov.cu: D x D).
// Eigenvectors overwrite dA; eigenvalues ascending in dvals (device).
int sdpsr_small_syevd(sdpsr_ctx* ctx, double* dA, int64_t n, int64_t lda, double* dvals) {
  if (!ctx->solver) {
    Solver* s = new Solver();
    ctx->solver = s;
    SDPSR_REQUIRE(sapi().ok, SDPSR_E_CUSOLVER, "libcusolver.so.11 could not be loaded");
    SDPSR_REQUIRE(solver_create(&s->h) == CUSOLVER_STATUS_SUCCESS, SDPSR_E_CUSOLVER, "cusolverDnCreate failed");
    SDPSR_REQUIRE(sapi().SetStream(s->h, ctx->stream) == CUSOLVER_STATUS_SUCCESS, SDPSR_E_CUSOLVER,
                  "cusolverDnSetStream failed");
    SDPSR_REQUIRE(sapi().CreateParams(&s->params) == CUSOLVER_STATUS_SUCCESS, SDPSR_E_CUSOLVER,
                  "cusolverDnCreateParams failed");
    SDPSR_CUDA(cudaMalloc(&s->d_vals, (size_t)ctx->n * sizeof(double)));
    SDPSR_CUDA(cudaMalloc(&ctx->solver_info, sizeof(int)));
  }
  Solver* s = reinterpret_cast<Solver*>(ctx->solver);
  size_t wdev = 0, whost = 0;
  cusolverStatus_t st = sapi().Xsyevd_bufferSize(s->h, s->params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n,
                                                 CUDA_R_64F, dA, lda, CUDA_R_64F, dvals, CUDA_R_64F, &wdev, &whost);
  SDPSR_REQUIRE(st == CUSOLVER_STATUS_SUCCESS, SDPSR_E_CUSOLVER, "cusolverDnXsyevd_bufferSize failed (small problem)");
  void* work = nullptr;
  SDPSR_TRY(sdpsr_scratch(ctx, 36, std::max<size_t>(wdev, 256), &work));
  std::vector<unsigned char> hwork(std::max<size_t>(whost, 1));
  {
    Timed tm(ctx, SDPSR_K_EIG, 0.0);
    st = sapi().Xsyevd(s->h, s->params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, dA, lda, CUDA_R_64F,
                       dvals, CUDA_R_64F, work, wdev, hwork.data(), whost, ctx->solver_info);
  }
  SDPSR_REQUIRE(st == CUSOLVER_STATUS_SUCCESS, SDPSR_E_CUSOLVER, "cusolverDnXsyevd failed (small problem)");
  int* hinfo = reinterpret_cast<int*>(ctx->h_pinned) + 64;
  SDPSR_CUDA(cudaMemcpyAsync(hinfo, ctx->solver_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  SDPSR_REQUIRE(*hinfo == 0, SDPSR_E_CUSOLVER, "syevd did not converge (small problem)");
  return SDPSR_OK;
}


// =============================================================================================
// Complex path: diagonalize(ComplexF64, P) of src/diagonalize.jl:25-40 for
// partitions that are not transpose-invariant.  Complex matrices are kept as split re/im planes so
// that every product runs on the real DMMA GEMM (4 real GEMMs with +/- accumulation); the general
// eigensolver is cuSOLVER Xgeev (library, like syevd on the real path).
// =============================================================================================
#include <cuComplex.h>

namespace {

__global__ void interleave_kernel(const double* __restrict__ re, const double* __restrict__ im,
                                  cuDoubleComplex* __restrict__ z, uint64_t total) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
    z[i] = make_cuDoubleComplex(re[i], im[i]);
}

// Q[:, c] = VR[:, order[c]] scaled to unit 2-norm (LAPACK's zgeev convention), split into planes
__global__ void __launch_bounds__(256) gather_unit_columns_kernel(const cuDoubleComplex* __restrict__ VR, int64_t n,
                                                                  int64_t ld, const int64_t* __restrict__ order,
                                                                  double* __restrict__ Qr, double* __restrict__ Qi) {
  __shared__ double ws[8];
  __shared__ double inv;
  const int64_t c = blockIdx.x;
  const cuDoubleComplex* src = VR + ld * order[c];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += src[i].x * src[i].x + src[i].y * src[i].y;
  for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += ws[w];
    inv = t > 0.0 ? 1.0 / sqrt(t) : 0.0;
  }
  __syncthreads();
  for (int64_t i = threadIdx.x; i < ld; i += blockDim.x) {
    const bool in = i < n;
    Qr[i + ld * c] = in ? src[i].x * inv : 0.0;
    Qi[i + ld * c] = in ? src[i].y * inv : 0.0;
  }
}

__global__ void scale_kernel(double* __restrict__ x, uint64_t total, double f) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
    x[i] *= f;
}

// out[pair*maxm + t] = sum_r conj(Q[r, qcol+t]) * V[r, vcol]      (Q_j^H v)
__global__ void __launch_bounds__(256) cpair_dot_kernel(const double* __restrict__ Qr, const double* __restrict__ Qi,
                                                        const double* __restrict__ Vr, const double* __restrict__ Vi,
                                                        int64_t n, int64_t ld, const int64_t* __restrict__ qcol,
                                                        const int64_t* __restrict__ vcol,
                                                        const int64_t* __restrict__ mult, int64_t maxm,
                                                        double* __restrict__ out_r, double* __restrict__ out_i) {
  __shared__ double wr[8], wi[8];
  const int64_t pair = blockIdx.y, t = blockIdx.x;
  if (t >= mult[pair]) return;
  const int64_t qc = ld * (qcol[pair] + t), vc = ld * vcol[pair];
  double sr = 0.0, si = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double ar = Qr[qc + i], ai = Qi[qc + i], br = Vr[vc + i], bi = Vi[vc + i];
    sr += ar * br + ai * bi;       // conj(a) * b
    si += ar * bi - ai * br;
  }
  for (int o = 16; o; o >>= 1) {
    sr += __shfl_down_sync(0xffffffffu, sr, o);
    si += __shfl_down_sync(0xffffffffu, si, o);
  }
  if ((threadIdx.x & 31) == 0) {
    wr[threadIdx.x >> 5] = sr;
    wi[threadIdx.x >> 5] = si;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tr = 0.0, ti = 0.0;
    for (int w = 0; w < 8; ++w) {
      tr += wr[w];
      ti += wi[w];
    }
    out_r[pair * maxm + t] = tr;
    out_i[pair * maxm + t] = ti;
  }
}

// Qhat[:, dst] = Q[:, qcol .. qcol+mult) * (u / ||w||)
__global__ void __launch_bounds__(256) cpair_combine_kernel(const double* __restrict__ Qr, const double* __restrict__ Qi,
                                                            int64_t n, int64_t ld, const int64_t* __restrict__ qcol,
                                                            const int64_t* __restrict__ mult, int64_t maxm,
                                                            const double* __restrict__ ur, const double* __restrict__ ui,
                                                            const double* __restrict__ wr, const double* __restrict__ wi,
                                                            const int64_t* __restrict__ dst, double* __restrict__ Hr,
                                                            double* __restrict__ Hi) {
  const int64_t pair = blockIdx.y;
  const int64_t mlt = mult[pair];
  double nrm2 = 0.0;
  for (int64_t t = 0; t < mlt; ++t)
    nrm2 += wr[pair * maxm + t] * wr[pair * maxm + t] + wi[pair * maxm + t] * wi[pair * maxm + t];
  const double inv = 1.0 / sqrt(nrm2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double sr = 0.0, si = 0.0;
    for (int64_t t = 0; t < mlt; ++t) {
      const double qr = Qr[i + ld * (qcol[pair] + t)], qi = Qi[i + ld * (qcol[pair] + t)];
      const double cr = ur[pair * maxm + t] * inv, ci = ui[pair * maxm + t] * inv;
      sr += qr * cr - qi * ci;
      si += qr * ci + qi * cr;
    }
    Hr[i + ld * dst[pair]] = sr;
    Hi[i + ld * dst[pair]] = si;
  }
}

__global__ void cclamp_kernel(double* __restrict__ xr, double* __restrict__ xi, uint64_t total, double atol) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
    if (hypot(xr[i], xi[i]) < atol) {
      xr[i] = 0.0;
      xi[i] = 0.0;
    }
}

// partial[chunk][p] = sum over entries (r,c) of conj(Qt[r][ca[p]]) * Qt[c][cb[p]]   (Q^H B Q)
__global__ void __launch_bounds__(256) cbasis_partial_kernel(const uint32_t* __restrict__ rows,
                                                             const uint32_t* __restrict__ cols,
                                                             const unsigned long long* __restrict__ chunk_beg,
                                                             const unsigned long long* __restrict__ chunk_end,
                                                             const double* __restrict__ Qtr, const double* __restrict__ Qti,
                                                             int64_t S, const int* __restrict__ ca,
                                                             const int* __restrict__ cb, int npairs,
                                                             double* __restrict__ part_r, double* __restrict__ part_i) {
  __shared__ int sa[BI_PAIRS], sb[BI_PAIRS];
  __shared__ double red[8][BI_PAIRS];
  if (threadIdx.x < BI_PAIRS) {
    sa[threadIdx.x] = threadIdx.x < npairs ? ca[threadIdx.x] : 0;
    sb[threadIdx.x] = threadIdx.x < npairs ? cb[threadIdx.x] : 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned long long e0 = chunk_beg[blockIdx.x], e1 = chunk_end[blockIdx.x];
  for (int part = 0; part < 2; ++part) {            // real part, then imaginary part (keeps registers low)
    double acc[BI_PAIRS];
#pragma unroll
    for (int p = 0; p < BI_PAIRS; ++p) acc[p] = 0.0;
    for (unsigned long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
      const int64_t r = rows[e], c = cols[e];
#pragma unroll
      for (int p = 0; p < BI_PAIRS; ++p)
        if (p < npairs) {
          const double ar = Qtr[S * r + sa[p]], ai = Qti[S * r + sa[p]];
          const double br = Qtr[S * c + sb[p]], bi = Qti[S * c + sb[p]];
          acc[p] += part == 0 ? (ar * br + ai * bi) : (ar * bi - ai * br);
        }
    }
#pragma unroll
    for (int p = 0; p < BI_PAIRS; ++p) {
      double v = acc[p];
      for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) red[wid][p] = v;
    }
    __syncthreads();
    if (threadIdx.x < npairs) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
      (part == 0 ? part_r : part_i)[(int64_t)blockIdx.x * BI_PAIRS + threadIdx.x] = t;
    }
    __syncthreads();
  }
}

int ensure_planes(sdpsr_ctx* ctx) {
  SDPSR_TRY(ensure_buffer(ctx, &ctx->Q));
  SDPSR_TRY(ensure_buffer(ctx, &ctx->W));
  SDPSR_TRY(ensure_buffer(ctx, &ctx->T));
  SDPSR_TRY(ensure_buffer(ctx, &ctx->Xi));
  SDPSR_TRY(ensure_buffer(ctx, &ctx->X2i));
  SDPSR_TRY(ensure_buffer(ctx, &ctx->Qi));
  SDPSR_TRY(ensure_buffer(ctx, &ctx->Wi));
  SDPSR_TRY(ensure_buffer(ctx, &ctx->Ti));
  return SDPSR_OK;
}

// fill(S, r) for complex r (interleaved re/im) into two planes
int cfill_into(sdpsr_ctx* ctx, const double* r, int64_t len, double* dre, double* dim_) {
  SDPSR_REQUIRE(len == ctx->dim, SDPSR_E_INVALID, "length(values) != dim(P) (src/partitions.jl:69)");
  SDPSR_REQUIRE(r != nullptr || len == 0, SDPSR_E_INVALID, "coefficient vector is NULL");
  std::vector<double> re((size_t)len), im((size_t)len);
  for (int64_t i = 0; i < len; ++i) {
    re[(size_t)i] = r[2 * i];
    im[(size_t)i] = r[2 * i + 1];
  }
  SDPSR_TRY(sdpsr_upload_values(ctx, re.data(), len));
  SDPSR_TRY(sdpsr_build_lut(ctx, ctx->d_values, len));
  SDPSR_TRY(sdpsr_materialize_fill(ctx, dre));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  SDPSR_TRY(sdpsr_upload_values(ctx, im.data(), len));
  SDPSR_TRY(sdpsr_build_lut(ctx, ctx->d_values, len));
  SDPSR_TRY(sdpsr_materialize_fill(ctx, dim_));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->x_is_fill = false;
  ctx->x_valid = false;
  return SDPSR_OK;
}

// C = A * B for split-plane complex matrices (M = ld rows, K = n): 4 real DMMA GEMMs
int cgemm(sdpsr_ctx* ctx, const double* Ar, const double* Ai, const double* Br, const double* Bi, double* Cr,
          double* Ci, int64_t ncols) {
  const int64_t ld = ctx->ld, n = ctx->n;
  SDPSR_TRY(sdpsr_gemm_f64(ctx, Ar, ld, Br, ld, Cr, ld, ld, ncols, n, false, false, 0));
  SDPSR_TRY(sdpsr_gemm_f64(ctx, Ai, ld, Bi, ld, Cr, ld, ld, ncols, n, false, false, -1));
  SDPSR_TRY(sdpsr_gemm_f64(ctx, Ar, ld, Bi, ld, Ci, ld, ld, ncols, n, false, false, 0));
  SDPSR_TRY(sdpsr_gemm_f64(ctx, Ai, ld, Br, ld, Ci, ld, ld, ncols, n, false, false, +1));
  return SDPSR_OK;
}

// (Or, Oi) = (Ir, Ii)^H
int cadjoint(sdpsr_ctx* ctx, const double* Ir, const double* Ii, double* Or, double* Oi) {
  const int64_t n = ctx->n, ld = ctx->ld;
  const unsigned nb = (unsigned)((n + 31) / 32);
  if (ld != n) {
    SDPSR_CUDA(cudaMemsetAsync(Or, 0, ctx->elems * 8, ctx->stream));
    SDPSR_CUDA(cudaMemsetAsync(Oi, 0, ctx->elems * 8, ctx->stream));
  }
  transpose_kernel<<<dim3(nb, nb), 256, 0, ctx->stream>>>(Ir, Or, n, ld);
  transpose_kernel<<<dim3(nb, nb), 256, 0, ctx->stream>>>(Ii, Oi, n, ld);
  const int grid = (int)std::min<uint64_t>((ctx->elems + 255) / 256, (uint64_t)ctx->sm_count * 8);
  scale_kernel<<<grid, 256, 0, ctx->stream>>>(Oi, ctx->elems, -1.0);
  count_launch(ctx, 3);
  SDPSR_CUDA(cudaGetLastError());
  return SDPSR_OK;
}

}  // namespace

// A = fill(S, r1) (complex); (vals, Q) = eigen(A), values sorted by (re, im) like Julia's eigsortby
extern "C" int sdpsr_eig_complex(sdpsr_ctx* ctx, const double* r1, int64_t len, double* vals) {
  CTX_ENTER();
  SDPSR_REQUIRE(vals != nullptr, SDPSR_E_INVALID, "vals is NULL");
  SDPSR_REQUIRE(ctx->nranks == 1, SDPSR_E_UNSUPPORTED, "the complex path is single-GPU");
  SDPSR_TRY(ensure_planes(ctx));
  const int64_t n = ctx->n, ld = ctx->ld;
  SDPSR_TRY(cfill_into(ctx, r1, len, ctx->X, ctx->Xi));
  if (!ctx->solver) {
    Solver* s = new Solver();
    ctx->solver = s;
    SDPSR_REQUIRE(sapi().ok, SDPSR_E_CUSOLVER, "libcusolver.so.11 could not be loaded");
    SDPSR_REQUIRE(solver_create(&s->h) == CUSOLVER_STATUS_SUCCESS, SDPSR_E_CUSOLVER, "cusolverDnCreate failed");
    SDPSR_REQUIRE(sapi().SetStream(s->h, ctx->stream) == CUSOLVER_STATUS_SUCCESS, SDPSR_E_CUSOLVER,
                  "cusolverDnSetStream failed");
    SDPSR_REQUIRE(sapi().CreateParams(&s->params) == CUSOLVER_STATUS_SUCCESS, SDPSR_E_CUSOLVER,
                  "cusolverDnCreateParams failed");
    SDPSR_CUDA(cudaMalloc(&s->d_vals, (size_t)ctx->n * sizeof(double)));
    SDPSR_CUDA(cudaMalloc(&ctx->solver_info, sizeof(int)));
  }
  Solver* s = reinterpret_cast<Solver*>(ctx->solver);
  cuDoubleComplex *Az = nullptr, *VR = nullptr, *Wz = nullptr;
  void* dwork = nullptr;
  void* hwork = nullptr;
  int64_t* d_order = nullptr;
  int status = SDPSR_OK;
  auto cleanup = [&]() {
    cudaFree(Az);
    cudaFree(VR);
    cudaFree(Wz);
    cudaFree(dwork);
    cudaFree(d_order);
    free(hwork);
  };
  do {
    if (cudaMalloc(&Az, ctx->elems * 16) != cudaSuccess || cudaMalloc(&VR, ctx->elems * 16) != cudaSuccess ||
        cudaMalloc(&Wz, (size_t)n * 16) != cudaSuccess || cudaMalloc(&d_order, (size_t)n * 8) != cudaSuccess) {
      status = ctx->fail(SDPSR_E_ALLOC, "complex eigensolver buffers");
      break;
    }
    const int grid = (int)std::min<uint64_t>((ctx->elems + 255) / 256, (uint64_t)ctx->sm_count * 8);
    interleave_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->X, ctx->Xi, Az, ctx->elems);
    count_launch(ctx);
    size_t wd = 0, wh = 0;
    if (sapi().Xgeev_bufferSize(s->h, s->params, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_VECTOR, n, CUDA_C_64F,
                                   Az, ld, CUDA_C_64F, Wz, CUDA_C_64F, nullptr, ld, CUDA_C_64F, VR, ld, CUDA_C_64F, &wd,
                                   &wh) != CUSOLVER_STATUS_SUCCESS) {
      status = ctx->fail(SDPSR_E_CUSOLVER, "cusolverDnXgeev_bufferSize failed");
      break;
    }
    if (cudaMalloc(&dwork, std::max<size_t>(wd, 16)) != cudaSuccess) {
      status = ctx->fail(SDPSR_E_ALLOC, "geev workspace");
      break;
    }
    hwork = malloc(std::max<size_t>(wh, 16));
    cusolverStatus_t st;
    {
      Timed tm(ctx, SDPSR_K_EIG, 0.0);
      st = sapi().Xgeev(s->h, s->params, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_VECTOR, n, CUDA_C_64F, Az, ld,
                           CUDA_C_64F, Wz, CUDA_C_64F, nullptr, ld, CUDA_C_64F, VR, ld, CUDA_C_64F, dwork, wd, hwork, wh,
                           ctx->solver_info);
    }
    if (st != CUSOLVER_STATUS_SUCCESS) {
      status = ctx->fail(SDPSR_E_CUSOLVER, "cusolverDnXgeev failed (status " + std::to_string((int)st) + ")");
      break;
    }
    std::vector<double> w((size_t)2 * n);
    int hinfo = 0;
    cudaMemcpyAsync(w.data(), Wz, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(&hinfo, ctx->solver_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      status = ctx->fail(SDPSR_E_CUDA, "geev synchronisation failed");
      break;
    }
    if (hinfo != 0) {
      status = ctx->fail(SDPSR_E_CUSOLVER, "geev did not converge (info = " + std::to_string(hinfo) + ")");
      break;
    }
    std::vector<int64_t> order((size_t)n);
    for (int64_t i = 0; i < n; ++i) order[(size_t)i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
      if (w[2 * a] != w[2 * b]) return w[2 * a] < w[2 * b];
      return w[2 * a + 1] < w[2 * b + 1];
    });
    for (int64_t i = 0; i < n; ++i) {
      vals[2 * i] = w[2 * (size_t)order[(size_t)i]];
      vals[2 * i + 1] = w[2 * (size_t)order[(size_t)i] + 1];
    }
    cudaMemcpyAsync(d_order, order.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream);
    gather_unit_columns_kernel<<<(unsigned)n, 256, 0, ctx->stream>>>(VR, n, ld, d_order, ctx->Q, ctx->Qi);
    count_launch(ctx);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess)
      status = ctx->fail(SDPSR_E_CUDA, "eigenvector gather failed");
  } while (0);
  cleanup();
  SDPSR_TRY(status);
  ctx->have_Q = true;
  ctx->q_complex = true;
  return SDPSR_OK;
}

extern "C" int sdpsr_block_norms_complex(sdpsr_ctx* ctx, const double* r2, int64_t len, const int64_t* ptrs,
                                         int64_t nptr, double* norms) {
  CTX_ENTER();
  SDPSR_REQUIRE(ctx->have_Q && ctx->q_complex, SDPSR_E_STATE, "sdpsr_block_norms_complex must follow sdpsr_eig_complex");
  SDPSR_REQUIRE(ptrs && nptr >= 2 && norms, SDPSR_E_INVALID, "bad eigenspace pointers");
  const int64_t ne = nptr - 1, n = ctx->n, ld = ctx->ld;
  SDPSR_REQUIRE(ptrs[0] == 0 && ptrs[ne] == n, SDPSR_E_INVALID, "ptrs must run from 0 to N");
  std::vector<uint32_t> space((size_t)n), sdim((size_t)ne);
  for (int64_t e = 0; e < ne; ++e) {
    SDPSR_REQUIRE(ptrs[e + 1] > ptrs[e], SDPSR_E_INVALID, "ptrs must be strictly increasing");
    sdim[(size_t)e] = (uint32_t)(ptrs[e + 1] - ptrs[e]);
    for (int64_t a = ptrs[e]; a < ptrs[e + 1]; ++a) space[(size_t)a] = (uint32_t)e;
  }
  SDPSR_TRY(cfill_into(ctx, r2, len, ctx->X, ctx->Xi));                          // A2
  SDPSR_TRY(cgemm(ctx, ctx->X, ctx->Xi, ctx->Q, ctx->Qi, ctx->T, ctx->Ti, n));   // T = A2 Q
  SDPSR_TRY(cadjoint(ctx, ctx->Q, ctx->Qi, ctx->X2, ctx->X2i));                  // Q'
  SDPSR_TRY(cgemm(ctx, ctx->X2, ctx->X2i, ctx->T, ctx->Ti, ctx->W, ctx->Wi, n)); // W = Q' T
  uint32_t *d_space = nullptr, *d_sdim = nullptr;
  unsigned long long* d_norms = nullptr;
  SDPSR_TRY(sdpsr_scratch_t(ctx, 0, (size_t)n, &d_space));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 1, (size_t)ne, &d_sdim));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 2, (size_t)ne * ne, &d_norms));
  SDPSR_CUDA(cudaMemcpyAsync(d_space, space.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
  SDPSR_CUDA(cudaMemcpyAsync(d_sdim, sdim.data(), (size_t)ne * 4, cudaMemcpyHostToDevice, ctx->stream));
  SDPSR_CUDA(cudaMemsetAsync(d_norms, 0, (size_t)ne * ne * 8, ctx->stream));
  block_max_kernel<<<dim3((unsigned)std::min<int64_t>((n + 255) / 256, 64), (unsigned)n), 256, 0, ctx->stream>>>(
      ctx->W, ctx->Wi, n, ld, d_space, d_sdim, ne, d_norms);
  count_launch(ctx);
  SDPSR_CUDA(cudaGetLastError());
  SDPSR_CUDA(cudaMemcpyAsync(norms, d_norms, (size_t)ne * ne * 8, cudaMemcpyDefault, ctx->stream));
  SDPSR_TRY(finish(ctx));
  for (int64_t j = 0; j < ne; ++j)
    for (int64_t i = 0; i < j; ++i) norms[j + ne * i] = norms[i + ne * j];
  return SDPSR_OK;
}

extern "C" int sdpsr_irreducible_complex(sdpsr_ctx* ctx, const double* r3, int64_t len, const int64_t* ptrs,
                                         int64_t nptr, const int64_t* kroot, double atol, int64_t* blk_sizes,
                                         int64_t* nblk) {
  CTX_ENTER();
  SDPSR_REQUIRE(ctx->have_Q && ctx->q_complex, SDPSR_E_STATE, "sdpsr_irreducible_complex must follow sdpsr_eig_complex");
  SDPSR_REQUIRE(ptrs && nptr >= 2 && kroot && blk_sizes && nblk, SDPSR_E_INVALID, "bad arguments");
  const int64_t ne = nptr - 1, n = ctx->n, ld = ctx->ld;
  std::vector<std::vector<int64_t>> classes;
  std::vector<int64_t> class_of((size_t)ne, -1);
  for (int64_t e = 0; e < ne; ++e) {
    const int64_t r = kroot[e];
    SDPSR_REQUIRE(r >= 0 && r <= e && kroot[r] == r, SDPSR_E_INVALID, "kroot[e] must be the smallest member of its class");
    if (r == e) {
      class_of[(size_t)e] = (int64_t)classes.size();
      classes.emplace_back();
    }
    classes[(size_t)class_of[(size_t)r]].push_back(e);
    class_of[(size_t)e] = class_of[(size_t)r];
  }
  int64_t S = 0;
  for (auto& k : classes) S += (int64_t)k.size();
  SDPSR_TRY(sdpsr_scratch_t(ctx, 16, (size_t)ld * (size_t)S, &ctx->Qhat));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 17, (size_t)ld * (size_t)S, &ctx->Qhat_i));
  SDPSR_CUDA(cudaMemsetAsync(ctx->Qhat, 0, (size_t)ld * (size_t)S * 8, ctx->stream));
  SDPSR_CUDA(cudaMemsetAsync(ctx->Qhat_i, 0, (size_t)ld * (size_t)S * 8, ctx->stream));
  ctx->qhat_cols = S;
  ctx->blk_sizes.clear();
  std::vector<int64_t> fcols, findex((size_t)ne, -1);
  struct Pair { int64_t i, j, dst; };
  std::vector<Pair> pairs;
  int64_t col = 0;
  for (auto& k : classes) {
    ctx->blk_sizes.push_back((int64_t)k.size());
    SDPSR_CUDA(cudaMemcpyAsync(ctx->Qhat + ld * col, ctx->Q + ld * ptrs[k[0]], (size_t)ld * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    SDPSR_CUDA(cudaMemcpyAsync(ctx->Qhat_i + ld * col, ctx->Qi + ld * ptrs[k[0]], (size_t)ld * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    if (k.size() > 1) {
      for (int64_t e : k)
        if (findex[(size_t)e] < 0) {
          findex[(size_t)e] = (int64_t)fcols.size();
          fcols.push_back(ptrs[e]);
        }
      for (size_t t = 1; t < k.size(); ++t) {
        SDPSR_REQUIRE(ptrs[k[t] + 1] - ptrs[k[t]] == ptrs[k[0] + 1] - ptrs[k[0]], SDPSR_E_INVALID,
                      "isomorphic eigenspaces must have equal dimension");
        pairs.push_back(Pair{k[0], k[t], col + (int64_t)t});
      }
    }
    col += (int64_t)k.size();
  }
  *nblk = (int64_t)classes.size();
  for (size_t k = 0; k < classes.size(); ++k) blk_sizes[k] = ctx->blk_sizes[k];
  SDPSR_REQUIRE(len == ctx->dim, SDPSR_E_INVALID, "length(values) != dim(P) (src/partitions.jl:69)");
  if (!pairs.empty()) {
    const int64_t nf = (int64_t)fcols.size(), np = (int64_t)pairs.size();
    SDPSR_REQUIRE(np <= 65535, SDPSR_E_UNSUPPORTED, "more than 65535 isomorphic eigenspace pairs");
    SDPSR_TRY(cfill_into(ctx, r3, len, ctx->X, ctx->Xi));                      // A3
    int64_t* d_fcols = nullptr;
    SDPSR_CUDA(cudaMalloc(&d_fcols, (size_t)nf * 8));
    SDPSR_CUDA(cudaMemcpyAsync(d_fcols, fcols.data(), (size_t)nf * 8, cudaMemcpyHostToDevice, ctx->stream));
    const dim3 gg((unsigned)std::min<int64_t>((ld + 255) / 256, 64), (unsigned)nf);
    gather_cols_kernel<<<gg, 256, 0, ctx->stream>>>(ctx->Q, ld, d_fcols, ctx->W);       // F (re)
    gather_cols_kernel<<<gg, 256, 0, ctx->stream>>>(ctx->Qi, ld, d_fcols, ctx->Wi);     // F (im)
    count_launch(ctx, 2);
    // V = A3 F (into T) ; V' = A3^H F (into Q'-scratch X2 after forming A3^H in place of A3)
    SDPSR_TRY(cgemm(ctx, ctx->X, ctx->Xi, ctx->W, ctx->Wi, ctx->T, ctx->Ti, nf));
    double *Vpr = nullptr, *Vpi = nullptr, *Ahr = nullptr, *Ahi = nullptr;
    SDPSR_CUDA(cudaMalloc(&Vpr, ctx->elems * 8));
    SDPSR_CUDA(cudaMalloc(&Vpi, ctx->elems * 8));
    Ahr = ctx->X2;
    Ahi = ctx->X2i;
    SDPSR_TRY(cadjoint(ctx, ctx->X, ctx->Xi, Ahr, Ahi));
    SDPSR_TRY(cgemm(ctx, Ahr, Ahi, ctx->W, ctx->Wi, Vpr, Vpi, nf));
    int64_t maxm = 1;
    std::vector<int64_t> h((size_t)np * 6);
    for (int64_t p = 0; p < np; ++p) {
      const Pair& pr = pairs[(size_t)p];
      const int64_t mlt = ptrs[pr.i + 1] - ptrs[pr.i];
      maxm = std::max(maxm, mlt);
      h[(size_t)p] = ptrs[pr.j];                        // Q_j for u
      h[(size_t)(np + p)] = findex[(size_t)pr.i];       // A3^H q_i1
      h[(size_t)(2 * np + p)] = ptrs[pr.i];             // Q_i for w
      h[(size_t)(3 * np + p)] = findex[(size_t)pr.j];   // A3 q_j1
      h[(size_t)(4 * np + p)] = mlt;
      h[(size_t)(5 * np + p)] = pr.dst;
    }
    int64_t* d_h = nullptr;
    double* d_uw = nullptr;
    SDPSR_CUDA(cudaMalloc(&d_h, h.size() * 8));
    SDPSR_CUDA(cudaMalloc(&d_uw, (size_t)np * (size_t)maxm * 4 * 8));
    SDPSR_CUDA(cudaMemcpyAsync(d_h, h.data(), h.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    double* ur = d_uw;
    double* ui = ur + np * maxm;
    double* wr = ui + np * maxm;
    double* wi = wr + np * maxm;
    cpair_dot_kernel<<<dim3((unsigned)maxm, (unsigned)np), 256, 0, ctx->stream>>>(ctx->Q, ctx->Qi, Vpr, Vpi, n, ld, d_h,
                                                                                   d_h + np, d_h + 4 * np, maxm, ur, ui);
    cpair_dot_kernel<<<dim3((unsigned)maxm, (unsigned)np), 256, 0, ctx->stream>>>(
        ctx->Q, ctx->Qi, ctx->T, ctx->Ti, n, ld, d_h + 2 * np, d_h + 3 * np, d_h + 4 * np, maxm, wr, wi);
    cpair_combine_kernel<<<dim3((unsigned)std::min<int64_t>((n + 255) / 256, 64), (unsigned)np), 256, 0, ctx->stream>>>(
        ctx->Q, ctx->Qi, n, ld, d_h, d_h + 4 * np, maxm, ur, ui, wr, wi, d_h + 5 * np, ctx->Qhat, ctx->Qhat_i);
    count_launch(ctx, 3);
    SDPSR_CUDA(cudaGetLastError());
    SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_h);
    cudaFree(d_uw);
    cudaFree(d_fcols);
    cudaFree(Vpr);
    cudaFree(Vpi);
  }
  {
    const uint64_t total = (uint64_t)ld * (uint64_t)S;
    const int grid = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)ctx->sm_count * 16);
    cclamp_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->Qhat, ctx->Qhat_i, total, atol);
    count_launch(ctx);
  }
  return finish(ctx);
}

// out: interleaved complex, packed [i][k] s_k x s_k column-major; out_len counts complex numbers
extern "C" int sdpsr_basis_image_complex(sdpsr_ctx* ctx, double atol, double* out, int64_t out_len) {
  CTX_ENTER();
  SDPSR_TRY(sdpsr_shard_ensure_full_labels(ctx));
  SDPSR_REQUIRE(ctx->Qhat && ctx->Qhat_i && out, SDPSR_E_STATE, "no complex Qhat (call sdpsr_irreducible_complex first)");
  const int64_t n = ctx->n, ld = ctx->ld, S = ctx->qhat_cols, d = ctx->dim;
  int64_t Sq = 0;
  for (int64_t s : ctx->blk_sizes) Sq += s * s;
  SDPSR_REQUIRE(out_len == d * Sq, SDPSR_E_INVALID, "out_len must be dim * sum(s_k^2)");
  if (d == 0) return SDPSR_OK;
  KeyTable& t = ctx->tab[ctx->cur];
  unsigned long long* d_cnt = nullptr;
  SDPSR_CUDA(cudaMalloc(&d_cnt, ((size_t)d + 2) * 8));
  SDPSR_CUDA(cudaMemsetAsync(d_cnt, 0, ((size_t)d + 2) * 8, ctx->stream));
  const unsigned g2 = (unsigned)std::min<int64_t>(n, (int64_t)ctx->sm_count * 8);
  class_count_kernel<<<g2, 256, 0, ctx->stream>>>(ctx->labels, t.rank, n, ld, d_cnt, d + 1);
  std::vector<unsigned long long> cnt((size_t)d + 2);
  SDPSR_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, cnt.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<unsigned long long> start((size_t)d + 2, 0);
  for (int64_t i = 1; i <= d; ++i) start[(size_t)i + 1] = start[(size_t)i] + cnt[(size_t)i];
  const unsigned long long nent = start[(size_t)d + 1];
  SDPSR_CUDA(cudaMemcpyAsync(d_cnt, start.data(), start.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  uint32_t* d_rc = nullptr;
  SDPSR_CUDA(cudaMalloc(&d_rc, std::max<size_t>(1, (size_t)nent) * 8));
  uint32_t* d_rows = d_rc;
  uint32_t* d_cols = d_rc + nent;
  class_scatter_kernel<<<g2, 256, 0, ctx->stream>>>(ctx->labels, t.rank, n, ld, d_cnt, d_rows, d_cols, d + 1);
  count_launch(ctx, 2);
  std::vector<unsigned long long> cbeg, cend;
  std::vector<int64_t> cstart((size_t)d + 2, 0);
  for (int64_t i = 1; i <= d; ++i) {
    cstart[(size_t)i] = (int64_t)cbeg.size();
    for (unsigned long long e = start[(size_t)i]; e < start[(size_t)i + 1]; e += BI_CHUNK) {
      cbeg.push_back(e);
      cend.push_back(std::min<unsigned long long>(e + BI_CHUNK, start[(size_t)i + 1]));
    }
  }
  cstart[(size_t)d + 1] = (int64_t)cbeg.size();
  const int64_t nchunks = (int64_t)cbeg.size();
  unsigned long long* d_cb = nullptr;
  double *d_part = nullptr, *d_qtr = nullptr, *d_qti = nullptr;
  int* d_pairs = nullptr;
  SDPSR_CUDA(cudaMalloc(&d_cb, std::max<size_t>(1, (size_t)nchunks) * 16));
  SDPSR_CUDA(cudaMalloc(&d_part, std::max<size_t>(1, (size_t)nchunks) * BI_PAIRS * 16));
  SDPSR_CUDA(cudaMalloc(&d_qtr, (size_t)n * (size_t)S * 8));
  SDPSR_CUDA(cudaMalloc(&d_qti, (size_t)n * (size_t)S * 8));
  SDPSR_CUDA(cudaMalloc(&d_pairs, 2 * BI_PAIRS * sizeof(int)));
  SDPSR_CUDA(cudaMemcpyAsync(d_cb, cbeg.data(), (size_t)nchunks * 8, cudaMemcpyHostToDevice, ctx->stream));
  SDPSR_CUDA(cudaMemcpyAsync(d_cb + nchunks, cend.data(), (size_t)nchunks * 8, cudaMemcpyHostToDevice, ctx->stream));
  const dim3 gq((unsigned)std::min<int64_t>((n + 255) / 256, 64), (unsigned)S);
  qhat_rowmajor_kernel<<<gq, 256, 0, ctx->stream>>>(ctx->Qhat, n, ld, S, d_qtr);
  qhat_rowmajor_kernel<<<gq, 256, 0, ctx->stream>>>(ctx->Qhat_i, n, ld, S, d_qti);
  count_launch(ctx, 2);
  std::vector<int> pa, pb;
  std::vector<int64_t> poff;
  {
    int64_t colbase = 0, off = 0;
    for (int64_t s : ctx->blk_sizes) {
      for (int64_t b = 0; b < s; ++b)
        for (int64_t a = 0; a < s; ++a) {
          pa.push_back((int)(colbase + a));
          pb.push_back((int)(colbase + b));
          poff.push_back(off + a + s * b);
        }
      colbase += s;
      off += s * s;
    }
  }
  const size_t pstride = (size_t)std::max<int64_t>(1, nchunks) * BI_PAIRS;
  std::vector<double> part(2 * pstride);
  std::vector<double> result((size_t)(2 * d * Sq), 0.0);
  int status = SDPSR_OK;
  for (size_t p0 = 0; p0 < pa.size() && status == SDPSR_OK; p0 += BI_PAIRS) {
    const int np = (int)std::min<size_t>(BI_PAIRS, pa.size() - p0);
    int hp[2 * BI_PAIRS] = {0};
    for (int p = 0; p < np; ++p) {
      hp[p] = pa[p0 + p];
      hp[BI_PAIRS + p] = pb[p0 + p];
    }
    cudaMemcpyAsync(d_pairs, hp, sizeof(hp), cudaMemcpyHostToDevice, ctx->stream);
    if (nchunks) {
      cbasis_partial_kernel<<<(unsigned)nchunks, 256, 0, ctx->stream>>>(d_rows, d_cols, d_cb, d_cb + nchunks, d_qtr, d_qti,
                                                                       S, d_pairs, d_pairs + BI_PAIRS, np, d_part,
                                                                       d_part + pstride);
      count_launch(ctx);
    }
    cudaMemcpyAsync(part.data(), d_part, 2 * pstride * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      status = ctx->fail(SDPSR_E_CUDA, "complex basis_image kernel failed");
      break;
    }
    for (int64_t i = 1; i <= d; ++i)
      for (int p = 0; p < np; ++p) {
        double sr = 0.0, si = 0.0;
        for (int64_t c = cstart[(size_t)i]; c < cstart[(size_t)i + 1]; ++c) {
          sr += part[(size_t)c * BI_PAIRS + p];
          si += part[pstride + (size_t)c * BI_PAIRS + p];
        }
        if (std::hypot(sr, si) < atol) sr = si = 0.0;
        const size_t o = (size_t)((i - 1) * Sq + poff[p0 + p]);
        result[2 * o] = sr;
        result[2 * o + 1] = si;
      }
  }
  cudaFree(d_cnt);
  cudaFree(d_rc);
  cudaFree(d_cb);
  cudaFree(d_part);
  cudaFree(d_qtr);
  cudaFree(d_qti);
  cudaFree(d_pairs);
  SDPSR_TRY(status);
  SDPSR_CUDA(cudaMemcpy(out, result.data(), result.size() * 8, cudaMemcpyDefault));
  return finish(ctx);
}

// interleaved complex N x S (column-major)
extern "C" int sdpsr_get_qhat_complex(sdpsr_ctx* ctx, double* qhat, int64_t len) {
  CTX_ENTER();
  SDPSR_REQUIRE(ctx->Qhat && ctx->Qhat_i && qhat, SDPSR_E_STATE, "no complex Qhat");
  SDPSR_REQUIRE(len == ctx->n * ctx->qhat_cols, SDPSR_E_INVALID, "len must be N * sum(blk_sizes) complex numbers");
  const size_t cnt = (size_t)ctx->n * (size_t)ctx->qhat_cols;
  std::vector<double> re(cnt), im(cnt);
  SDPSR_CUDA(cudaMemcpy2DAsync(re.data(), (size_t)ctx->n * 8, ctx->Qhat, (size_t)ctx->ld * 8, (size_t)ctx->n * 8,
                               (size_t)ctx->qhat_cols, cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaMemcpy2DAsync(im.data(), (size_t)ctx->n * 8, ctx->Qhat_i, (size_t)ctx->ld * 8, (size_t)ctx->n * 8,
                               (size_t)ctx->qhat_cols, cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_TRY(finish(ctx));
  for (size_t i = 0; i < cnt; ++i) {
    qhat[2 * i] = re[i];
    qhat[2 * i + 1] = im[i];
  }
  return SDPSR_OK;
}
