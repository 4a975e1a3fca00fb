// krylov.cu -- blockDiagonalize without a dense eigendecomposition, for partitions whose generic
// element has FEW distinct eigenvalues (association schemes, small coherent algebras).
//
// Same outputs as the dense path of blockdiag.cu, i.e. as the reference's
//   eigen_decomposition        src/eigen_decomposition.jl:236-273
//   isomorphism_partition      src/eigen_decomposition.jl:201-219
//   irreducible_decomposition  src/eigen_decomposition.jl:295-348
// but the work is O(ne) matrix-vector products instead of an O(N^3) `eigen`:
//
//  * A1 = fill(S, r1) has ne = sum(s_k) distinct eigenvalues (one per eigenspace E_i).  Lanczos with
//    full reorthogonalisation from a generic start vector breaks down after exactly ne steps; the
//    Ritz pairs at breakdown are exact eigenpairs: one unit vector y_i in every E_i.  The reference
//    only ever uses ONE vector of the root eigenspace of a class ("first column", :311-314) plus
//    projections P_j A3 y_i (:327-336), and any unit vector of E_i gives the same blocks
//    (SURVEY.md A.6), so one vector per eigenspace is all that is needed.
//  * dim E_i = trace of the spectral projector P_i.  P_i lies in the algebra, so its diagonal is
//    constant on every diagonal class; (P_i)_rr = the Gauss-quadrature weight of eigenvalue i in a
//    Lanczos run started at the unit vector e_r.  One short run per diagonal class.
//  * isomorphism test (:203-217): ||P_j A2 y_i|| -- the quadrature weights of a Lanczos run
//    started at A2 y_i -- plays the role of max|Q_i' A2 Q_j|; same equal-dimension mask; the Otsu
//    threshold, union-find and consistency check stay in the host language as before.
//  * Qhat columns (:327-336): the Ritz vectors of a |K_i|-step Lanczos run started at A3 y_i are
//    +-P_j A3 y_i / ||.||; the sign is fixed by the first component of the tridiagonal eigenvector.
//
// The matrix is never materialised: y = A v is evaluated from the u32 label matrix and the
// coefficient LUT (4 B/entry of HBM traffic instead of 8).
//
// Rounding noise in directions of a degenerate eigenspace that are orthogonal to the Krylov space
// is amplified as the Ritz values converge, so a CLEAN breakdown (beta <= tol * ||A||) is only
// observed for small ne (about a dozen).  The path is self-validating: without a clean breakdown,
// with non-integer multiplicities or unmatched Ritz values it returns SDPSR_E_KRYLOV and the
// caller runs the dense path (sdpsr_eig ...) with the same coefficient vectors.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "sdpsr_internal.cuh"

namespace {

constexpr int KR_MAX = 48;       // hard cap on Lanczos steps / eigenspaces
constexpr int KR_MAX_DIAG = 32;  // hard cap on diagonal classes (one run each)

struct Krylov {
  int ne = 0;
  bool ready = false;
  double anorm = 0.0;
  double tol = 1e-10;
  // Breakdown threshold of the follow-up runs (isomorphism, class): their start vectors A y_i carry the
  // residual of y_i (up to tol * ||A||, amplified by ||A2||/gap), so demanding tol again fails about one
  // draw in four at N = 16384.  1e-7 is the tolerance the Ritz values are matched with anyway.
  double soft_tol = 1e-7;
  std::vector<double> th;       // distinct eigenvalues of A1, ascending
  std::vector<int64_t> mult;    // dim E_i
  double* lut1 = nullptr;       // device: coefficient LUT of A1 by provisional id   (scratch 20)
  double* V = nullptr;          // device: ld x (KR_MAX+1) Lanczos basis             (scratch 21)
  double* Y = nullptr;          // device: ld x ne, one unit eigenvector per E_i     (scratch 22)
  double* U = nullptr;          // device: ld x ne work                              (scratch 23)
  double* w = nullptr;          // device: ld work vector                            (scratch 24)
  double* small = nullptr;      // device: dots / coefficients / S matrix            (scratch 25)
};

// ---------------------------------------------------------------------------------------------
// y[j] = sum_i lut[lab[i + ld*j]] * v[i]   (A symmetric, so A v = A' v: a column dot product).
// One warp per pair of columns: the v chunk is loaded once and used for both columns.  Padding
// rows hold label 0 (lut[0] == 0) and v == 0.  Fixed summation order -> bit-reproducible.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) labmv_kernel(const uint32_t* __restrict__ lab, const double* __restrict__ lut,
                                                    const double* __restrict__ v, double* __restrict__ y, int64_t n,
                                                    int64_t ld) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j0 = warp * 2; j0 < n; j0 += nwarps * 2) {
    const bool two = j0 + 1 < n;
    const uint32_t* c0 = lab + ld * j0;
    const uint32_t* c1 = lab + ld * (two ? j0 + 1 : j0);
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    for (int64_t i = (int64_t)lane * 4; i < ld; i += 128) {
      const uint4 l0 = __ldcs(reinterpret_cast<const uint4*>(c0 + i));
      const uint4 l1 = __ldcs(reinterpret_cast<const uint4*>(c1 + i));
      const double2 v0 = __ldg(reinterpret_cast<const double2*>(v + i));
      const double2 v1 = __ldg(reinterpret_cast<const double2*>(v + i + 2));
      a0 = fma(__ldg(lut + l0.x), v0.x, a0);
      a1 = fma(__ldg(lut + l0.y), v0.y, a1);
      a0 = fma(__ldg(lut + l0.z), v1.x, a0);
      a1 = fma(__ldg(lut + l0.w), v1.y, a1);
      b0 = fma(__ldg(lut + l1.x), v0.x, b0);
      b1 = fma(__ldg(lut + l1.y), v0.y, b1);
      b0 = fma(__ldg(lut + l1.z), v1.x, b0);
      b1 = fma(__ldg(lut + l1.w), v1.y, b1);
    }
    double a = a0 + a1, b = b0 + b1;
    for (int o = 16; o; o >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, o);
      b += __shfl_down_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      y[j0] = a;
      if (two) y[j0 + 1] = b;
    }
  }
}

// out[c] = dot(V[:, c], w), c < k : one CTA per column, fixed-order reduction
__global__ void __launch_bounds__(256) kr_dots_kernel(const double* __restrict__ V, int64_t ldv,
                                                      const double* __restrict__ w, int64_t n,
                                                      double* __restrict__ out) {
  __shared__ double ws[8];
  const double* col = V + ldv * blockIdx.x;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s = fma(col[i], w[i], s);
  for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int q = 0; q < 8; ++q) t += ws[q];
    out[blockIdx.x] = t;
  }
}

// w[i] -= sum_c h[c] * V[i + ldv*c]
__global__ void __launch_bounds__(256) kr_axpy_kernel(const double* __restrict__ V, int64_t ldv, int k,
                                                      const double* __restrict__ h, double* __restrict__ w,
                                                      int64_t n) {
  __shared__ double sh[KR_MAX + 1];
  for (int c = threadIdx.x; c < k; c += blockDim.x) sh[c] = h[c];
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < k; ++c) s = fma(sh[c], V[i + ldv * c], s);
    w[i] -= s;
  }
}

// dst[i] = src[i] * s  (i < n), 0 in the padding rows
__global__ void kr_scale_kernel(const double* __restrict__ src, double s, double* __restrict__ dst, int64_t n,
                                int64_t ld) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ld; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = i < n ? src[i] * s : 0.0;
}

__global__ void kr_unit_kernel(double* __restrict__ dst, int64_t ld, int64_t r) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ld; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = i == r ? 1.0 : 0.0;
}

// dst[:, t] = sgn[t] * V[:, 0..k) * S[:, sel[t]]   (S is k x k column-major; tiny)
__global__ void __launch_bounds__(256) kr_ritz_kernel(const double* __restrict__ V, int64_t ldv, int k,
                                                      const double* __restrict__ S, const int* __restrict__ sel,
                                                      const double* __restrict__ sgn, int nsel,
                                                      double* __restrict__ dst, int64_t lddst, int64_t n) {
  extern __shared__ double sS[];   // k * nsel
  for (int t = threadIdx.x; t < k * nsel; t += blockDim.x) {
    const int c = t / k, r = t - c * k;
    sS[t] = S[r + k * sel[c]] * sgn[c];
  }
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    for (int c = 0; c < nsel; ++c) {
      double s = 0.0;
      for (int r = 0; r < k; ++r) s = fma(V[i + ldv * r], sS[r + k * c], s);
      dst[i + lddst * c] = s;
    }
  }
}

// canonical labels of the diagonal
__global__ void kr_diag_kernel(const uint32_t* __restrict__ lab, const uint32_t* __restrict__ rank, int64_t n,
                               int64_t ld, uint32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = rank[lab[i + ld * i]];
}

__global__ void kr_clamp_kernel(double* __restrict__ x, uint64_t total, double atol) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
    if (fabs(x[i]) < atol) x[i] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// host: eigen-decomposition of a k x k symmetric tridiagonal matrix by cyclic Jacobi (k <= 48;
// accurate small eigenvector components matter more here than speed).  Ascending eigenvalues;
// Z column-major, Z[:, c] the eigenvector of val[c].
// ---------------------------------------------------------------------------------------------
void tridiag_eig(const std::vector<double>& al, const std::vector<double>& be, int k, std::vector<double>& val,
                 std::vector<double>& Z) {
  std::vector<double> A((size_t)k * k, 0.0);
  Z.assign((size_t)k * k, 0.0);
  for (int i = 0; i < k; ++i) {
    A[(size_t)i + (size_t)k * i] = al[(size_t)i];
    Z[(size_t)i + (size_t)k * i] = 1.0;
    if (i + 1 < k) A[(size_t)i + 1 + (size_t)k * i] = A[(size_t)i + (size_t)k * (i + 1)] = be[(size_t)i];
  }
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int p = 0; p < k; ++p) {
      diag += A[(size_t)p + (size_t)k * p] * A[(size_t)p + (size_t)k * p];
      for (int q = p + 1; q < k; ++q) off += A[(size_t)p + (size_t)k * q] * A[(size_t)p + (size_t)k * q];
    }
    if (off <= 1e-34 * (diag + off) || off == 0.0) break;
    for (int p = 0; p < k; ++p)
      for (int q = p + 1; q < k; ++q) {
        const double apq = A[(size_t)p + (size_t)k * q];
        if (apq == 0.0) continue;
        const double app = A[(size_t)p + (size_t)k * p], aqq = A[(size_t)q + (size_t)k * q];
        const double tau = (aqq - app) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int r = 0; r < k; ++r) {          // columns p, q
          const double arp = A[(size_t)r + (size_t)k * p], arq = A[(size_t)r + (size_t)k * q];
          A[(size_t)r + (size_t)k * p] = c * arp - s * arq;
          A[(size_t)r + (size_t)k * q] = s * arp + c * arq;
        }
        for (int r = 0; r < k; ++r) {          // rows p, q
          const double apr = A[(size_t)p + (size_t)k * r], aqr = A[(size_t)q + (size_t)k * r];
          A[(size_t)p + (size_t)k * r] = c * apr - s * aqr;
          A[(size_t)q + (size_t)k * r] = s * apr + c * aqr;
        }
        for (int r = 0; r < k; ++r) {
          const double zrp = Z[(size_t)r + (size_t)k * p], zrq = Z[(size_t)r + (size_t)k * q];
          Z[(size_t)r + (size_t)k * p] = c * zrp - s * zrq;
          Z[(size_t)r + (size_t)k * q] = s * zrp + c * zrq;
        }
      }
  }
  std::vector<int> order((size_t)k);
  for (int i = 0; i < k; ++i) order[(size_t)i] = i;
  std::sort(order.begin(), order.end(),
            [&](int a, int b) { return A[(size_t)a + (size_t)k * a] < A[(size_t)b + (size_t)k * b]; });
  val.resize((size_t)k);
  std::vector<double> Zs((size_t)k * k);
  for (int c = 0; c < k; ++c) {
    const int src = order[(size_t)c];
    val[(size_t)c] = A[(size_t)src + (size_t)k * src];
    for (int r = 0; r < k; ++r) Zs[(size_t)r + (size_t)k * c] = Z[(size_t)r + (size_t)k * src];
  }
  Z.swap(Zs);
}

struct LanczosOut {
  std::vector<double> alpha, beta;   // beta[k] couples steps k and k+1
  int steps = 0;
  bool breakdown = false;
  double start_norm = 0.0;
  double scale = 0.0;                // running estimate of ||A||
};

int matvec(sdpsr_ctx* ctx, const double* lut, const double* v, double* y) {
  Timed tm(ctx, SDPSR_K_KRYLOV, (double)ctx->elems * 4.0);
  const int64_t warps = (ctx->n + 1) / 2;
  const int grid = (int)std::min<int64_t>((warps + 7) / 8, (int64_t)ctx->sm_count * 8);
  labmv_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->labels, lut, v, y, ctx->n, ctx->ld);
  count_launch(ctx);
  SDPSR_CUDA(cudaGetLastError());
  return SDPSR_OK;
}

int vec_grid(sdpsr_ctx* ctx) { return (int)std::min<int64_t>((ctx->ld + 255) / 256, (int64_t)ctx->sm_count * 4); }

// Lanczos with full (twice classical Gram-Schmidt) reorthogonalisation on A = lut[labels].
// start: device vector (ld doubles, zero padding).  Basis in kr.V.  Stops at the first
// beta <= tol * (anorm > 0 ? anorm : running scale) ("breakdown") or after kmax steps.
int lanczos(sdpsr_ctx* ctx, Krylov& kr, const double* lut, const double* start, int kmax, double tol, double anorm,
            LanczosOut& out) {
  const int64_t n = ctx->n, ld = ctx->ld;
  double* hs = reinterpret_cast<double*>(ctx->h_pinned) + 128;   // pinned: 2*(KR_MAX+1)+2 doubles fit in 4 KB
  out = LanczosOut();
  const int vg = vec_grid(ctx);
  kr_dots_kernel<<<1, 256, 0, ctx->stream>>>(start, ld, start, n, kr.small);
  count_launch(ctx);
  SDPSR_CUDA(cudaMemcpyAsync(hs, kr.small, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  out.start_norm = std::sqrt(hs[0]);
  if (!(out.start_norm > 0.0) || !std::isfinite(out.start_norm)) {   // zero start vector: empty Krylov space
    out.breakdown = true;
    return SDPSR_OK;
  }
  kr_scale_kernel<<<vg, 256, 0, ctx->stream>>>(start, 1.0 / out.start_norm, kr.V, n, ld);
  count_launch(ctx);
  double beta_prev = 0.0;
  for (int k = 0; k < kmax; ++k) {
    const double* vk = kr.V + ld * k;
    SDPSR_TRY(matvec(ctx, lut, vk, kr.w));
    double* h1 = kr.small;
    double* h2 = kr.small + (KR_MAX + 1);
    double* nb = kr.small + 2 * (KR_MAX + 1);
    kr_dots_kernel<<<k + 1, 256, 0, ctx->stream>>>(kr.V, ld, kr.w, n, h1);
    kr_axpy_kernel<<<vg, 256, 0, ctx->stream>>>(kr.V, ld, k + 1, h1, kr.w, n);
    kr_dots_kernel<<<k + 1, 256, 0, ctx->stream>>>(kr.V, ld, kr.w, n, h2);
    kr_axpy_kernel<<<vg, 256, 0, ctx->stream>>>(kr.V, ld, k + 1, h2, kr.w, n);
    kr_dots_kernel<<<1, 256, 0, ctx->stream>>>(kr.w, ld, kr.w, n, nb);
    count_launch(ctx, 5);
    SDPSR_CUDA(cudaGetLastError());
    SDPSR_CUDA(cudaMemcpyAsync(hs, kr.small, (size_t)(2 * (KR_MAX + 1) + 1) * sizeof(double), cudaMemcpyDeviceToHost,
                               ctx->stream));
    SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
    const double a = hs[k] + hs[(KR_MAX + 1) + k];
    const double b = std::sqrt(std::max(0.0, hs[2 * (KR_MAX + 1)]));
    if (!std::isfinite(a) || !std::isfinite(b)) return ctx->fail(SDPSR_E_KRYLOV, "Lanczos produced a non-finite value");
    out.alpha.push_back(a);
    out.steps = k + 1;
    out.scale = std::max(out.scale, std::sqrt(a * a + b * b + beta_prev * beta_prev));
    if (b <= tol * (anorm > 0.0 ? anorm : out.scale)) {
      out.breakdown = true;
      return SDPSR_OK;
    }
    if (k + 1 < kmax) {
      out.beta.push_back(b);
      kr_scale_kernel<<<vg, 256, 0, ctx->stream>>>(kr.w, 1.0 / b, kr.V + ld * (k + 1), n, ld);
      count_launch(ctx);
    }
    beta_prev = b;
  }
  return SDPSR_OK;
}

// index of the eigenvalue of A1 closest to t, or -1 when none is within match_tol
int match_value(const Krylov& kr, double t, double match_tol) {
  int best = -1;
  double bd = 0.0;
  for (int i = 0; i < kr.ne; ++i) {
    const double d = std::fabs(kr.th[(size_t)i] - t);
    if (best < 0 || d < bd) {
      best = i;
      bd = d;
    }
  }
  return (best >= 0 && bd <= match_tol) ? best : -1;
}

int ensure_state(sdpsr_ctx* ctx, Krylov** out) {
  if (!ctx->krylov) ctx->krylov = new Krylov();
  Krylov* kr = reinterpret_cast<Krylov*>(ctx->krylov);
  const size_t ld = (size_t)ctx->ld;
  SDPSR_TRY(sdpsr_scratch_t(ctx, 21, ld * (KR_MAX + 1), &kr->V));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 24, ld, &kr->w));
  SDPSR_TRY(sdpsr_scratch_t(ctx, 25, (size_t)(KR_MAX + 1) * (KR_MAX + 4), &kr->small));
  *out = kr;
  return SDPSR_OK;
}

int build_lut_from(sdpsr_ctx* ctx, const double* r, int64_t len) {
  SDPSR_REQUIRE(len == ctx->dim, SDPSR_E_INVALID, "length(values) != dim(P) (src/partitions.jl:69)");
  SDPSR_REQUIRE(r != nullptr || len == 0, SDPSR_E_INVALID, "coefficient vector is NULL");
  SDPSR_TRY(sdpsr_upload_values(ctx, r, len));
  SDPSR_TRY(sdpsr_build_lut(ctx, ctx->d_values, len));
  ctx->x_is_fill = false;   // the lut no longer belongs to X
  return SDPSR_OK;
}

// Ritz vectors: dst[:, t] = sgn[t] * V[:, 0..k) * Z[:, sel[t]]
int ritz_vectors(sdpsr_ctx* ctx, Krylov& kr, int k, const std::vector<double>& Z, const std::vector<int>& sel,
                 const std::vector<double>& sgn, double* dst) {
  const int nsel = (int)sel.size();
  if (nsel == 0) return SDPSR_OK;
  double* dS = kr.small + 3 * (KR_MAX + 1);
  double* dsgn = dS + (size_t)KR_MAX * KR_MAX;
  int* dsel = reinterpret_cast<int*>(dsgn + KR_MAX);
  SDPSR_CUDA(cudaMemcpyAsync(dS, Z.data(), (size_t)k * k * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  SDPSR_CUDA(cudaMemcpyAsync(dsgn, sgn.data(), (size_t)nsel * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  SDPSR_CUDA(cudaMemcpyAsync(dsel, sel.data(), (size_t)nsel * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  kr_ritz_kernel<<<vec_grid(ctx), 256, (size_t)k * nsel * sizeof(double), ctx->stream>>>(
      kr.V, ctx->ld, k, dS, dsel, dsgn, nsel, dst, ctx->ld, ctx->n);
  count_launch(ctx);
  SDPSR_CUDA(cudaGetLastError());
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));   // Z / sel / sgn are caller stack data
  return SDPSR_OK;
}

}  // namespace

void sdpsr_krylov_free(sdpsr_ctx* ctx) {
  if (ctx->krylov) {
    delete reinterpret_cast<Krylov*>(ctx->krylov);
    ctx->krylov = nullptr;
  }
}

#define CTX_ENTER()                 \
  if (!ctx) return SDPSR_E_INVALID; \
  if (cudaSetDevice(ctx->device) != cudaSuccess) return ctx->fail(SDPSR_E_CUDA, "cudaSetDevice failed")

static int finish(sdpsr_ctx* ctx) {
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  SDPSR_CUDA(cudaGetLastError());
  return SDPSR_OK;
}

extern "C" int sdpsr_eig_krylov(sdpsr_ctx* ctx, const double* r1, int64_t len, int64_t max_steps, double tol,
                                double* vals, int64_t* mult, int64_t* ne_out) {
  CTX_ENTER();
  SDPSR_REQUIRE(vals && mult && ne_out, SDPSR_E_INVALID, "vals / mult / ne is NULL");
  SDPSR_REQUIRE(max_steps >= 1, SDPSR_E_INVALID, "max_steps must be positive");
  SDPSR_REQUIRE(tol > 0.0 && tol < 1e-4, SDPSR_E_INVALID, "tol must be in (0, 1e-4)");
  SDPSR_TRY(sdpsr_shard_ensure_full_labels(ctx));
  int sym = 0;
  SDPSR_TRY(sdpsr_symmetric_check(ctx, &sym));
  SDPSR_REQUIRE(sym, SDPSR_E_NOT_SYMMETRIC,
                "partition is not transpose-invariant: no real symmetric eigendecomposition "
                "(InvalidDecompositionField, src/eigen_decomposition.jl:247-253)");
  Krylov* kr = nullptr;
  SDPSR_TRY(ensure_state(ctx, &kr));
  kr->ready = false;
  kr->tol = tol;
  const int64_t n = ctx->n, ld = ctx->ld;
  // A1 = fill(S, r1) as a LUT (kept for the later calls; ctx->lut is overwritten by r2 / r3)
  SDPSR_TRY(build_lut_from(ctx, r1, len));
  const size_t lut_len = (size_t)ctx->tab[ctx->cur].cap + 1;
  SDPSR_TRY(sdpsr_scratch_t(ctx, 20, lut_len, &kr->lut1));
  SDPSR_CUDA(cudaMemcpyAsync(kr->lut1, ctx->lut, lut_len * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  // generic start vector: fixed pseudo-random sequence (identical on every rank).  Start vectors live
  // in the spare last column of the basis buffer (Lanczos uses columns 0 .. KR_MAX-1 only).
  double* startv = kr->V + ld * KR_MAX;
  {
    std::vector<double> v0((size_t)ld, 0.0);
    uint64_t s = 0x9e3779b97f4a7c15ull;
    for (int64_t i = 0; i < n; ++i) {
      s += 0x9e3779b97f4a7c15ull;
      uint64_t z = s;
      z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
      z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
      z ^= z >> 31;
      v0[(size_t)i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    }
    SDPSR_CUDA(cudaMemcpyAsync(startv, v0.data(), (size_t)ld * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  const int kmax = (int)std::min<int64_t>(std::min<int64_t>(max_steps, n), KR_MAX);
  LanczosOut lo;
  SDPSR_TRY(lanczos(ctx, *kr, kr->lut1, startv, kmax, tol, 0.0, lo));
  if (!lo.breakdown)
    return ctx->fail(SDPSR_E_KRYLOV, "no clean Lanczos breakdown within " + std::to_string(kmax) +
                                         " steps: use the dense path (sdpsr_eig)");
  const int ne = lo.steps;
  std::vector<double> Z;
  tridiag_eig(lo.alpha, lo.beta, ne, kr->th, Z);
  kr->ne = ne;
  kr->anorm = 0.0;
  for (double t : kr->th) kr->anorm = std::max(kr->anorm, std::fabs(t));
  if (!(kr->anorm > 0.0)) kr->anorm = 1.0;
  for (int i = 0; i + 1 < ne; ++i)
    if (!(kr->th[(size_t)i + 1] - kr->th[(size_t)i] > 1e-6 * kr->anorm))
      return ctx->fail(SDPSR_E_KRYLOV, "Ritz values (nearly) coincide: a degenerate copy entered the Krylov space");
  // Y = V * Z : one unit eigenvector per eigenspace, ascending eigenvalue
  SDPSR_TRY(sdpsr_scratch_t(ctx, 22, (size_t)ld * (size_t)ne, &kr->Y));
  SDPSR_CUDA(cudaMemsetAsync(kr->Y, 0, (size_t)ld * (size_t)ne * sizeof(double), ctx->stream));
  {
    std::vector<int> sel((size_t)ne);
    std::vector<double> sgn((size_t)ne, 1.0);
    for (int i = 0; i < ne; ++i) sel[(size_t)i] = i;
    SDPSR_TRY(ritz_vectors(ctx, *kr, ne, Z, sel, sgn, kr->Y));
  }
  // multiplicities: trace(P_i) = sum over diagonal classes c of |c| * (P_i)_rr, r in c
  std::vector<uint32_t> dlab((size_t)n);
  {
    uint32_t* d_diag = nullptr;
    SDPSR_TRY(sdpsr_scratch_t(ctx, 0, (size_t)n, &d_diag));
    kr_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->labels, ctx->tab[ctx->cur].rank, n, ld,
                                                                         d_diag);
    count_launch(ctx);
    SDPSR_CUDA(cudaMemcpyAsync(dlab.data(), d_diag, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  std::vector<std::pair<uint32_t, int64_t>> reps;   // (label, first row) in order of first occurrence
  std::vector<int64_t> sizes;
  for (int64_t i = 0; i < n; ++i) {
    size_t c = 0;
    while (c < reps.size() && reps[c].first != dlab[(size_t)i]) ++c;
    if (c == reps.size()) {
      if ((int)reps.size() >= KR_MAX_DIAG)
        return ctx->fail(SDPSR_E_KRYLOV, "more than " + std::to_string(KR_MAX_DIAG) + " diagonal classes");
      reps.emplace_back(dlab[(size_t)i], i);
      sizes.push_back(0);
    }
    sizes[c] += 1;
  }
  std::vector<double> m((size_t)ne, 0.0);
  const double match_tol = 1e-7 * kr->anorm;
  for (size_t c = 0; c < reps.size(); ++c) {
    kr_unit_kernel<<<vec_grid(ctx), 256, 0, ctx->stream>>>(startv, ld, reps[c].second);
    count_launch(ctx);
    LanczosOut l2;
    // The Krylov space of e_r has at most ne dimensions, so ne steps give the exact quadrature whether or
    // not the last beta drops below the threshold; the run validates itself below (every weighted Ritz
    // value must match an eigenvalue, the dimensions must be integers and add up to N).
    SDPSR_TRY(lanczos(ctx, *kr, kr->lut1, startv, ne, kr->soft_tol, kr->anorm, l2));
    std::vector<double> t2, Z2;
    tridiag_eig(l2.alpha, l2.beta, l2.steps, t2, Z2);
    for (int k = 0; k < l2.steps; ++k) {
      const double w = Z2[(size_t)0 + (size_t)l2.steps * k];
      const int i = match_value(*kr, t2[(size_t)k], match_tol);
      if (i < 0) {
        if (w * w > 1e-12) return ctx->fail(SDPSR_E_KRYLOV, "multiplicity run: unmatched Ritz value");
        continue;
      }
      m[(size_t)i] += w * w * (double)sizes[c];
    }
  }
  int64_t total = 0;
  kr->mult.assign((size_t)ne, 0);
  for (int i = 0; i < ne; ++i) {
    const double r = std::nearbyint(m[(size_t)i]);
    if (!(std::fabs(m[(size_t)i] - r) <= 1e-6 * std::max(1.0, r)) || r < 1.0)
      return ctx->fail(SDPSR_E_KRYLOV, "non-integer eigenspace dimension " + std::to_string(m[(size_t)i]));
    kr->mult[(size_t)i] = (int64_t)r;
    total += (int64_t)r;
  }
  if (total != n) return ctx->fail(SDPSR_E_KRYLOV, "eigenspace dimensions do not add up to N");
  for (int i = 0; i < ne; ++i) {
    vals[i] = kr->th[(size_t)i];
    mult[i] = kr->mult[(size_t)i];
  }
  *ne_out = ne;
  kr->ready = true;
  ctx->have_Q = false;
  return finish(ctx);
}

extern "C" int sdpsr_block_norms_krylov(sdpsr_ctx* ctx, const double* r2, int64_t len, double* norms) {
  CTX_ENTER();
  Krylov* kr = reinterpret_cast<Krylov*>(ctx->krylov);
  SDPSR_REQUIRE(kr && kr->ready, SDPSR_E_STATE, "sdpsr_block_norms_krylov must follow sdpsr_eig_krylov");
  SDPSR_REQUIRE(norms != nullptr, SDPSR_E_INVALID, "norms is NULL");
  const int ne = kr->ne;
  const int64_t ld = ctx->ld;
  SDPSR_TRY(build_lut_from(ctx, r2, len));                      // A2 = fill(S, r2)          (:259)
  SDPSR_TRY(sdpsr_scratch_t(ctx, 23, (size_t)ld * (size_t)ne, &kr->U));
  SDPSR_CUDA(cudaMemsetAsync(kr->U, 0, (size_t)ld * (size_t)ne * sizeof(double), ctx->stream));
  for (int i = 0; i < ne; ++i) SDPSR_TRY(matvec(ctx, ctx->lut, kr->Y + ld * i, kr->U + ld * i));
  std::vector<double> T((size_t)ne * ne, 0.0);
  const double match_tol = 1e-7 * kr->anorm;
  for (int i = 0; i < ne; ++i) {
    LanczosOut lo;
    SDPSR_TRY(lanczos(ctx, *kr, kr->lut1, kr->U + ld * i, ne, kr->soft_tol, kr->anorm, lo));
    if (lo.steps == 0) continue;                                 // A2 y_i == 0
    if (!lo.breakdown) return ctx->fail(SDPSR_E_KRYLOV, "isomorphism run: no clean breakdown");
    std::vector<double> t, Z;
    tridiag_eig(lo.alpha, lo.beta, lo.steps, t, Z);
    for (int k = 0; k < lo.steps; ++k) {
      const double w = std::fabs(Z[(size_t)0 + (size_t)lo.steps * k]);
      const int j = match_value(*kr, t[(size_t)k], match_tol);
      if (j < 0) {
        if (w > 1e-6) return ctx->fail(SDPSR_E_KRYLOV, "isomorphism run: unmatched Ritz value");
        continue;
      }
      T[(size_t)i + (size_t)ne * j] = std::max(T[(size_t)i + (size_t)ne * j], w * lo.start_norm);
    }
  }
  // ||P_j A2 y_i|| is symmetric in exact arithmetic; equal-dimension mask of block_norms (:183-190)
  for (int j = 0; j < ne; ++j)
    for (int i = 0; i < ne; ++i) {
      const double v = std::max(T[(size_t)i + (size_t)ne * j], T[(size_t)j + (size_t)ne * i]);
      norms[(size_t)i + (size_t)ne * j] = kr->mult[(size_t)i] == kr->mult[(size_t)j] ? v : 0.0;
    }
  return finish(ctx);
}

extern "C" int sdpsr_irreducible_krylov(sdpsr_ctx* ctx, const double* r3, int64_t len, const int64_t* kroot,
                                        double atol, int64_t* blk_sizes, int64_t* nblk) {
  CTX_ENTER();
  Krylov* kr = reinterpret_cast<Krylov*>(ctx->krylov);
  SDPSR_REQUIRE(kr && kr->ready, SDPSR_E_STATE, "sdpsr_irreducible_krylov must follow sdpsr_eig_krylov");
  SDPSR_REQUIRE(kroot && blk_sizes && nblk, SDPSR_E_INVALID, "bad arguments");
  const int ne = kr->ne;
  const int64_t ld = ctx->ld;
  std::vector<std::vector<int>> classes;                         // in order of the root eigenspace (:299-309)
  std::vector<int> class_of((size_t)ne, -1);
  for (int e = 0; e < ne; ++e) {
    const int64_t r = kroot[e];
    SDPSR_REQUIRE(r >= 0 && r <= e && kroot[r] == r, SDPSR_E_INVALID,
                  "kroot[e] must be the smallest member of e's class (src/eigen_decomposition.jl:310)");
    if (r == e) {
      class_of[(size_t)e] = (int)classes.size();
      classes.emplace_back();
    }
    classes[(size_t)class_of[(size_t)r]].push_back(e);
    class_of[(size_t)e] = class_of[(size_t)r];
  }
  int64_t S = 0;
  for (auto& k : classes) S += (int64_t)k.size();
  SDPSR_TRY(sdpsr_scratch_t(ctx, 16, (size_t)ld * (size_t)S, &ctx->Qhat));
  SDPSR_CUDA(cudaMemsetAsync(ctx->Qhat, 0, (size_t)ld * (size_t)S * sizeof(double), ctx->stream));
  ctx->qhat_cols = S;
  ctx->blk_sizes.clear();
  // draw #3 is consumed whether or not it is needed, like the reference (:306)
  SDPSR_TRY(build_lut_from(ctx, r3, len));
  const double match_tol = 1e-7 * kr->anorm;
  double* startv = kr->V + ld * KR_MAX;
  int64_t col = 0;
  for (auto& k : classes) {
    const int s = (int)k.size();
    ctx->blk_sizes.push_back(s);
    const int root = k[0];
    // first column = the eigenvector of the root eigenspace                   (:311-314)
    SDPSR_CUDA(cudaMemcpyAsync(ctx->Qhat + ld * col, kr->Y + ld * root, (size_t)ld * sizeof(double),
                               cudaMemcpyDeviceToDevice, ctx->stream));
    if (s > 1) {
      for (int t = 1; t < s; ++t)
        SDPSR_REQUIRE(kr->mult[(size_t)k[(size_t)t]] == kr->mult[(size_t)root], SDPSR_E_INVALID,
                      "isomorphic eigenspaces must have equal dimension");
      // u = A3 y_root spans, under A1, exactly the s eigen-directions P_j u, j in the class  (:327-336)
      SDPSR_TRY(matvec(ctx, ctx->lut, kr->Y + ld * root, startv));
      LanczosOut lo;
      SDPSR_TRY(lanczos(ctx, *kr, kr->lut1, startv, s, kr->soft_tol, kr->anorm, lo));
      if (!lo.breakdown || lo.steps != s)
        return ctx->fail(SDPSR_E_KRYLOV, "class run: Krylov dimension " + std::to_string(lo.steps) +
                                             (lo.breakdown ? "" : "+") + " != class size " + std::to_string(s));
      std::vector<double> t3, Z3;
      tridiag_eig(lo.alpha, lo.beta, s, t3, Z3);
      std::vector<int> sel;
      std::vector<double> sgn;
      for (int t = 1; t < s; ++t) {
        const int j = k[(size_t)t];
        int best = -1;
        for (int q = 0; q < s; ++q)
          if (std::fabs(t3[(size_t)q] - kr->th[(size_t)j]) <= match_tol &&
              (best < 0 || std::fabs(t3[(size_t)q] - kr->th[(size_t)j]) < std::fabs(t3[(size_t)best] - kr->th[(size_t)j])))
            best = q;
        if (best < 0) return ctx->fail(SDPSR_E_KRYLOV, "class run: eigenvalue of a member not found");
        sel.push_back(best);
        // Ritz vector . u = ||u|| * Z3[0, best]: orient it along the projection of u
        sgn.push_back(Z3[(size_t)0 + (size_t)s * best] < 0.0 ? -1.0 : 1.0);
      }
      SDPSR_TRY(ritz_vectors(ctx, *kr, s, Z3, sel, sgn, ctx->Qhat + ld * (col + 1)));
    }
    col += s;
  }
  *nblk = (int64_t)classes.size();
  for (size_t k = 0; k < classes.size(); ++k) blk_sizes[k] = ctx->blk_sizes[k];
  {                                                             // clamptol!.(Q_hat, atol)   (src/diagonalize.jl:39)
    const uint64_t total = (uint64_t)ld * (uint64_t)S;
    const int grid = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)ctx->sm_count * 16);
    kr_clamp_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->Qhat, total, atol);
    count_launch(ctx);
  }
  return finish(ctx);
}
