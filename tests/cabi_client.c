/* cabi_client.c -- a plain C client of libsdpsr_cuda.so (include/sdpsr.h only; no Python, no torch).
 *
 * Replays, call for call, what integration/julia/SDPSRCuda.jl does for
 *     P = admissible_subspace(CuPartition, C, A, b)        (src/partitions.jl:109-190)
 *     blockDiagonalize(P)                                   (src/compat.jl:26-68)
 * with host-supplied initial elements CL, X0 (the Julia wrapper computes them with the reference's own
 * qr / Krylov.craig) and host-supplied coefficient vectors (the wrapper's `rand`), so that a test can hold the
 * result against the oracle: tests/test_gpu_cabi_client.py writes the input file, runs this program and
 * compares the output file.  The scalar steps in between (Otsu threshold, union-find, consistency check:
 * src/eigen_decomposition.jl:83-139,163-219) are restated here in C, as the Julia wrapper reuses the
 * reference's functions for them.
 *
 * input  (little endian): int64 n, m, nnz | int64 rowptr[m+1] | int64 col[nnz] | double val[nnz] |
 *                         double CL[n*n] | double X0[n*n] | double atol | double epsilon |
 *                         int64 ncoef | ncoef x ( int64 len | double v[len] )
 * output: int64 dim | int64 iterations | uint32 labels[n*n] | int64 mode (0 module path, 1 dense) | int64 nblk |
 *         int64 sizes[nblk] | double blocks[dim * sum(s^2)] | int64 ncons_total | int64 cptr[dim+1] | uint32 cidx[...]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/sdpsr.h"

static FILE* fin;
static sdpsr_ctx* ctx;

#define CHECK(call)                                                                                   \
  do {                                                                                                \
    int st_ = (call);                                                                                 \
    if (st_ != SDPSR_OK) {                                                                            \
      fprintf(stderr, "%s failed (%d): %s\n", #call, st_, sdpsr_last_error(ctx));                     \
      exit(2);                                                                                        \
    }                                                                                                 \
  } while (0)

static void rd(void* p, size_t bytes) {
  if (bytes && fread(p, 1, bytes, fin) != bytes) {
    fprintf(stderr, "short read\n");
    exit(3);
  }
}

/* the wrapper's rand(k): the next recorded vector, which must have the length the reference would draw */
static int64_t ncoef_left;
static double* next_coeffs(int64_t k) {
  int64_t len;
  if (ncoef_left-- <= 0) {
    fprintf(stderr, "coefficient vectors exhausted\n");
    exit(4);
  }
  rd(&len, 8);
  if (len != k) {
    fprintf(stderr, "draw of length %lld where the recording has %lld\n", (long long)k, (long long)len);
    exit(4);
  }
  double* v = (double*)malloc((size_t)(len > 0 ? len : 1) * 8);
  rd(v, (size_t)len * 8);
  return v;
}

/* ---- otsu_threshold (src/eigen_decomposition.jl:83-139): 16-bin log histogram + Otsu split ---- */
static double otsu_threshold(const double* X, int64_t count, double atol) {
  int nb = (int)ceil(-log10(2.220446049250313e-16));
  if (nb < 4) nb = 4;
  double lo = INFINITY, hi = 0.0;
  for (int64_t i = 0; i < count; ++i) {
    const double a = fabs(X[i]);
    if (a < lo) lo = a;
    if (a > hi) hi = a;
  }
  if (lo < atol) lo = atol;
  double* edges = (double*)malloc((size_t)(nb + 1) * 8);
  const double l0 = log(lo), l1 = log(hi);
  for (int k = 0; k <= nb; ++k) edges[k] = exp(l0 + (l1 - l0) * (double)k / (double)nb);   /* range(..., length = nb + 1) */
  double* cnt = (double*)calloc((size_t)nb, 8);
  for (int64_t i = 0; i < count; ++i) {
    const double a = fabs(X[i]);
    int k = nb; /* findfirst(e -> e > x) - 1, clamped to 1..nb (1-based) */
    for (int e = 0; e <= nb; ++e)
      if (edges[e] > a) {
        k = e;
        break;
      }
    if (k < 1) k = 1;
    if (k > nb) k = nb;
    cnt[k - 1] += 1.0;
  }
  double w = 0.0, mu = 0.0, muT = 0.0;
  for (int k = 0; k < nb; ++k) muT += log(edges[k]) * cnt[k] / (double)count;
  double best = -1.0;
  int bestk = 0, have_nan = 0;
  for (int k = 0; k + 1 < nb; ++k) {
    const double p = cnt[k] / (double)count;
    w += p;
    mu += log(edges[k]) * p;
    const double s2 = (muT * w - mu) * (muT * w - mu) / (w * (1.0 - w));
    if (s2 != s2) { /* Julia's argmax: NaN is maximal, the first one wins */
      if (!have_nan) {
        bestk = k;
        have_nan = 1;
      }
    } else if (!have_nan && s2 > best) {
      best = s2;
      bestk = k;
    }
  }
  const double thr = edges[bestk + 1];
  free(edges);
  free(cnt);
  return thr;
}

/* ---- IntDisjointSets (DataStructures.jl): union by rank, ties -> the first argument's root ---- */
static int64_t find_root(int64_t* par, int64_t x) {
  int64_t r = x;
  while (par[r] != r) r = par[r];
  while (par[x] != r) {
    const int64_t nx = par[x];
    par[x] = r;
    x = nx;
  }
  return r;
}
static void set_union(int64_t* par, int64_t* rnk, int64_t x, int64_t y) {
  int64_t xr = find_root(par, x), yr = find_root(par, y);
  if (xr == yr) return;
  if (rnk[xr] < rnk[yr]) {
    const int64_t t = xr;
    xr = yr;
    yr = t;
  } else if (rnk[xr] == rnk[yr]) {
    rnk[xr] += 1;
  }
  par[yr] = xr;
}

/* isomorphism classes from the block norms (:205-219) + __isconsistent (:163-167); kroot 0-based */
static void isomorphism_classes(const double* norms, int64_t ne, double atol, int64_t* kroot) {
  const double thr = otsu_threshold(norms, ne * ne, atol);
  int64_t* par = (int64_t*)malloc((size_t)ne * 8);
  int64_t* rnk = (int64_t*)calloc((size_t)ne, 8);
  for (int64_t i = 0; i < ne; ++i) par[i] = i;
  for (int64_t i = 0; i < ne; ++i)
    for (int64_t j = i + 1; j < ne; ++j)
      if (norms[i + ne * j] >= thr) set_union(par, rnk, i, j);
  for (int64_t i = 0; i < ne; ++i) kroot[i] = find_root(par, i);
  for (int64_t i = 0; i < ne; ++i) { /* every root must be the smallest member of its class */
    int64_t first = -1;
    for (int64_t j = 0; j < ne && first < 0; ++j)
      if (kroot[j] == kroot[i]) first = j;
    if (first != kroot[i]) {
      fprintf(stderr, "NumericalInconsistency: the K-partition is inconsistent with the eigenspaces\n");
      exit(5);
    }
  }
  free(par);
  free(rnk);
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s input.bin output.bin\n", argv[0]);
    return 1;
  }
  fin = fopen(argv[1], "rb");
  if (!fin) return 1;
  int64_t n, m, nnz;
  rd(&n, 8);
  rd(&m, 8);
  rd(&nnz, 8);
  int64_t* rowptr = (int64_t*)malloc((size_t)(m + 1) * 8);
  int64_t* col = (int64_t*)malloc((size_t)(nnz ? nnz : 1) * 8);
  double* val = (double*)malloc((size_t)(nnz ? nnz : 1) * 8);
  double* CL = (double*)malloc((size_t)(n * n) * 8);
  double* X0 = (double*)malloc((size_t)(n * n) * 8);
  double atol, epsilon;
  rd(rowptr, (size_t)(m + 1) * 8);
  rd(col, (size_t)nnz * 8);
  rd(val, (size_t)nnz * 8);
  rd(CL, (size_t)(n * n) * 8);
  rd(X0, (size_t)(n * n) * 8);
  rd(&atol, 8);
  rd(&epsilon, 8);
  rd(&ncoef_left, 8);

  if (sdpsr_create(&ctx, n, 0, SDPSR_F_DEFAULT) != SDPSR_OK) {
    fprintf(stderr, "sdpsr_create: %s\n", sdpsr_last_error(NULL));
    return 2;
  }
  /* ---- admissible_subspace(CuPartition, C, A, b) ------------------------------------------------ */
  int64_t d = 0, cur, iterations = 0;
  CHECK(sdpsr_set_constraints_csr(ctx, m, rowptr, col, val, 0));
  CHECK(sdpsr_partition_reset(ctx));
  CHECK(sdpsr_refine_values(ctx, CL, atol, 0, &d));   /* S = Part(CL)            :145 */
  CHECK(sdpsr_refine_values(ctx, X0, atol, 0, &d));   /* refine!(S, Part(X0))    :146 */
  cur = d;
  const int64_t maxdim = (n * n + n) / 2;
  while (cur < maxdim) { /* :154 */
    ++iterations;
    double* r = next_coeffs(cur); /* randomize!(X, S)        :159 */
    CHECK(sdpsr_fill(ctx, r, cur));
    free(r);
    CHECK(sdpsr_project_round_refine(ctx, atol, &d)); /* :160-164 */
    if (d != cur) {                                   /* :166-168 */
      r = next_coeffs(d);
      CHECK(sdpsr_fill(ctx, r, d));
      free(r);
    }
    CHECK(sdpsr_square_round_refine(ctx, atol, &d)); /* :172-174 */
    if (cur == d) break;                              /* :180-182 */
    cur = d;
  }
  CHECK(sdpsr_partition_dim(ctx, &d));
  uint32_t* labels = (uint32_t*)malloc((size_t)(n * n) * 4);
  CHECK(sdpsr_partition_get_labels(ctx, labels, 4));
  /* the same export without the wait (copy stream; pageable memory here, so it degrades to a plain copy) and
   * as the reference's default UInt16 */
  uint16_t* labels16 = (uint16_t*)malloc((size_t)(n * n) * 2);
  CHECK(sdpsr_partition_get_labels_async(ctx, labels16, 2));
  CHECK(sdpsr_partition_labels_wait(ctx));
  for (int64_t i = 0; i < n * n; ++i)
    if ((uint32_t)labels16[i] != labels[i]) {
      fprintf(stderr, "asynchronous label export differs at %lld\n", (long long)i);
      return 1;
    }
  free(labels16);

  /* ---- blockDiagonalize(P): module path first, dense path with the SAME draws when it does not apply -- */
  double* r1 = next_coeffs(d);
  double* r2 = next_coeffs(d);
  double* r3 = next_coeffs(d);
  int64_t mode = 0, nblk = 0;
  int64_t* sizes = (int64_t*)malloc((size_t)n * 8);
  {
    const int64_t maxmod = 2 * d + 16 < n ? 2 * d + 16 : n;
    double* vals = (double*)malloc((size_t)maxmod * 8);
    int64_t* mult = (int64_t*)malloc((size_t)maxmod * 8);
    int64_t ne = 0;
    int st = sdpsr_eig_krylov(ctx, r1, d, maxmod, epsilon, vals, mult, &ne);
    if (st == SDPSR_OK) {
      double* norms = (double*)malloc((size_t)(ne * ne) * 8);
      int64_t* kroot = (int64_t*)malloc((size_t)ne * 8);
      st = sdpsr_block_norms_krylov(ctx, r2, d, norms);
      if (st == SDPSR_OK) {
        isomorphism_classes(norms, ne, epsilon, kroot);
        st = sdpsr_irreducible_krylov(ctx, r3, d, kroot, epsilon, sizes, &nblk);
      }
      free(norms);
      free(kroot);
    }
    if (st == SDPSR_E_KRYLOV) { /* the reference's algorithm step by step */
      mode = 1;
      double* all = (double*)malloc((size_t)n * 8);
      CHECK(sdpsr_eig(ctx, r1, d, all));
      int64_t* ptrs = (int64_t*)malloc((size_t)(n + 1) * 8);
      int64_t np = 0;
      ptrs[np++] = 0;
      for (int64_t i = 1; i < n; ++i)
        if (fabs(all[i] - all[i - 1]) > epsilon) ptrs[np++] = i; /* EigenDecomposition, :19-40 */
      ptrs[np++] = n;
      const int64_t ne2 = np - 1;
      double* norms = (double*)malloc((size_t)(ne2 * ne2) * 8);
      int64_t* kroot = (int64_t*)malloc((size_t)ne2 * 8);
      CHECK(sdpsr_block_norms(ctx, r2, d, ptrs, np, norms));
      isomorphism_classes(norms, ne2, epsilon, kroot);
      CHECK(sdpsr_irreducible(ctx, r3, d, ptrs, np, kroot, epsilon, sizes, &nblk));
      free(all);
      free(ptrs);
      free(norms);
      free(kroot);
    } else if (st != SDPSR_OK) {
      fprintf(stderr, "module path failed (%d): %s\n", st, sdpsr_last_error(ctx));
      return 2;
    }
    free(vals);
    free(mult);
  }
  int64_t fin_dim = 0, sq = 0;
  for (int64_t k = 0; k < nblk; ++k) {
    fin_dim += sizes[k] * (sizes[k] + 1) / 2;
    sq += sizes[k] * sizes[k];
  }
  if (fin_dim != d) { /* check_block_sizes, src/diagonalize.jl:1-23 */
    fprintf(stderr, "DimensionMismatch: final_dim=%lld expected %lld\n", (long long)fin_dim, (long long)d);
    return 6;
  }
  double* blocks = (double*)malloc((size_t)(d * sq > 0 ? d * sq : 1) * 8);
  CHECK(sdpsr_basis_image(ctx, 1e-12 * (double)n, blocks, d * sq));

  /* ---- _constraints(P) (the generic AbstractPartition slow path) ---------------------------------- */
  int64_t zeros = 0;
  CHECK(sdpsr_partition_zero_count(ctx, &zeros));
  const int64_t ncons = n * n - zeros;
  int64_t* cptr = (int64_t*)malloc((size_t)(d + 1) * 8);
  uint32_t* cidx = (uint32_t*)malloc((size_t)(ncons > 0 ? ncons : 1) * 4);
  CHECK(sdpsr_partition_constraints(ctx, cptr, cidx, ncons, 1));

  FILE* fo = fopen(argv[2], "wb");
  if (!fo) return 1;
  fwrite(&d, 8, 1, fo);
  fwrite(&iterations, 8, 1, fo);
  fwrite(labels, 4, (size_t)(n * n), fo);
  fwrite(&mode, 8, 1, fo);
  fwrite(&nblk, 8, 1, fo);
  fwrite(sizes, 8, (size_t)nblk, fo);
  fwrite(blocks, 8, (size_t)(d * sq), fo);
  fwrite(&ncons, 8, 1, fo);
  fwrite(cptr, 8, (size_t)(d + 1), fo);
  fwrite(cidx, 4, (size_t)ncons, fo);
  fclose(fo);
  CHECK(sdpsr_destroy(ctx));
  printf("cabi_client ok: n=%lld dim=%lld iterations=%lld blocks=%lld path=%s\n", (long long)n, (long long)d,
         (long long)iterations, (long long)nblk, mode ? "dense" : "module");
  return 0;
}
