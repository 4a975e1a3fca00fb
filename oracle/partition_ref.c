/*
 * partition_ref.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * A single-threaded C restatement of the reference's partition primitives, used
 * (a) to cross-check the numpy oracle at sizes where np.unique is slow and
 * (b) as the CPU baseline timed by bench.py (`cpu_baseline`, `--impl reference`).
 * The reference runs these loops single-threaded in Julia; only BLAS/LAPACK are
 * threaded there, which bench.py mirrors with numpy/OpenBLAS.
 *
 *   clamp_round        src/utils.jl:34-53   (_clamp_round! / unsafe_round)
 *   part_from_values   src/partitions.jl:24-35  (Dict pass, first-occurrence labels)
 *   sort_unique        src/partitions.jl:44-60  (__sort_unique!: unique + LUT + relabel)
 *   refine             src/partitions.jl:62-66
 *   fill               src/partitions.jl:68-75
 *
 * Julia's Dict is an open-addressing hash table keyed by isequal/hash; the table
 * below plays that role (linear probing, 64-bit mix).  Only the label semantics
 * are pinned (tests/test_oracle_pins.py), not Julia's exact probing order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t mix64(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}

/* _clamp_round!(A; atol): abs(a) < atol -> 0 else ldexp(trunc(scale*x)/scale, n) */
void oracle_clamp_round(double* a, int64_t len, double atol) {
  const int sig = (int)floor(-log10(atol));
  double scale = 1.0;
  for (int i = 0; i < sig; ++i) scale *= 10.0;
  for (int64_t i = 0; i < len; ++i) {
    const double v = a[i];
    if (fabs(v) < atol) { a[i] = 0.0; continue; }
    int n;
    const double x = frexp(v, &n);
    const double q = trunc(scale * x);
    a[i] = ldexp(q / scale, n);
  }
}

typedef struct { uint64_t* keys; uint32_t* vals; uint64_t cap, cnt; } table_t;

static void table_init(table_t* t, uint64_t cap) {
  t->cap = cap; t->cnt = 0;
  t->keys = (uint64_t*)malloc(cap * sizeof(uint64_t));
  t->vals = (uint32_t*)malloc(cap * sizeof(uint32_t));
  memset(t->keys, 0xff, cap * sizeof(uint64_t));
}
static void table_free(table_t* t) { free(t->keys); free(t->vals); }
static void table_grow(table_t* t);
/* get!(d, key, dflt): returns the stored value */
static inline uint32_t table_get_or_set(table_t* t, uint64_t key, uint32_t dflt) {
  if (2 * (t->cnt + 1) > t->cap) table_grow(t);
  uint64_t s = mix64(key) & (t->cap - 1);
  for (;;) {
    if (t->keys[s] == key) return t->vals[s];
    if (t->keys[s] == ~0ULL) { t->keys[s] = key; t->vals[s] = dflt; t->cnt++; return dflt; }
    s = (s + 1) & (t->cap - 1);
  }
}
static void table_grow(table_t* t) {
  table_t n; table_init(&n, t->cap * 4);
  for (uint64_t i = 0; i < t->cap; ++i)
    if (t->keys[i] != ~0ULL) {
      uint64_t s = mix64(t->keys[i]) & (n.cap - 1);
      while (n.keys[s] != ~0ULL) s = (s + 1) & (n.cap - 1);
      n.keys[s] = t->keys[i]; n.vals[s] = t->vals[i]; n.cnt++;
    }
  table_free(t); *t = n;
}

/* Partition{T}(M): Dict(0.0 => 0); labels by first occurrence (column-major = memory order). */
int64_t oracle_part_from_values(const double* M, int64_t len, uint32_t* out) {
  table_t t; table_init(&t, 1024);
  uint32_t l = 0;
  table_get_or_set(&t, 0ULL, 0u);                      /* zero(eltype(M)) => 0; -0.0 has other bits */
  for (int64_t i = 0; i < len; ++i) {
    uint64_t bits; memcpy(&bits, &M[i], 8);
    if (M[i] != M[i]) bits = 0x7ff8000000000000ULL;    /* isequal: NaN == NaN */
    const uint32_t k = table_get_or_set(&t, bits, l + 1);
    if (k == l + 1) l = k;
    out[i] = k;
  }
  table_free(&t);
  return (int64_t)l;
}

/* __sort_unique!(P): unique (hash, first-occurrence order) -> dense LUT -> relabel; 0 kept. */
int64_t oracle_sort_unique(uint64_t* m, int64_t len) {
  table_t seen; table_init(&seen, 1024);
  uint64_t* uniq = (uint64_t*)malloc((size_t)(len > 0 ? len : 1) * sizeof(uint64_t));
  int64_t nu = 0; uint64_t mx = 0;
  for (int64_t i = 0; i < len; ++i) {                   /* unique(P.matrix) */
    const uint64_t before = seen.cnt;
    table_get_or_set(&seen, m[i], 0u);
    if (seen.cnt != before) { uniq[nu++] = m[i]; if (m[i] > mx) mx = m[i]; }
  }
  table_free(&seen);
  int64_t* lut = (int64_t*)calloc((size_t)mx + 1, sizeof(int64_t));   /* zeros(Int, maximum + 1) */
  int64_t dim = 0;
  for (int64_t u = 0; u < nu; ++u) { if (uniq[u] == 0) continue; lut[uniq[u]] = ++dim; }
  for (int64_t i = 0; i < len; ++i) m[i] = (uint64_t)lut[m[i]];
  free(lut); free(uniq);
  return dim;
}

/* refine!(P1, P2): P1 .+= P2 .* (dim(P1)+1); __sort_unique! */
int64_t oracle_refine(uint64_t* p1, int64_t dim1, const uint32_t* p2, int64_t len) {
  for (int64_t i = 0; i < len; ++i) p1[i] += (uint64_t)p2[i] * (uint64_t)(dim1 + 1);
  return oracle_sort_unique(p1, len);
}

/* fill!(M, P; values) */
void oracle_fill(double* M, const uint64_t* p, const double* values, int64_t len) {
  for (int64_t i = 0; i < len; ++i) M[i] = p[i] == 0 ? 0.0 : values[p[i] - 1];
}

/* one loop step of admissible_subspace on existing data: round; Part; refine  (:173-174) */
int64_t oracle_round_refine(uint64_t* labels, int64_t dim1, double* M, int64_t len, double atol, uint32_t* scratch) {
  oracle_clamp_round(M, len, atol);
  oracle_part_from_values(M, len, scratch);
  return oracle_refine(labels, dim1, scratch, len);
}
