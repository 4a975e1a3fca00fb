// refine.cu -- the partition kernels of the Jordan-reduction hot path (sm_100a).
//
// Replaces, as ONE streaming pass over the N x N entries,
//   _clamp_round!            src/utils.jl:34-53
//   Partition{T}(M)          src/partitions.jl:24-35   (Dict pass)
//   refine! + __sort_unique! src/partitions.jl:44-66   (unique + LUT + relabel)
// and fill! (src/partitions.jl:68-75).
//
// Design (DESIGN.md "refine pass"): each entry forms a 64-bit key
//   (old provisional id, code of the rounded value)
// and obtains the slot of that key in a global open-addressing table; slot+1 is
// the entry's new *provisional* id.  A per-CTA shared-memory cache of
// key -> slot keeps global atomics down to O(distinct keys) per CTA.  The table
// records the first (smallest) linear index of every key; canonical labels (the
// reference numbers classes by first occurrence in column-major order) are the
// ranks of those first indices and are computed on the table only -- O(dim) work,
// not O(N^2) -- and applied lazily (fill LUT composition, label export).
// HBM traffic of a pass: 8 B value + 4 B old id + 4 B new id = 16 B / entry.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "sdpsr_internal.cuh"

namespace {

constexpr int RT = 768;           // threads per CTA (one CTA per SM shares one key cache)
constexpr int EPT = 4;            // consecutive entries per thread per tile
constexpr int TILE = RT * EPT;    // entries per tile
constexpr int SC = 8192;          // slots of the per-CTA key cache (128 KB)
constexpr int SC_LIMIT = SC / 2;      // keys cached per CTA; the rest always go to the global table
// current dim above which a pass skips the cache altogether: with 16-byte table slots (one L2 sector per lookup) the
// global path holds 2.3-2.5 TB/s from ~2000 classes up, the cache falls below that at ~2300 (probe chains at load
// factors > 0.28 cost more warp-level iterations than the L2 round trip; sweep in DESIGN.md section 4)
constexpr int CACHE_DIM_LIMIT = 2304;
constexpr int JOINT_MIN_DIM = 1500;    // classes above which refine_fast_kernel probes its four entries jointly
constexpr int CACHE_PROBES = 8;       // linear-probe window of the cache (lookups and inserts)
constexpr uint32_t PROBE_LIMIT = 4096;
constexpr int RANK_BRUTE_MAX = 16384;

struct RefineArgs {
  const uint32_t* lab_in;
  const uint32_t* lab2;
  const double* vals;
  double* vals_out;
  const double* lut;
  const double* tpat;
  const uint32_t* pid;
  uint32_t* lab_out;
  uint64_t total;
  uint32_t idx0;        // padded linear index of the first entry of the range (sharded passes)
  double atol;
  double scale;
  long long iscale;
  int qbits;
  int lbits;
  int do_round;
  int fillproj;
  int use_cache;
  int raw_bits;
  KeySlot* gslots;
  uint32_t* gocc;
  uint32_t* gmeta;
  uint32_t gmask;
  uint32_t glimit;
};

__device__ __forceinline__ uint64_t mix64(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

__device__ __forceinline__ uint64_t ld_vol64(const uint64_t* p) {
  return *reinterpret_cast<const volatile uint64_t*>(p);
}
__device__ __forceinline__ uint32_t ld_vol32(const uint32_t* p) {
  return *reinterpret_cast<const volatile uint32_t*>(p);
}
__device__ __forceinline__ uint4 ld_vol128(const void* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// _clamp_round! of one value (src/utils.jl:34-53) as an integer code:
//   0                       if |a| < atol
//   sign | biased exp | |q| otherwise, with a = x * 2^n, |x| in [0.5,1), q = trunc(scale*x).
// The code is an injective function of the rounded value ldexp(q/scale, n); the one
// non-canonical representation (|q| == scale, i.e. y == 1.0) is folded onto
// (scale/2, n+1), which denotes the same double.  `rounded` receives the value the
// reference would store.
template <bool NEED_VALUE>
__device__ __forceinline__ uint64_t round_code(double a, double atol, double scale, long long iscale,
                                               int qbits, double& rounded) {
  if (fabs(a) < atol) {
    rounded = 0.0;
    return 0ull;
  }
  const uint64_t bits = (uint64_t)__double_as_longlong(a);
  const uint64_t sign = bits >> 63;
  uint32_t e = (uint32_t)(bits >> 52) & 0x7ffu;   // biased exponent, n = e - 1022
  const double x = __longlong_as_double((long long)((bits & 0x800fffffffffffffull) | (0x3feull << 52)));
  const double p = __dmul_rn(scale, x);            // one IEEE multiply, no FMA contraction
  const long long q = __double2ll_rz(p);           // unsafe_trunc(Int, .)
  if (NEED_VALUE) {
    // ldexp(q/scale, n) with n = e - 1022: y in [0.5, 1] and e is a normal exponent, so scaling
    // by 2^n is an exact exponent adjustment (done in two exact steps to stay in range)
    const double y = __ddiv_rn((double)q, scale);
    const int n = (int)e - 1022;
    const int n1 = n / 2, n2 = n - n1;
    rounded = y * __longlong_as_double((long long)(1023 + n1) << 52) * __longlong_as_double((long long)(1023 + n2) << 52);
  }
  unsigned long long aq = (unsigned long long)(q < 0 ? -q : q);
  if (aq == (unsigned long long)iscale) {          // y == 1.0: same double as (scale/2, n+1)
    aq = (unsigned long long)(iscale >> 1);
    e += 1;
  }
  return (sign << (11 + qbits)) | ((uint64_t)e << qbits) | aq;
}

__device__ __forceinline__ uint64_t raw_code(double v) {
  uint64_t b = (uint64_t)__double_as_longlong(v);
  if (v != v) b = 0x7ff8000000000000ull;           // isequal: all NaNs are one key
  return b;                                         // +0.0 -> 0 (the zero key); -0.0 stays distinct
}

__device__ __forceinline__ uint32_t global_insert(const RefineArgs& a, uint64_t k, uint64_t h) {
  uint32_t s = (uint32_t)(h >> 20) & a.gmask;
  for (uint32_t probe = 0; probe < PROBE_LIMIT; ++probe) {
    const uint64_t kk = ld_vol64(&a.gslots[s].key);
    if (kk == k) return s + 1;
    if (kk == KEY_EMPTY) {
      const unsigned long long old =
          atomicCAS(reinterpret_cast<unsigned long long*>(&a.gslots[s].key), KEY_EMPTY, (unsigned long long)k);
      if (old == KEY_EMPTY) {
        const uint32_t pos = atomicAdd(a.gmeta, 1u);
        if (pos < a.glimit)
          a.gocc[pos] = s;
        else
          a.gmeta[1] = 1u;                           // table too small: host grows it and reruns
        return s + 1;
      }
      if (old == (unsigned long long)k) return s + 1;
    }
    s = (s + 1) & a.gmask;
  }
  a.gmeta[1] = 1u;
  return 1u;
}

// One cache slot: {key, provisional id, tile of insertion}; 16 bytes, read with one LDS.128.
struct __align__(16) CacheSlot {
  uint64_t key;
  uint32_t gid;     // 0 while the inserting thread has not published yet
  uint32_t tile;    // CTA tile during which the key entered this cache
};

struct Entry4 {
  uint4 lab;
  uint4 aux;        // KM_PAIR: second ids; fill-projection: pattern ids
  double v[EPT];
};

template <int MODE>
__device__ __forceinline__ void load_entries(const RefineArgs& a, uint64_t base, Entry4& e) {
  e.lab = make_uint4(0u, 0u, 0u, 0u);
  if (a.lab_in) e.lab = __ldcs(reinterpret_cast<const uint4*>(a.lab_in + base));
  if (MODE == KM_PAIR) {
    e.aux = __ldcs(reinterpret_cast<const uint4*>(a.lab2 + base));
  } else if (a.fillproj) {
    e.aux = __ldcs(reinterpret_cast<const uint4*>(a.pid + base));
  } else {
    const double2 v0 = __ldcs(reinterpret_cast<const double2*>(a.vals + base));
    const double2 v1 = __ldcs(reinterpret_cast<const double2*>(a.vals + base + 2));
    e.v[0] = v0.x; e.v[1] = v0.y; e.v[2] = v1.x; e.v[3] = v1.y;
  }
}

// Keep the first occurrence: the table entry only ever decreases, so a stale (larger) read
// merely costs one redundant atomic.
__device__ __forceinline__ void note_first(const RefineArgs& a, uint32_t g, uint32_t idx) {
  if (idx < ld_vol32(&a.gslots[g - 1].minidx)) atomicMin(&a.gslots[g - 1].minidx, idx);
}

__device__ __noinline__ uint32_t refine_miss(KeySlot* gslots, uint32_t* gocc, uint32_t* gmeta,
                                             uint32_t gmask, uint32_t glimit, CacheSlot* cache, uint32_t* s_count,
                                             uint64_t k, uint32_t idx, int free_slot, uint32_t mytile) {
  RefineArgs a;                      // only the table fields are read below
  a.gslots = gslots;
  a.gocc = gocc;
  a.gmeta = gmeta;
  a.gmask = gmask;
  a.glimit = glimit;
  const uint32_t g = global_insert(a, k, mix64(k));
  note_first(a, g, idx);
  if (free_slot >= 0 && *reinterpret_cast<volatile uint32_t*>(s_count) < (uint32_t)SC_LIMIT) {
    const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(&cache[free_slot].key), KEY_EMPTY,
                                             (unsigned long long)k);
    if (old == KEY_EMPTY) {
      *reinterpret_cast<volatile unsigned long long*>(&cache[free_slot].gid) = ((unsigned long long)mytile << 32) | g;
      atomicAdd(s_count, 1u);
    }
  }
  return g;
}

// The pass is barrier-free.  A thread resolves a key in the CTA cache with one 16-byte shared
// load; on a miss it goes to the global table itself and then publishes (key, id, tile) to the
// cache.  First-occurrence bookkeeping: a CTA walks its tiles in increasing index order, so a
// hit on a slot inserted during an EARLIER tile has a larger index than the inserter's and
// needs no update; a hit from the same or an earlier tile (warps of a CTA may be a tile
// apart) updates the global minimum itself.
template <int MODE, bool WRITEBACK>
__global__ void __launch_bounds__(RT) refine_kernel(const RefineArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CacheSlot* cache = reinterpret_cast<CacheSlot*>(smem_raw);
  __shared__ uint32_t s_count;

  for (int i = threadIdx.x; i < SC; i += RT) {
    cache[i].key = KEY_EMPTY;
    cache[i].gid = 0u;
    cache[i].tile = 0xffffffffu;
  }
  if (threadIdx.x == 0) s_count = 0u;
  __syncthreads();

  const uint64_t ntiles = (a.total + TILE - 1) / TILE;
  uint64_t tile = blockIdx.x;
  if (tile >= ntiles) return;
  Entry4 cur, nxt;
  uint64_t base = tile * TILE + (uint64_t)threadIdx.x * EPT;
  if (base < a.total) load_entries<MODE>(a, base, cur);
  uint32_t iter = 0;

  for (; tile < ntiles; tile += gridDim.x, ++iter) {
    base = tile * TILE + (uint64_t)threadIdx.x * EPT;
    const bool active = base < a.total;               // total % EPT == 0: whole vector or nothing
    // prefetch the next tile's operands before touching this one (memory-level parallelism)
    const uint64_t nbase = (tile + gridDim.x) * TILE + (uint64_t)threadIdx.x * EPT;
    if (tile + gridDim.x < ntiles && nbase < a.total) load_entries<MODE>(a, nbase, nxt);
    if ((iter & 15u) == 15u && ld_vol32(a.gmeta + 1)) return;   // table overflowed: host reruns
    const unsigned wmask = __ballot_sync(0xffffffffu, active);

    if (active) {
      const uint32_t lv[EPT] = {cur.lab.x, cur.lab.y, cur.lab.z, cur.lab.w};
      const uint32_t av[EPT] = {cur.aux.x, cur.aux.y, cur.aux.z, cur.aux.w};
      uint64_t key[EPT];
      if (MODE == KM_PAIR) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) key[e] = ((uint64_t)av[e] << a.lbits) | lv[e];
      } else {
        double r[EPT];
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          const double v = a.fillproj ? __dsub_rn(__ldg(a.lut + lv[e]), __ldg(a.tpat + av[e])) : cur.v[e];
          uint64_t code;
          if (a.do_round) {
            code = round_code<WRITEBACK || MODE == KM_RAW>(v, a.atol, a.scale, a.iscale, a.qbits, r[e]);
            if (MODE == KM_RAW) code = raw_code(r[e]);
          } else {
            r[e] = v;
            code = a.raw_bits ? (uint64_t)__double_as_longlong(v) : raw_code(v);
          }
          key[e] = (MODE == KM_RAW) ? code : ((code << a.lbits) | lv[e]);
        }
        if (WRITEBACK) {
          __stcs(reinterpret_cast<double2*>(a.vals_out + base), make_double2(r[0], r[1]));
          __stcs(reinterpret_cast<double2*>(a.vals_out + base + 2), make_double2(r[2], r[3]));
        }
      }

      uint32_t gid[EPT];
      const uint32_t mytile = (uint32_t)iter;
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const uint64_t k = key[e];
        const uint32_t idx = a.idx0 + (uint32_t)(base + e);
        const bool same = e > 0 && k == key[e - 1];      // run of equal keys: first one did the work
        const bool need = k != 0ull && !same;            // the zero class keeps id 0
        uint32_t g = same ? gid[e - 1] : 0u;
        const uint64_t hf = k * 0x9e3779b97f4a7c15ull;   // Fibonacci hash for the cache
        uint32_t s = (uint32_t)(hf >> 40) & (SC - 1);
        int free_slot = -1;
        if (need && a.use_cache) {
#pragma unroll 1
          for (int probe = 0; probe < CACHE_PROBES; ++probe) {
            uint4 raw;   // one LDS.128 (volatile: other warps publish concurrently)
            asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w)
                         : "r"((uint32_t)__cvta_generic_to_shared(&cache[s])));
            const uint64_t kk = ((uint64_t)raw.y << 32) | raw.x;
            if (kk == k) {
              g = raw.z;                                  // 0: publisher in flight -> miss path
              if (g != 0u && mytile <= raw.w) note_first(a, g, idx);
              break;
            }
            if (kk == KEY_EMPTY) {
              free_slot = (int)s;
              break;
            }
            s = (s + 1) & (SC - 1);
          }
        }
        __syncwarp(wmask);      // reconverge (see refine_fast_kernel)
        if (need && g == 0u)
          g = refine_miss(a.gslots, a.gocc, a.gmeta, a.gmask, a.glimit, cache, &s_count, k, idx, free_slot,
                          mytile);
        __syncwarp(wmask);
        gid[e] = g;
      }
      __stcs(reinterpret_cast<uint4*>(a.lab_out + base), make_uint4(gid[0], gid[1], gid[2], gid[3]));
    }
    cur = nxt;
  }
}

// ---------------------------------------------------------------------------
// Fast path of the round+refine pass for the reference's default tolerance
// (atol = sqrt(eps) -> scale = 10^7, |q| < 2^24) and provisional ids < 2^28:
// the key has a FIXED layout   hi = sign(1) | biased exp(11) | q>>4 (20)
//                              lo = q&15 (4) | old id (28)
// so that it is assembled with a handful of 32-bit operations, hashed with two 32-bit
// multiplies and compared word-wise.  Same protocol as refine_kernel otherwise; the cold
// miss path is out of line.
// ---------------------------------------------------------------------------
template <bool WRITEBACK, bool FILLPROJ, bool JOINT>
__global__ void __launch_bounds__(RT, 1) refine_fast_kernel(const RefineArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CacheSlot* cache = reinterpret_cast<CacheSlot*>(smem_raw);
  __shared__ uint32_t s_count;
  for (int i = threadIdx.x; i < SC; i += RT) {
    cache[i].key = KEY_EMPTY;
    cache[i].gid = 0u;
    cache[i].tile = 0xffffffffu;
  }
  if (threadIdx.x == 0) s_count = 0u;
  __syncthreads();

  const uint64_t total = a.total;
  const uint64_t ntiles = (total + TILE - 1) / TILE;
  const double atol = a.atol;
  const uint32_t* __restrict__ lab_in = a.lab_in;
  const double* __restrict__ vals = a.vals;
  const uint32_t* __restrict__ pid = a.pid;
  const double* __restrict__ lut = a.lut;
  const double* __restrict__ tpat = a.tpat;
  uint32_t* __restrict__ lab_out = a.lab_out;
  double* __restrict__ vals_out = a.vals_out;
  const uint32_t cache_base = (uint32_t)__cvta_generic_to_shared(cache);
  const bool use_cache = a.use_cache != 0;

  uint64_t tile = blockIdx.x;
  if (tile >= ntiles) return;
  // Two tiles of operands are kept in flight per thread (registers): with 768 threads per SM that
  // is ~72 KB outstanding, enough to cover HBM latency at full bandwidth (one tile was not: ncu
  // showed 40 % of the stall samples on the first use of the loaded values).
  struct Stage {
    uint4 lab, aux;
    double2 v0, v1;
  };
  auto load_stage = [&](Stage& st, uint64_t b) {
    st.lab = make_uint4(0u, 0u, 0u, 0u);
    if (b < total) {
      if (lab_in) st.lab = __ldcs(reinterpret_cast<const uint4*>(lab_in + b));
      if (FILLPROJ) {
        st.aux = __ldcs(reinterpret_cast<const uint4*>(pid + b));
      } else {
        st.v0 = __ldcs(reinterpret_cast<const double2*>(vals + b));
        st.v1 = __ldcs(reinterpret_cast<const double2*>(vals + b + 2));
      }
    }
  };
  const uint64_t tstride = gridDim.x;
  const uint64_t toff = (uint64_t)threadIdx.x * EPT;
  Stage cur, nx1, nx2;
  cur.aux = nx1.aux = nx2.aux = make_uint4(0u, 0u, 0u, 0u);
  cur.v0 = cur.v1 = nx1.v0 = nx1.v1 = nx2.v0 = nx2.v1 = make_double2(0.0, 0.0);
  uint64_t base = tile * TILE + toff;
  load_stage(cur, base);
  load_stage(nx1, base + tstride * TILE);
  uint32_t iter = 0;
  for (; tile < ntiles; tile += tstride, ++iter) {
    base = tile * TILE + toff;
    load_stage(nx2, base + 2 * tstride * TILE);
    const uint4 clab = cur.lab, caux = cur.aux;
    const double2 cv0 = cur.v0, cv1 = cur.v1;
    if ((iter & 15u) == 15u && ld_vol32(a.gmeta + 1)) return;
    const unsigned wmask = __ballot_sync(0xffffffffu, base < total);
    if (base < total) {
      const uint32_t lv[EPT] = {clab.x, clab.y, clab.z, clab.w};
      const uint32_t av[EPT] = {caux.x, caux.y, caux.z, caux.w};
      const double vv[EPT] = {cv0.x, cv0.y, cv1.x, cv1.y};
      uint32_t khi[EPT], klo[EPT], gid[EPT];
      double r[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const double v = FILLPROJ ? __dsub_rn(__ldg(lut + lv[e]), __ldg(tpat + av[e])) : vv[e];
        const uint32_t hi = (uint32_t)__double2hiint(v);
        if (fabs(v) < atol) {
          khi[e] = 0u;
          klo[e] = lv[e];
          r[e] = 0.0;
        } else {
          uint32_t e11 = (hi >> 20) & 0x7ffu;
          const double x = __hiloint2double((int)((hi & 0x800fffffu) | 0x3fe00000u), __double2loint(v));
          const int q = __double2int_rz(__dmul_rn(1.0e7, x));
          if (WRITEBACK) {
            const double y = __ddiv_rn((double)q, 1.0e7);
            const int n = (int)e11 - 1022;
            const int n1 = n / 2, n2 = n - n1;
            r[e] = y * __hiloint2double((1023 + n1) << 20, 0) * __hiloint2double((1023 + n2) << 20, 0);
          }
          uint32_t aq = (uint32_t)(q < 0 ? -q : q);
          if (aq == 10000000u) {
            aq = 5000000u;
            e11 += 1u;
          }
          khi[e] = (hi & 0x80000000u) | (e11 << 20) | (aq >> 4);
          klo[e] = (aq << 28) | lv[e];
        }
      }
      if (WRITEBACK) {
        __stcs(reinterpret_cast<double2*>(vals_out + base), make_double2(r[0], r[1]));
        __stcs(reinterpret_cast<double2*>(vals_out + base + 2), make_double2(r[2], r[3]));
      }
      if (!JOINT) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          // Lanes need different probe counts; without the explicit __syncwarp below each lane
          // would walk the rest of the tile alone (measured: 3.4 active lanes per instruction).
          const bool same = e > 0 && khi[e] == khi[e - 1] && klo[e] == klo[e - 1];
          const bool need = (khi[e] | klo[e]) != 0u && !same;
          uint32_t g = same ? gid[e - 1] : 0u;
          uint32_t s = (((klo[e] * 0x9e3779b1u) ^ (khi[e] * 0x85ebca77u)) >> 19) & (SC - 1);
          int free_slot = -1;
          if (need && use_cache) {
#pragma unroll 1
            for (int probe = 0; probe < CACHE_PROBES; ++probe) {
              uint4 raw;
              asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                           : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w)
                           : "r"(cache_base + s * 16u));
              if (raw.x == klo[e] && raw.y == khi[e]) {
                g = raw.z;                                   // 0 while the publisher is in flight
                if (g != 0u && iter <= raw.w) {
                  const uint32_t idx = a.idx0 + (uint32_t)(base + e);
                  if (idx < ld_vol32(&a.gslots[g - 1].minidx)) atomicMin(&a.gslots[g - 1].minidx, idx);
                }
                break;
              }
              if ((raw.x & raw.y) == 0xffffffffu) {
                free_slot = (int)s;
                break;
              }
              s = (s + 1) & (SC - 1);
            }
          }
          __syncwarp(wmask);       // lanes leave the probe loop at different times: reconverge here
          if (need && g == 0u)
            g = refine_miss(a.gslots, a.gocc, a.gmeta, a.gmask, a.glimit, cache, &s_count,
                            ((uint64_t)khi[e] << 32) | klo[e], a.idx0 + (uint32_t)(base + e), free_slot, iter);
          __syncwarp(wmask);
          gid[e] = g;
        }
      } else {
        // The four entries of a thread are resolved TOGETHER: every probe round issues the (predicated) 16-byte
        // shared loads of all entries that are still searching before any of them is compared, so the rounds of
        // a warp cost max-over-lanes iterations ONCE instead of once per entry, with four independent loads in
        // flight (the sequential version was bound by branch resolution and fixed-latency waits at ~3000 classes:
        // 730 warp instructions per 128 entries against 295 at 8 classes).
        bool same[EPT], need[EPT], pend[EPT];
        uint32_t g[EPT], slot[EPT];
        int free_slot[EPT];
        bool any = false;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          same[e] = e > 0 && khi[e] == khi[e - 1] && klo[e] == klo[e - 1];
          need[e] = (khi[e] | klo[e]) != 0u && !same[e];
          g[e] = 0u;
          free_slot[e] = -1;
          slot[e] = (((klo[e] * 0x9e3779b1u) ^ (khi[e] * 0x85ebca77u)) >> 19) & (SC - 1);
          pend[e] = need[e] && use_cache;
          any |= pend[e];
        }
#pragma unroll 1
        for (int probe = 0; probe < CACHE_PROBES && __any_sync(wmask, any); ++probe) {
          uint4 raw[EPT];
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            raw[e] = make_uint4(0u, 0u, 0u, 0u);
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t"
                "@p ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
                : "+r"(raw[e].x), "+r"(raw[e].y), "+r"(raw[e].z), "+r"(raw[e].w)
                : "r"(cache_base + slot[e] * 16u), "r"((uint32_t)pend[e]));
          }
          any = false;
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            if (pend[e]) {
              if (raw[e].x == klo[e] && raw[e].y == khi[e]) {
                g[e] = raw[e].z;                             // 0 while the publisher is in flight
                if (g[e] != 0u && iter <= raw[e].w) {
                  const uint32_t idx = a.idx0 + (uint32_t)(base + e);
                  if (idx < ld_vol32(&a.gslots[g[e] - 1].minidx)) atomicMin(&a.gslots[g[e] - 1].minidx, idx);
                }
                pend[e] = false;
              } else if ((raw[e].x & raw[e].y) == 0xffffffffu) {
                free_slot[e] = (int)slot[e];
                pend[e] = false;
              } else {
                slot[e] = (slot[e] + 1) & (SC - 1);
                any = true;
              }
            }
          }
        }
        __syncwarp(wmask);
        if (!use_cache) {
          // More classes than the CTA cache holds: every entry goes to the global table.  The first probe and the
          // first-index check of the four entries are issued together (two dependent L2 round trips per tile
          // instead of eight); whatever is not settled by the first probe takes the out-of-line path below.
          uint4 sl[EPT];
          uint32_t gs[EPT];
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            const uint64_t k = ((uint64_t)khi[e] << 32) | klo[e];
            gs[e] = (uint32_t)(mix64(k) >> 20) & a.gmask;
            sl[e] = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
            if (need[e]) sl[e] = ld_vol128(a.gslots + gs[e]);   // {key lo, key hi, minidx, -}
          }
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            if (need[e] && sl[e].x == klo[e] && sl[e].y == khi[e]) {
              g[e] = gs[e] + 1u;
              // (a 16-byte load need not be single-copy atomic: a stale or torn first index can only be LARGER
              //  than the true one, which costs a redundant atomicMin, never a missed one)
              const uint32_t idx = a.idx0 + (uint32_t)(base + e);
              if (idx < sl[e].z) atomicMin(&a.gslots[gs[e]].minidx, idx);
            }
          }
          __syncwarp(wmask);
        }
        // misses (first sight of a key in this CTA, or a publisher still in flight): in entry order, because an
        // entry equal to its predecessor takes the predecessor's id
        bool miss = false;
#pragma unroll
        for (int e = 0; e < EPT; ++e) miss |= need[e] && g[e] == 0u;
        if (__any_sync(wmask, miss)) {
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            if (need[e] && g[e] == 0u)
              g[e] = refine_miss(a.gslots, a.gocc, a.gmeta, a.gmask, a.glimit, cache, &s_count,
                                 ((uint64_t)khi[e] << 32) | klo[e], a.idx0 + (uint32_t)(base + e), free_slot[e], iter);
            __syncwarp(wmask);
          }
        }
        gid[0] = g[0];
#pragma unroll
        for (int e = 1; e < EPT; ++e) gid[e] = same[e] ? gid[e - 1] : g[e];
      }
      __stcs(reinterpret_cast<uint4*>(lab_out + base), make_uint4(gid[0], gid[1], gid[2], gid[3]));
    }
    cur = nx1;
    nx1 = nx2;
  }
}

// ---------------------------------------------------------------------------
// canonical ranking of a table: rank[slot+1] = 1 + #{keys whose first index is smaller}
// ---------------------------------------------------------------------------
__global__ void gather_minidx_kernel(const uint32_t* __restrict__ occ, const KeySlot* __restrict__ gslots,
                                     uint32_t* __restrict__ mi, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) mi[i] = gslots[occ[i]].minidx;
}

__global__ void __launch_bounds__(256) rank_brute_kernel(const uint32_t* __restrict__ occ,
                                                         const uint32_t* __restrict__ mi,
                                                         uint32_t* __restrict__ rank, uint32_t count) {
  __shared__ uint32_t tile[1024];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t mine = i < count ? mi[i] : 0u;
  uint32_t r = 0;
  for (uint32_t j0 = 0; j0 < count; j0 += 1024) {
    for (uint32_t t = threadIdx.x; t < 1024; t += blockDim.x) tile[t] = (j0 + t < count) ? mi[j0 + t] : 0xffffffffu;
    __syncthreads();
    const uint32_t lim = min(1024u, count - j0);
    for (uint32_t t = 0; t < lim; ++t) r += (tile[t] < mine) ? 1u : 0u;
    __syncthreads();
  }
  if (i < count) rank[occ[i] + 1] = r + 1;
}

__global__ void bitmap_set_kernel(const uint32_t* __restrict__ occ, const KeySlot* __restrict__ gslots,
                                  uint32_t* __restrict__ bitmap, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    const uint32_t idx = gslots[occ[i]].minidx;
    atomicOr(bitmap + (idx >> 5), 1u << (idx & 31u));
  }
}

// one CTA of 1024 threads per 1024 bitmap words: exclusive popcount prefix per word + block total
__global__ void __launch_bounds__(1024) bitmap_scan_kernel(const uint32_t* __restrict__ bitmap,
                                                           uint32_t* __restrict__ wordpre,
                                                           uint32_t* __restrict__ blocksum, uint64_t nwords) {
  __shared__ uint32_t wsum[32];
  const uint64_t w = (uint64_t)blockIdx.x * 1024 + threadIdx.x;
  const uint32_t c = w < nwords ? __popc(bitmap[w]) : 0u;
  uint32_t incl = c;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t v = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    wsum[lane] = v;   // inclusive over warps
  }
  __syncthreads();
  const uint32_t warp_off = wid ? wsum[wid - 1] : 0u;
  if (w < nwords) wordpre[w] = warp_off + incl - c;
  if (threadIdx.x == 1023) blocksum[blockIdx.x] = warp_off + incl;
}

// single CTA: exclusive scan of the block totals in place
__global__ void __launch_bounds__(1024) blocksum_scan_kernel(uint32_t* __restrict__ blocksum, uint32_t nblocks) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (uint32_t b0 = 0; b0 < nblocks; b0 += 1024) {
    const uint32_t i = b0 + threadIdx.x;
    const uint32_t c = i < nblocks ? blocksum[i] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      uint32_t v = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      wsum[lane] = v;
    }
    __syncthreads();
    const uint32_t off = carry + (wid ? wsum[wid - 1] : 0u);
    if (i < nblocks) blocksum[i] = off + incl - c;
    __syncthreads();
    if (threadIdx.x == 1023) carry = off + incl;
    __syncthreads();
  }
}

__global__ void bitmap_rank_kernel(const uint32_t* __restrict__ occ, const KeySlot* __restrict__ gslots,
                                   const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ wordpre,
                                   const uint32_t* __restrict__ blocksum, uint32_t* __restrict__ rank,
                                   uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    const uint32_t s = occ[i];
    const uint32_t idx = gslots[s].minidx;
    const uint32_t w = idx >> 5;
    const uint32_t below = __popc(bitmap[w] & ((1u << (idx & 31u)) - 1u));
    rank[s + 1] = blocksum[w >> 10] + wordpre[w] + below + 1u;
  }
}

// ---------------------------------------------------------------------------
// fill! (src/partitions.jl:68-75): lut composition and the label -> value gather
// ---------------------------------------------------------------------------
__global__ void build_lut_kernel(const uint32_t* __restrict__ occ, const uint32_t* __restrict__ rank,
                                 const double* __restrict__ values, double* __restrict__ lut, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) lut[0] = 0.0;
  if (i < count) {
    const uint32_t p = occ[i] + 1;
    lut[p] = values[rank[p] - 1];
  }
}

// lut[slot + 1] = the rounded value encoded in the key of `slot` (the value every entry of that class holds
// after `_clamp_round!`): after a round+refine pass the rounded matrix IS fill(S_new, lut), so it never has
// to be written to HBM (src/partitions.jl:160-164: "X stays the projected, rounded element").
//   fast layout   hi = sign | e11 | q>>4,  lo = (q & 15) << 28 | old id          (refine_fast_kernel)
//   generic       key = (sign | e11 | q) << lbits | old id                        (refine_kernel<KM_ROUND>)
__global__ void decode_lut_kernel(const uint32_t* __restrict__ occ, const KeySlot* __restrict__ slots, uint32_t count,
                                  int fast, int lbits, int qbits, double scale, double* __restrict__ lut) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) lut[0] = 0.0;
  if (i >= count) return;
  const uint32_t s = occ[i];
  const uint64_t k = slots[s].key;
  uint64_t aq, e11, sign;
  if (fast) {
    const uint32_t hi = (uint32_t)(k >> 32), lo = (uint32_t)k;
    aq = ((uint64_t)(hi & 0xFFFFFu) << 4) | (lo >> 28);
    e11 = (hi >> 20) & 0x7ffu;
    sign = hi >> 31;
  } else {
    const uint64_t code = k >> lbits;
    aq = code & ((1ull << qbits) - 1ull);
    e11 = (code >> qbits) & 0x7ffull;
    sign = (code >> (11 + qbits)) & 1ull;
  }
  double v = 0.0;
  if (aq != 0ull || e11 != 0ull) {
    const double y = __ddiv_rn((double)(long long)aq, scale);
    const int n = (int)e11 - 1022;
    const int n1 = n / 2, n2 = n - n1;
    v = y * __longlong_as_double((long long)(1023 + n1) << 52) * __longlong_as_double((long long)(1023 + n2) << 52);
    if (sign) v = -v;
  }
  lut[s + 1] = v;
}

// out[0] = max |lut| over the occupied ids, out[1] = min non-zero |lut| (bit patterns of non-negative doubles)
__global__ void lut_stats_kernel(const uint32_t* __restrict__ occ, const double* __restrict__ lut, uint32_t count,
                                 unsigned long long* __restrict__ out) {
  unsigned long long mx = 0ull, mn = ~0ull;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const unsigned long long a = (unsigned long long)__double_as_longlong(lut[occ[i] + 1]) & 0x7FFFFFFFFFFFFFFFull;
    mx = max(mx, a);
    if (a) mn = min(mn, a);
  }
  for (int o = 16; o; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (mx) atomicMax(out, mx);
    atomicMin(out + 1, mn);
  }
}

__global__ void __launch_bounds__(256) fill_kernel(const uint32_t* __restrict__ labels,
                                                   const double* __restrict__ lut, double* __restrict__ X,
                                                   uint64_t total) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
  for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; base < total; base += stride) {
    const uint4 l = __ldcs(reinterpret_cast<const uint4*>(labels + base));
    const double a = __ldg(lut + l.x), b = __ldg(lut + l.y), c = __ldg(lut + l.z), d = __ldg(lut + l.w);
    __stcs(reinterpret_cast<double2*>(X + base), make_double2(a, b));
    __stcs(reinterpret_cast<double2*>(X + base + 2), make_double2(c, d));
  }
}

// canonical labels, unpadded: out[i + n*j] = rank[labels[i + ld*j]]
__global__ void canonical_kernel(const uint32_t* __restrict__ labels, const uint32_t* __restrict__ rank,
                                 uint32_t* __restrict__ out, int64_t n, int64_t ld) {
  const int64_t j = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i + n * j] = rank[labels[i + ld * j]];
}

}  // namespace

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
void sdpsr_table_free(KeyTable& t) {
  cudaFree(t.slots);
  cudaFree(t.rank);
  cudaFree(t.occ);
  cudaFree(t.meta);
  t = KeyTable();
}

int sdpsr_table_alloc(sdpsr_ctx* ctx, KeyTable& t, size_t cap) {
  if (t.alloc >= cap && t.slots) {
    t.cap = (uint32_t)cap;
    return SDPSR_OK;
  }
  sdpsr_table_free(t);
  // a single memset(0xff) clears a slot: key = KEY_EMPTY, first index = 2^32 - 1
  SDPSR_CUDA(cudaMalloc(&t.slots, cap * sizeof(KeySlot)));
  SDPSR_CUDA(cudaMalloc(&t.rank, (cap + 1) * sizeof(uint32_t)));
  SDPSR_CUDA(cudaMalloc(&t.occ, cap * sizeof(uint32_t)));
  SDPSR_CUDA(cudaMalloc(&t.meta, 4 * sizeof(uint32_t)));
  t.alloc = cap;
  t.cap = (uint32_t)cap;
  return SDPSR_OK;
}

int sdpsr_round_params(sdpsr_ctx* ctx, double atol, double* scale, long long* iscale, int* qbits) {
  SDPSR_REQUIRE(atol > 1e-300 && atol <= 0.1 && std::isfinite(atol), SDPSR_E_INVALID,
                "atol must be in (1e-300, 0.1]");
  const int sig = (int)std::floor(-std::log10(atol));   // src/utils.jl:37
  SDPSR_REQUIRE(sig >= 1 && sig <= 15, SDPSR_E_INVALID, "floor(-log10(atol)) must be in 1..15");
  long long s = 1;
  for (int i = 0; i < sig; ++i) s *= 10;
  *iscale = s;
  *scale = (double)s;
  *qbits = bits_for((uint64_t)s);
  return SDPSR_OK;
}

static size_t initial_cap(sdpsr_ctx* ctx) {
  if (ctx->flags & SDPSR_F_TINY_TABLE) return 64;
  size_t want = next_pow2(2 * (uint64_t)ctx->elems);
  return std::max<size_t>(64, std::min<size_t>(want, (size_t)1 << 20));
}

int sdpsr_rank_table(sdpsr_ctx* ctx, KeyTable& t) {
  const uint32_t count = t.count;
  Timed tm(ctx, SDPSR_K_RANK, (double)count * 8.0);
  SDPSR_CUDA(cudaMemsetAsync(t.rank, 0, sizeof(uint32_t), ctx->stream));
  if (count == 0) return SDPSR_OK;
  const KeySlot* gmin = t.slots;
  const int blocks = (int)((count + 255) / 256);
  const bool brute = count <= (uint32_t)RANK_BRUTE_MAX && !(ctx->flags & SDPSR_F_FORCE_BITMAP_RANK);
  if (brute) {
    if (ctx->rk_alloc < count) {
      cudaFree(ctx->rk_mi);
      ctx->rk_mi = nullptr;
      const size_t want = std::max<size_t>(count, 4096);
      SDPSR_CUDA(cudaMalloc(&ctx->rk_mi, want * sizeof(uint32_t)));
      ctx->rk_alloc = want;
    }
    gather_minidx_kernel<<<blocks, 256, 0, ctx->stream>>>(t.occ, gmin, ctx->rk_mi, count);
    rank_brute_kernel<<<blocks, 256, 0, ctx->stream>>>(t.occ, ctx->rk_mi, t.rank, count);
    count_launch(ctx, 2);
  } else {
    const uint64_t nwords = (ctx->elems + 31) / 32;
    const uint64_t nblk = (nwords + 1023) / 1024;
    if (!ctx->bitmap) {
      SDPSR_CUDA(cudaMalloc(&ctx->bitmap, nwords * 2 * sizeof(uint32_t)));   // bitmap + wordpre
      SDPSR_CUDA(cudaMalloc(&ctx->bm_block, nblk * sizeof(uint32_t)));
      ctx->bm_blocks = nblk;
    }
    uint32_t* wordpre = ctx->bitmap + nwords;
    SDPSR_CUDA(cudaMemsetAsync(ctx->bitmap, 0, nwords * sizeof(uint32_t), ctx->stream));
    bitmap_set_kernel<<<blocks, 256, 0, ctx->stream>>>(t.occ, gmin, ctx->bitmap, count);
    bitmap_scan_kernel<<<(unsigned)nblk, 1024, 0, ctx->stream>>>(ctx->bitmap, wordpre, ctx->bm_block, nwords);
    blocksum_scan_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->bm_block, (uint32_t)nblk);
    bitmap_rank_kernel<<<blocks, 256, 0, ctx->stream>>>(t.occ, gmin, ctx->bitmap, wordpre, ctx->bm_block, t.rank,
                                                        count);
    count_launch(ctx, 4);
  }
  SDPSR_CUDA(cudaGetLastError());
  return SDPSR_OK;
}

int sdpsr_refine_pass(sdpsr_ctx* ctx, const RefineSpec& spec, int64_t* dim) {
  static bool attr_set_dev[64] = {false};      // the attribute is per device
  bool& attr_set = attr_set_dev[ctx->device & 63];
  const size_t smem = (size_t)SC * 16;
  if (!attr_set) {
    SDPSR_CUDA(cudaFuncSetAttribute(refine_kernel<KM_ROUND, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDPSR_CUDA(cudaFuncSetAttribute(refine_kernel<KM_ROUND, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDPSR_CUDA(cudaFuncSetAttribute(refine_kernel<KM_RAW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDPSR_CUDA(cudaFuncSetAttribute(refine_kernel<KM_RAW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDPSR_CUDA(cudaFuncSetAttribute(refine_kernel<KM_PAIR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDPSR_CUDA(cudaFuncSetAttribute(refine_fast_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDPSR_CUDA(cudaFuncSetAttribute(refine_fast_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDPSR_CUDA(cudaFuncSetAttribute(refine_fast_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDPSR_CUDA(cudaFuncSetAttribute(refine_fast_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDPSR_CUDA(cudaFuncSetAttribute(refine_fast_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDPSR_CUDA(cudaFuncSetAttribute(refine_fast_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDPSR_CUDA(cudaFuncSetAttribute(refine_fast_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SDPSR_CUDA(cudaFuncSetAttribute(refine_fast_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  KeyTable& tnew = spec.table_override ? *spec.table_override : ctx->tab[ctx->cur ^ 1];

  // Sharded partition (shard.cu): the pass runs over this rank's column block only and is followed by the
  // key-table merge.  Passes with an output / table override (pattern ids, the id pass of the two-step
  // refine) are rank-local full-range passes.
  const bool sharded = sdpsr_shard_active(ctx) && ((!spec.out_override && !spec.table_override) || spec.shard_block);
  uint64_t rb = 0, re = ctx->elems;
  if (sharded) sdpsr_shard_block(ctx, ctx->rank, &rb, &re);
  else if (sdpsr_shard_active(ctx) && !spec.ignore_labels) SDPSR_TRY(sdpsr_shard_ensure_full_labels(ctx));
  RefineArgs a;
  std::memset(&a, 0, sizeof(a));
  a.lab_in = spec.ignore_labels ? nullptr : ctx->labels + rb;
  a.lab2 = spec.lab2 ? spec.lab2 + rb : nullptr;
  a.vals = spec.vals ? spec.vals + rb : nullptr;
  a.vals_out = spec.vals_out ? spec.vals_out + rb : nullptr;
  a.lut = spec.lut;
  a.tpat = spec.tpat;
  a.pid = spec.pid ? spec.pid + rb : nullptr;
  a.lab_out = (spec.out_override ? spec.out_override : ctx->labels_alt) + rb;
  a.total = re - rb;
  a.idx0 = (uint32_t)rb;
  a.atol = spec.atol;
  a.do_round = spec.do_round ? 1 : 0;
  a.fillproj = spec.fillproj ? 1 : 0;
  a.raw_bits = spec.raw_bits ? 1 : 0;
  // the per-CTA cache only pays while the classes fit it; the new dim is >= the current one
  // (SDPSR_REFINE_CACHE_LIMIT / SDPSR_REFINE_JOINT_MIN: experiment knobs of tools/refine_bench.py)
  static const int64_t cache_limit = [] {
    const char* e = getenv("SDPSR_REFINE_CACHE_LIMIT");
    return e ? std::min<int64_t>(atoll(e), SC_LIMIT) : (int64_t)CACHE_DIM_LIMIT;
  }();
  static const int64_t joint_min = [] {
    const char* e = getenv("SDPSR_REFINE_JOINT_MIN");
    return e ? (int64_t)atoll(e) : (int64_t)JOINT_MIN_DIM;
  }();
  a.use_cache = ((ctx->flags & SDPSR_F_NO_SMEM_CACHE) || (!spec.table_override && ctx->dim > cache_limit)) ? 0 : 1;
  // provisional ids are <= cap; sharded: canonical ids <= dim, and the key layout must not depend on a
  // rank's table capacity (keys are compared across ranks)
  a.lbits = spec.ignore_labels ? 1 : sdpsr_label_bits(ctx);
  if (spec.do_round && spec.mode != KM_PAIR) SDPSR_TRY(sdpsr_round_params(ctx, spec.atol, &a.scale, &a.iscale, &a.qbits));
  if (spec.mode == KM_ROUND) {
    SDPSR_REQUIRE(spec.do_round, SDPSR_E_INVALID, "KM_ROUND needs rounding");
    SDPSR_REQUIRE(12 + a.qbits + a.lbits <= 64, SDPSR_E_UNSUPPORTED, "fused key does not fit 64 bits");
  }

  size_t cap = tnew.cap ? tnew.cap : initial_cap(ctx);
  if (cap < initial_cap(ctx)) cap = initial_cap(ctx);
  const size_t cap_max = std::max<size_t>(64, next_pow2(2 * (uint64_t)ctx->elems));
  const uint64_t ntiles = (a.total + TILE - 1) / TILE;
  const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(ntiles, (uint64_t)ctx->sm_count));
  bool last_fast = false;

  for (;;) {
    SDPSR_TRY(sdpsr_table_alloc(ctx, tnew, cap));
    SDPSR_CUDA(cudaMemsetAsync(tnew.slots, 0xff, (size_t)tnew.cap * sizeof(KeySlot), ctx->stream));
    SDPSR_CUDA(cudaMemsetAsync(tnew.meta, 0, 4 * sizeof(uint32_t), ctx->stream));
    a.gslots = tnew.slots;
    a.gocc = tnew.occ;
    a.gmeta = tnew.meta;
    a.gmask = tnew.cap - 1;
    a.glimit = tnew.cap / 2;
    {
      Timed tm(ctx, SDPSR_K_REFINE, (double)a.total * 16.0);
      const bool wb = a.vals_out != nullptr;
      const bool fast = spec.mode == KM_ROUND && a.iscale == 10000000ll && a.lbits <= 28 &&
                        !(ctx->flags & SDPSR_F_NO_SMEM_CACHE);
      last_fast = fast;
      if (fast) {
        // few classes: nearly every lane hits on its first probe and the entry-by-entry loop is the shorter code;
        // many classes: the joint probe rounds win (sweep at N = 16384: 216 classes 4.72 vs 4.44 TB/s,
        // 3000 classes 1.79 vs 1.93 TB/s)
        const bool joint = ctx->dim > joint_min;
#define SDPSR_LAUNCH_FAST(WB, FP)                                                      \
  do {                                                                                 \
    if (joint) refine_fast_kernel<WB, FP, true><<<grid, RT, smem, ctx->stream>>>(a);   \
    else refine_fast_kernel<WB, FP, false><<<grid, RT, smem, ctx->stream>>>(a);        \
  } while (0)
        if (a.fillproj) {
          if (wb) SDPSR_LAUNCH_FAST(true, true);
          else SDPSR_LAUNCH_FAST(false, true);
        } else {
          if (wb) SDPSR_LAUNCH_FAST(true, false);
          else SDPSR_LAUNCH_FAST(false, false);
        }
#undef SDPSR_LAUNCH_FAST
      } else
      switch (spec.mode) {
        case KM_ROUND:
          if (wb) refine_kernel<KM_ROUND, true><<<grid, RT, smem, ctx->stream>>>(a);
          else refine_kernel<KM_ROUND, false><<<grid, RT, smem, ctx->stream>>>(a);
          break;
        case KM_RAW:
          if (wb) refine_kernel<KM_RAW, true><<<grid, RT, smem, ctx->stream>>>(a);
          else refine_kernel<KM_RAW, false><<<grid, RT, smem, ctx->stream>>>(a);
          break;
        case KM_PAIR: refine_kernel<KM_PAIR, false><<<grid, RT, smem, ctx->stream>>>(a); break;
      }
      count_launch(ctx);
    }
    SDPSR_CUDA(cudaGetLastError());
    uint32_t* hm = reinterpret_cast<uint32_t*>(ctx->h_pinned);
    SDPSR_CUDA(cudaMemcpyAsync(hm, tnew.meta, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hm[1] == 0u) {
      tnew.count = hm[0];
      break;
    }
    SDPSR_REQUIRE(cap < cap_max, SDPSR_E_ALLOC, "key table overflow at maximum capacity");
    cap = std::min(cap * 4, cap_max);
  }
  if (sharded) {
    int64_t dg = 0;
    SDPSR_TRY(sdpsr_shard_merge(ctx, tnew, spec.out_override ? spec.out_override : ctx->labels_alt, &dg));
    if (!spec.out_override) {
      ctx->labels_full = false;
      ctx->clabels_valid = false;
    }
  } else {
    SDPSR_TRY(sdpsr_rank_table(ctx, tnew));
  }
  if (!spec.out_override) {
    std::swap(ctx->labels, ctx->labels_alt);
    ctx->cur ^= 1;
    ctx->dim = tnew.count;
    ctx->part_epoch += 1;
    ctx->x_is_fill = false;
    // value-coded keys: the class values can be decoded from the table (sdpsr_decode_lut)
    ctx->key_decodable = spec.mode == KM_ROUND;
    ctx->key_fast = last_fast;
    ctx->key_lbits = a.lbits;
    ctx->sym_state = spec.keeps_symmetry && ctx->sym_state == 1 ? 1 : -1;
  }
  if (dim) *dim = tnew.count;
  return SDPSR_OK;
}

// Bits of the old-label field of a key.  Provisional ids are <= cap; in a sharded run the labels are canonical
// (<= dim) and the layout must not depend on a rank's table capacity: keys are compared across ranks, and the
// choice between the fused pass and the two-step refine (which contains collectives) must be the same everywhere.
int sdpsr_label_bits(const sdpsr_ctx* ctx) {
  return sdpsr_shard_active(ctx) ? bits_for((uint64_t)ctx->dim + 1) : bits_for((uint64_t)ctx->tab[ctx->cur].cap);
}

int sdpsr_ensure_tmp_labels(sdpsr_ctx* ctx) {
  if (!ctx->labels_tmp) {
    SDPSR_CUDA(cudaMalloc(&ctx->labels_tmp, ctx->elems * sizeof(uint32_t)));
    SDPSR_CUDA(cudaMemsetAsync(ctx->labels_tmp, 0, ctx->elems * sizeof(uint32_t), ctx->stream));
  }
  return SDPSR_OK;
}

// S = refine!(S, Partition(round?(M))) for a device-resident padded matrix.
// Fused single pass when the (id, value code) pair fits 64 bits, otherwise the
// reference's own two steps: ids of Part(M) first, then the pair pass.
int sdpsr_generic_refine_values(sdpsr_ctx* ctx, const double* dvals, double atol, bool do_round,
                                double* vals_out, int64_t* dim, bool keeps_symmetry) {
  bool fused = do_round;
  if (fused) {
    double sc;
    long long isc;
    int qb;
    SDPSR_TRY(sdpsr_round_params(ctx, atol, &sc, &isc, &qb));
    fused = 12 + qb + sdpsr_label_bits(ctx) <= 64;
  }
  RefineSpec sp;
  sp.vals = dvals;
  sp.vals_out = vals_out;
  sp.atol = atol;
  sp.do_round = do_round;
  sp.keeps_symmetry = keeps_symmetry;
  if (fused) {
    sp.mode = KM_ROUND;
    return sdpsr_refine_pass(ctx, sp, dim);
  }
  if (ctx->dim == 0 && ctx->tab[ctx->cur].count == 0) {   // empty S: Part(M) itself
    sp.mode = KM_RAW;
    sp.ignore_labels = true;
    return sdpsr_refine_pass(ctx, sp, dim);
  }
  SDPSR_TRY(sdpsr_ensure_tmp_labels(ctx));
  KeyTable& scratch = ctx->tab_scratch;   // reused across calls, freed with the context
  sp.mode = KM_RAW;
  sp.ignore_labels = true;
  sp.out_override = ctx->labels_tmp;
  sp.table_override = &scratch;
  sp.shard_block = true;      // sharded run: ids of Part(M) on this rank's block, merged to canonical ids
  int64_t d2 = 0;
  int st = sdpsr_refine_pass(ctx, sp, &d2);
  if (st == SDPSR_OK) {
    RefineSpec pr;
    pr.mode = KM_PAIR;
    pr.lab2 = ctx->labels_tmp;
    pr.do_round = false;
    pr.keeps_symmetry = keeps_symmetry;
    st = sdpsr_refine_pass(ctx, pr, dim);
  }
  return st;
}

int sdpsr_upload_values(sdpsr_ctx* ctx, const double* values, int64_t len) {
  if (ctx->values_alloc < (size_t)len) {
    cudaFree(ctx->d_values);
    ctx->d_values = nullptr;
    const size_t want = std::max<size_t>((size_t)len, 4096);
    SDPSR_CUDA(cudaMalloc(&ctx->d_values, want * sizeof(double)));
    ctx->values_alloc = want;
  }
  if (len > 0)
    SDPSR_CUDA(cudaMemcpyAsync(ctx->d_values, values, (size_t)len * sizeof(double), cudaMemcpyDefault, ctx->stream));
  return SDPSR_OK;
}

static int ensure_lut(sdpsr_ctx* ctx, size_t need);

int sdpsr_build_lut(sdpsr_ctx* ctx, const double* d_values, int64_t len) {
  KeyTable& t = ctx->tab[ctx->cur];
  SDPSR_REQUIRE(len == (int64_t)t.count, SDPSR_E_INVALID, "length(values) != dim(P) (src/partitions.jl:69)");
  SDPSR_TRY(ensure_lut(ctx, (size_t)t.cap + 1));
  const int blocks = (int)std::max<uint32_t>(1, (t.count + 255) / 256);
  build_lut_kernel<<<blocks, 256, 0, ctx->stream>>>(t.occ, t.rank, d_values, ctx->lut, t.count);
  count_launch(ctx);
  SDPSR_CUDA(cudaGetLastError());
  return SDPSR_OK;
}

static int ensure_lut(sdpsr_ctx* ctx, size_t need) {
  if (ctx->lut_alloc < need) {
    cudaFree(ctx->lut);
    ctx->lut = nullptr;
    SDPSR_CUDA(cudaMalloc(&ctx->lut, need * sizeof(double)));
    ctx->lut_alloc = need;
  }
  return SDPSR_OK;
}

// After a fused round+refine pass (KM_ROUND keys): ctx->lut[provisional id] = the rounded value of the class.
int sdpsr_decode_lut(sdpsr_ctx* ctx, double atol) {
  KeyTable& t = ctx->tab[ctx->cur];
  SDPSR_REQUIRE(ctx->key_decodable, SDPSR_E_STATE, "the last refine pass did not use value-coded keys");
  SDPSR_TRY(ensure_lut(ctx, (size_t)t.cap + 1));
  double scale;
  long long iscale;
  int qbits;
  SDPSR_TRY(sdpsr_round_params(ctx, atol, &scale, &iscale, &qbits));
  const int blocks = (int)std::max<uint32_t>(1, (t.count + 255) / 256);
  decode_lut_kernel<<<blocks, 256, 0, ctx->stream>>>(t.occ, t.slots, t.count, ctx->key_fast ? 1 : 0, ctx->key_lbits, qbits,
                                                      scale, ctx->lut);
  count_launch(ctx);
  SDPSR_CUDA(cudaGetLastError());
  return SDPSR_OK;
}

// max |lut| and the smallest non-zero |lut| over the classes of S (host doubles)
int sdpsr_lut_stats(sdpsr_ctx* ctx, double* vmax, double* vmin_nz) {
  KeyTable& t = ctx->tab[ctx->cur];
  unsigned long long* d = reinterpret_cast<unsigned long long*>(ctx->d_scalars + 44);
  unsigned long long* h = reinterpret_cast<unsigned long long*>(ctx->h_pinned) + 44;
  h[0] = 0ull;
  h[1] = ~0ull;
  SDPSR_CUDA(cudaMemcpyAsync(d, h, 2 * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
  if (t.count) {
    const int blocks = (int)std::min<uint32_t>(64, (t.count + 255) / 256);
    lut_stats_kernel<<<blocks, 256, 0, ctx->stream>>>(t.occ, ctx->lut, t.count, d);
    count_launch(ctx);
  }
  SDPSR_CUDA(cudaMemcpyAsync(h, d, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  SDPSR_CUDA(cudaStreamSynchronize(ctx->stream));
  std::memcpy(vmax, h, sizeof(double));
  if (h[1] == ~0ull)
    *vmin_nz = 0.0;
  else
    std::memcpy(vmin_nz, h + 1, sizeof(double));
  return SDPSR_OK;
}

int sdpsr_materialize_fill(sdpsr_ctx* ctx, double* dst) {
  SDPSR_TRY(sdpsr_shard_ensure_full_labels(ctx));
  Timed tm(ctx, SDPSR_K_FILL, (double)ctx->elems * 12.0);
  const uint64_t per = 256 * 4;
  const int grid = (int)std::min<uint64_t>((ctx->elems + per - 1) / per, (uint64_t)ctx->sm_count * 8);
  fill_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->labels, ctx->lut, dst, ctx->elems);
  count_launch(ctx);
  SDPSR_CUDA(cudaGetLastError());
  return SDPSR_OK;
}

int sdpsr_canonical_labels(sdpsr_ctx* ctx, uint32_t* dst_unpadded) {
  SDPSR_TRY(sdpsr_shard_ensure_full_labels(ctx));
  KeyTable& t = ctx->tab[ctx->cur];
  dim3 grid((unsigned)std::min<int64_t>((ctx->n + 255) / 256, 64), (unsigned)ctx->n);
  canonical_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->labels, t.rank, dst_unpadded, ctx->n, ctx->ld);
  count_launch(ctx);
  SDPSR_CUDA(cudaGetLastError());
  return SDPSR_OK;
}
