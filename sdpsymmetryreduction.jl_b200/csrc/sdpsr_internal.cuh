// Internal declarations shared by the translation units of libsdpsr_cuda.so.
// Nothing here is part of the C ABI (include/sdpsr.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/sdpsr.h"

// --------------------------------------------------------------------------
// error plumbing
// --------------------------------------------------------------------------
#define SDPSR_CUDA(call)                                                               \
  do {                                                                                 \
    cudaError_t _e = (call);                                                           \
    if (_e != cudaSuccess) {                                                           \
      return ctx->fail(SDPSR_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
    }                                                                                  \
  } while (0)

#define SDPSR_TRY(expr)          \
  do {                           \
    int _s = (expr);             \
    if (_s != SDPSR_OK) return _s; \
  } while (0)

#define SDPSR_REQUIRE(cond, code, msg)              \
  do {                                              \
    if (!(cond)) return ctx->fail((code), (msg));   \
  } while (0)

// --------------------------------------------------------------------------
// key table: open-addressing hash  key -> (first occurrence idx, canonical rank)
// A "provisional id" of an entry is slot+1 of its key (0 is reserved for the
// zero key: old label 0 and value +0.0), see DESIGN.md "lazy canonical labels".
// --------------------------------------------------------------------------
constexpr uint64_t KEY_EMPTY = ~0ull;

// One slot of the table: the key and the first index of its class share a 16-byte slot, i.e. one 32-byte L2
// sector serves the lookup AND the first-occurrence check of an entry (round 1 kept them in two arrays: two
// sectors per entry, which is what bounded the pass once the classes no longer fit the per-CTA cache).
struct __align__(16) KeySlot {
  uint64_t key;      // KEY_EMPTY when free
  uint32_t minidx;   // smallest padded linear index holding the key
  uint32_t pad;
};

struct KeyTable {
  KeySlot* slots = nullptr;    // [cap]
  uint32_t* rank = nullptr;    // [cap + 1]  canonical label of provisional id (rank[0] = 0)
  uint32_t* occ = nullptr;     // [cap]      list of occupied slots, in claim order
  uint32_t* meta = nullptr;    // [4]        {count, overflow, -, -}
  uint32_t cap = 0;            // power of two
  uint32_t count = 0;          // host copy of meta[0] after the pass == dim
  size_t alloc = 0;            // allocated capacity
};

struct ConstraintSet {
  int64_t m = 0;
  int64_t nnz = 0;
  // CSR of A on the device (column index = padded linear index)
  uint32_t* d_col = nullptr;     // [nnz]
  double* d_val = nullptr;       // [nnz]
  uint32_t* d_chunk_row = nullptr;   // [nchunks] row of each reduction chunk
  uint32_t* d_chunk_beg = nullptr;   // [nchunks+1] nnz range of each chunk
  int64_t nchunks = 0;
  double* d_partial = nullptr;   // [nchunks]
  // per-entry pattern ids and the pattern table
  uint32_t* d_pid = nullptr;     // [elems]  0 = entry touched by no constraint
  int64_t npat = 0;              // patterns 1..npat
  std::vector<int64_t> pat_ptr;  // [npat+2]
  std::vector<int32_t> pat_row;
  std::vector<double> pat_val;
  std::vector<int64_t> pat_cnt;  // [npat+1]
  std::vector<uint32_t> chunk_row;   // host copy
  // sharded row dots (C3): the chunks of every rank's column block; d_sh_* hold this rank's
  int sh_nranks = 0, sh_rank = -1;
  std::vector<std::vector<uint32_t>> sh_chunk_row;   // [rank][chunk] constraint row
  size_t sh_maxchunks = 0;
  uint32_t* d_sh_beg = nullptr;      // [2 * nchunks of this rank] begin | end
  double* d_sh_partial = nullptr;    // [nranks * sh_maxchunks]
  std::vector<long double> gram_chol;   // m x m pivoted Cholesky factor (row-major) of A A'
  std::vector<int> gram_perm;           // pivot order: gram_perm[i] = constraint row eliminated i-th
  int gram_rank = 0;                    // numerical rank of A; rows gram_perm[rank..] are dependent (dropped)
  double* d_tpat = nullptr;      // [npat+1]
  int pid_sym = -1;              // pattern ids transpose-invariant? (checked once; decides whether the
                                 // projection keeps a symmetric partition symmetric)
  std::vector<int64_t> h_rowptr; // host copy of the row pointers (the stored entries live on the device only)
  int64_t* d_rowptr = nullptr;   // [m + 1]
  bool ready = false;
};

struct EventPair {
  cudaEvent_t a, b;
  int family;
  double work;
};

struct sdpsr_ctx {
  int device = 0;
  int64_t n = 0;        // matrix order N
  int64_t ld = 0;       // leading dimension of every device matrix (N rounded up to 16)
  size_t elems = 0;     // ld * n
  uint32_t flags = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  int sm_count = 148;

  // Partition S: provisional ids, double buffered; tab[cur] maps them to canonical labels
  uint32_t* labels = nullptr;
  uint32_t* labels_alt = nullptr;
  uint32_t* labels_tmp = nullptr;   // lazily allocated third buffer (generic two-step refine, IO)
  KeyTable tab[2];
  KeyTable tab_scratch;             // reusable third table (two-step refine, pattern ids)
  KeyTable tab_merge;               // sharded refine: the merged (key -> first index) table (shard.cu)
  // Sharded partition (nranks > 1, shard.cu): this rank's column block of `labels` is always current; the
  // other blocks only while labels_full.  clabels = all blocks as 1- / 2-byte canonical labels (gathered
  // lazily, the INT8 square reads its digit slices through them).
  bool labels_full = true;
  uint8_t* clabels = nullptr;       // [elems * 2] bytes, lazily
  int clabel_width = 0;             // 1, 2 (clabels) or 4 (labels gathered directly)
  bool clabels_valid = false;
  int cur = 0;
  int64_t dim = 0;
  uint64_t part_epoch = 0;          // bumped whenever the partition changes (derived state checks it)

  // destination columns of the lower -> upper mirror after a symmetric product (sharded closure step: the
  // refine pass that follows reads this rank's column block of X2 only); col1 < 0: all columns
  int64_t mirror_col0 = 0, mirror_col1 = -1;
  int i8_pair = -1;     // INT8 square on CTA pairs (cta_group::2): -1 = for N > 16384, 0 / 1 = SDPSR_I8_PAIR
  int i8_segblocks = 0; // test hook (SDPSR_I8_SEGBLOCKS): K segment length of the INT8 square in 128-byte blocks
  int i8_slices = 0;    // int8 digits per entry of the INT8 square (gemm_i8.cu); 0 = 7 x 8-bit / 8 x 7-bit

  double* X = nullptr;   // [elems]
  double* X2 = nullptr;  // [elems]
  double* Q = nullptr;   // lazily allocated (blockDiagonalize)
  double* W = nullptr;
  double* T = nullptr;   // scratch N x N (Q', products)
  double* Xi = nullptr;  // imaginary planes (complex path only, lazily allocated)
  double* X2i = nullptr;
  double* Qi = nullptr;
  double* Wi = nullptr;
  double* Ti = nullptr;
  double* Qhat_i = nullptr;
  bool q_complex = false;
  double* Qhat = nullptr;
  int64_t qhat_cols = 0;
  std::vector<int64_t> blk_sizes;
  bool have_Q = false;

  // fill state: lut[provisional id] for the last sdpsr_fill
  double* lut = nullptr;
  size_t lut_alloc = 0;
  double* d_values = nullptr;   // staging for host coefficient vectors
  size_t values_alloc = 0;
  bool x_is_fill = false;       // X == fill(S, lut) and S unchanged since
  bool x_valid = false;
  // layout of the keys of tab[cur] (set by the last refine pass): value-coded keys can be decoded back
  // into the rounded class values (sdpsr_decode_lut)
  bool key_decodable = false;
  bool key_fast = false;
  int key_lbits = 0;
  // Is the label matrix transpose-invariant?  1 yes, 0 no, -1 unknown (checked lazily, 4 B/entry).  A refine
  // by a matrix that is symmetric by construction keeps a symmetric partition symmetric.
  int sym_state = -1;

  // ranking scratch
  uint32_t* rk_mi = nullptr;    // dense minidx list
  size_t rk_alloc = 0;
  uint32_t* bitmap = nullptr;   // [elems/32] first-occurrence bitmap
  uint32_t* bm_block = nullptr; // per-block popcount prefix
  size_t bm_blocks = 0;

  uint32_t* d_scalars = nullptr;   // small device scratch (64 words)
  void* h_pinned = nullptr;        // small pinned host scratch (4 KB)
  // copy stream (sdpsr_stage_objective, sdpsr_partition_get_labels_async): host <-> device transfers that overlap
  // the kernels of ctx->stream
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_gate = nullptr;       // recorded on ctx->stream: the copy may start
  cudaEvent_t staged_ev = nullptr;       // recorded on the copy stream: C has arrived in X
  cudaEvent_t labels_ev = nullptr;       // recorded on the copy stream: the labels have arrived on the host
  const double* staged_src = nullptr;    // host matrix whose upload into X is in flight / done
  uint64_t staged_seq = 0;               // api_seq at which the staging is still valid
  uint64_t api_seq = 0;                  // C-ABI calls entered so far
  bool labels_pending = false;

  ConstraintSet cons;

  // cusolver state (opaque here)
  void* solver = nullptr;
  void* solver_work = nullptr;
  size_t solver_work_bytes = 0;
  int* solver_info = nullptr;
  void* solver_hwork = nullptr;
  size_t solver_hwork_bytes = 0;

  // Krylov block-diagonalisation state (opaque here, krylov.cu)
  void* krylov = nullptr;

  // comm (multi-GPU)
  void* nccl = nullptr;
  void* local_group = nullptr;  // in-process transport (comm.cu): G contexts of one process, one thread each
  int nranks = 1, rank = 0;
  void* d_tiles = nullptr;      // this rank's (tm, tn) tile list for sharded GEMMs
  size_t tile_alloc = 0;
  // peer-mapped output buffers (CUDA IPC): peer_ptr[b][r] = rank r's copy of shardable buffer b
  // (0: X2, 1: T, 2: W); d_peer is the same table on the device.  The sharded GEMM stores its tiles
  // into every rank's buffer from its epilogue, so the exchange overlaps the math.
  static constexpr int MAX_RANKS = 16;
  double* peer_ptr[3][MAX_RANKS] = {};
  double** d_peer = nullptr;
  bool peer_ok = false;
  int* d_barrier = nullptr;

  // timing
  std::vector<EventPair> ev_pending;
  std::vector<cudaEvent_t> ev_pool;
  double t_ms[SDPSR_K_COUNT] = {0};
  int64_t t_launch[SDPSR_K_COUNT] = {0};
  double t_work[SDPSR_K_COUNT] = {0};
  int64_t launches = 0;

  // grow-only scratch buffers (slot -> device allocation): no cudaMalloc/cudaFree on the steady-state
  // path (allocation calls serialise across processes and jitter by up to seconds with IPC mappings)
  static constexpr int SCRATCH_SLOTS = 48;
  void* scratch_ptr[SCRATCH_SLOTS] = {};
  size_t scratch_bytes[SCRATCH_SLOTS] = {};

  std::string err;

  int fail(int code, const std::string& msg) {
    err = msg;
    return code;
  }
};

// Scope timer: records a CUDA-event pair around the kernels launched in its scope (SDPSR_F_TIMING).
struct Timed {
  sdpsr_ctx* c;
  int fam;
  double work;
  cudaEvent_t a = nullptr, b = nullptr;
  Timed(sdpsr_ctx* ctx, int family, double work_);
  ~Timed();
};

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline uint32_t next_pow2(uint64_t x) {
  uint64_t p = 1;
  while (p < x) p <<= 1;
  return (uint32_t)(p > 0x80000000ull ? 0x80000000ull : p);
}
static inline int bits_for(uint64_t maxval) {
  int b = 1;
  while ((maxval >> b) != 0) ++b;
  return b;
}

// --------------------------------------------------------------------------
// cross-TU entry points (host functions)
// --------------------------------------------------------------------------
// refine.cu
enum KeyMode { KM_ROUND = 0, KM_RAW = 1, KM_PAIR = 2 };
struct RefineSpec {
  KeyMode mode = KM_ROUND;
  const double* vals = nullptr;     // value source array (padded layout) or null
  double* vals_out = nullptr;       // optional: write the rounded values here
  const uint32_t* lab2 = nullptr;   // KM_PAIR: second provisional-id array
  bool fillproj = false;            // value = lut[label] - tpat[pid] instead of vals[idx]
  const double* lut = nullptr;
  const double* tpat = nullptr;
  const uint32_t* pid = nullptr;
  bool do_round = true;
  bool raw_bits = false;            // KM_RAW: key = the 64 bits of vals[idx] verbatim (signatures, not floats)
  double atol = 0;
  bool ignore_labels = false;       // treat the current labels as all-zero (fresh partition)
  bool keeps_symmetry = false;      // the refining values are symmetric by construction
  uint32_t* out_override = nullptr; // write provisional ids here instead of labels_alt (no swap)
  bool shard_block = false;         // sharded run: a pass with overrides also runs on this rank's block + merge
  KeyTable* table_override = nullptr;
};
int sdpsr_refine_pass(sdpsr_ctx* ctx, const RefineSpec& spec, int64_t* dim);
int sdpsr_generic_refine_values(sdpsr_ctx* ctx, const double* dvals, double atol, bool do_round,
                                double* vals_out, int64_t* dim, bool keeps_symmetry = false);
int sdpsr_table_alloc(sdpsr_ctx* ctx, KeyTable& t, size_t cap);
void sdpsr_table_free(KeyTable& t);
int sdpsr_rank_table(sdpsr_ctx* ctx, KeyTable& t);
int sdpsr_build_lut(sdpsr_ctx* ctx, const double* d_values, int64_t len);
int sdpsr_materialize_fill(sdpsr_ctx* ctx, double* dst);
int sdpsr_decode_lut(sdpsr_ctx* ctx, double atol);
int sdpsr_lut_stats(sdpsr_ctx* ctx, double* vmax, double* vmin_nz);
int sdpsr_canonical_labels(sdpsr_ctx* ctx, uint32_t* dst_unpadded);
int sdpsr_round_params(sdpsr_ctx* ctx, double atol, double* scale, long long* iscale, int* qbits);
int sdpsr_ensure_tmp_labels(sdpsr_ctx* ctx);
int sdpsr_label_bits(const sdpsr_ctx* ctx);
int sdpsr_upload_values(sdpsr_ctx* ctx, const double* values, int64_t len);

// gemm_f64.cu :  C[M x Nc] = A[M x K] * B[K x Nc], all column-major with the given lds
int sdpsr_gemm_f64(sdpsr_ctx* ctx, const double* A, int64_t lda, const double* B, int64_t ldb,
                   double* C, int64_t ldc, int64_t M, int64_t Nc, int64_t K, bool symmetric_out,
                   bool shard = false, int accum = 0);
int sdpsr_mirror_lower(sdpsr_ctx* ctx, double* C, int64_t ldc, int64_t n, int64_t col_begin = 0, int64_t col_end = -1);

// gemm_i8.cu : C = X * X for bit-for-bit symmetric X on the tcgen05 INT8 tensor path
int sdpsr_square_i8(sdpsr_ctx* ctx, const double* X, double* C, int slices, int bits, bool shard, bool force_range,
                    int* done);

// project.cu
int sdpsr_constraints_finalize(sdpsr_ctx* ctx);
void sdpsr_constraints_free(sdpsr_ctx* ctx);
int sdpsr_rowdots(sdpsr_ctx* ctx, const double* x_array, const double* lut, std::vector<double>& out,
                  bool x_own_block = false);
int sdpsr_solve_gram(sdpsr_ctx* ctx, std::vector<double>& rhs);
int sdpsr_upload_tpat(sdpsr_ctx* ctx, const std::vector<double>& coef);
int sdpsr_symmetric_check(sdpsr_ctx* ctx, int* is_sym);
int sdpsr_matrix_symmetric(sdpsr_ctx* ctx, const double* x, int* is_sym);

// blockdiag.cu
void sdpsr_blockdiag_free(sdpsr_ctx* ctx);
void sdpsr_blockdiag_rebind(sdpsr_ctx* ctx);

// krylov.cu
void sdpsr_krylov_free(sdpsr_ctx* ctx);
bool sdpsr_module_basis_available(const sdpsr_ctx* ctx);
int sdpsr_module_basis_image(sdpsr_ctx* ctx, double atol, double* out, int64_t out_len);
void sdpsr_module_invalidate_qhat(sdpsr_ctx* ctx);

// comm.cu
void sdpsr_comm_free(sdpsr_ctx* ctx);
int sdpsr_comm_exchange_tilecols(sdpsr_ctx* ctx, double* C, int64_t ldc, int64_t ncols, int tile_cols, int ntilecols,
                                 bool snake = false);
// Owner of tile-column tn of a lower-triangle product.  Column tn holds (tiles_m - c*tn) tiles, so a plain round-robin
// gives rank 0 the G heaviest columns of every round (576 vs 464 tiles at N = 16384 on 8 ranks); in snake order
// (0..G-1, G-1..0, ...) the columns of two consecutive rounds pair up to the same weight on every rank.
static inline int sdpsr_tilecol_owner_snake(int tn, int nranks) {
  const int q = tn / nranks, p = tn % nranks;
  return (q & 1) ? nranks - 1 - p : p;
}
int sdpsr_comm_bcast(sdpsr_ctx* ctx, void* buf, size_t bytes, int root);
int sdpsr_comm_allgather(sdpsr_ctx* ctx, void* recv, size_t bytes_per_rank);
int sdpsr_comm_allgatherv(sdpsr_ctx* ctx, void* base, const size_t* offset, const size_t* bytes);

// shard.cu
void sdpsr_shard_block(const sdpsr_ctx* ctx, int r, uint64_t* begin, uint64_t* end);
bool sdpsr_shard_active(const sdpsr_ctx* ctx);
int sdpsr_shard_merge(sdpsr_ctx* ctx, KeyTable& tloc, uint32_t* lab, int64_t* dim);
int sdpsr_shard_gather_compact(sdpsr_ctx* ctx);
int sdpsr_shard_ensure_full_labels(sdpsr_ctx* ctx);
int sdpsr_comm_agree_min(sdpsr_ctx* ctx, int* flag);
int sdpsr_comm_barrier(sdpsr_ctx* ctx);
bool sdpsr_comm_shares_device(const sdpsr_ctx* ctx);
double* const* sdpsr_comm_peer_table(sdpsr_ctx* ctx, const double* C);
int sdpsr_ensure_matrix(sdpsr_ctx* ctx, double** p);

// context.cu: grow-only scratch (slots 0-7 transient per call, 8-15 constraint set, 16+ persistent)
int sdpsr_scratch(sdpsr_ctx* ctx, int slot, size_t bytes, void** out);
template <typename T>
static inline int sdpsr_scratch_t(sdpsr_ctx* ctx, int slot, size_t count, T** out) {
  return sdpsr_scratch(ctx, slot, count * sizeof(T), reinterpret_cast<void**>(out));
}

// launch bookkeeping
static inline void count_launch(sdpsr_ctx* ctx, int k = 1) { ctx->launches += k; }
