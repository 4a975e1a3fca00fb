/*
 * sdpsr.h -- C ABI of libsdpsr_cuda.so: the B200-native engine behind
 * SDPSymmetryReduction.jl's Jordan-reduction hot path.
 *
 * The reference (DanielBrosch/SDPSymmetryReduction.jl v0.2.1) is pure Julia and
 * has no FFI; its only extension point is the AbstractPartition contract
 * (src/abstract_part.jl:1-18).  This header is therefore NEW surface: each entry
 * point names the reference code it replaces (file:line relative to the
 * reference root) so that a `CuPartition <: AbstractPartition` Julia back-end can
 * bind it with `ccall`.
 *
 * Conventions
 *  - Every matrix is COLUMN-MAJOR (Julia), order N, linear index idx = i + N*j
 *    (0-based); first-occurrence order == ascending idx (src/partitions.jl:24-35).
 *  - Pointers are plain host pointers unless stated otherwise.  Every bulk
 *    buffer argument may ALSO be a CUDA device pointer of the context's device
 *    (copies use cudaMemcpyDefault); this is how a caller keeps data resident.
 *  - Every function returns 0 (SDPSR_OK) or a negative sdpsr_status; the message
 *    of the last failure is available from sdpsr_last_error().
 *  - Calls are blocking: they return after the context's stream has drained.
 *    A context is not thread-safe; distinct contexts are independent.
 *  - No torch / C++ types cross this boundary.
 *  - There is no CPU fallback: without a CUDA device sdpsr_create() fails.
 */
#ifndef SDPSR_H
#define SDPSR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDPSR_VERSION 100 /* 0.1.0 */

typedef struct sdpsr_ctx sdpsr_ctx;

typedef enum sdpsr_status {
  SDPSR_OK = 0,
  SDPSR_E_INVALID = -1,        /* bad argument (AssertionError in the reference)        */
  SDPSR_E_CUDA = -2,           /* CUDA runtime / driver failure                         */
  SDPSR_E_NO_DEVICE = -3,      /* no CUDA device: there is no CPU fallback              */
  SDPSR_E_ALLOC = -4,          /* out of device / host memory                           */
  SDPSR_E_LABEL_OVERFLOW = -5, /* label does not fit the requested integer width
                                  (InexactError, src/partitions.jl:29,63)              */
  SDPSR_E_NOT_SYMMETRIC = -6,  /* partition is not transpose-invariant: real eigen path
                                  impossible (InvalidDecompositionField,
                                  src/eigen_decomposition.jl:247-253)                  */
  SDPSR_E_CUSOLVER = -7,       /* cuSOLVER syevd failed / did not converge              */
  SDPSR_E_STATE = -8,          /* call sequence error (e.g. square before fill)         */
  SDPSR_E_SINGULAR = -9,       /* the constraint matrix has rank 0 (A == 0)             */
  SDPSR_E_NCCL = -10,          /* NCCL failure / NCCL library not loadable              */
  SDPSR_E_UNSUPPORTED = -11,   /* valid request outside this build's limits             */
  SDPSR_E_KRYLOV = -12         /* the Krylov block-diagonalisation path is not applicable
                                  (no clean Lanczos breakdown): run the dense path      */
} sdpsr_status;

/* flags for sdpsr_create */
#define SDPSR_F_DEFAULT 0u
#define SDPSR_F_FORCE_BITMAP_RANK 1u /* test hook: always use the bitmap-scan ranking path   */
#define SDPSR_F_TINY_TABLE 2u        /* test hook: start with a 64-slot table (forces growth) */
#define SDPSR_F_NO_SMEM_CACHE 4u     /* test hook: bypass the per-CTA key cache              */
#define SDPSR_F_TIMING 8u            /* record CUDA-event timings per kernel family          */
#define SDPSR_F_NO_SYRK 16u          /* square with the full GEMM even for symmetric X       */
#define SDPSR_F_NCCL_EXCHANGE 32u    /* multi-GPU: exchange GEMM tiles with NCCL broadcasts instead
                                        of peer stores from the GEMM epilogue (A/B switch)    */
#define SDPSR_F_NO_I8 64u            /* never square on the INT8 tensor path (always DMMA)    */
#define SDPSR_F_FORCE_I8 128u        /* square every symmetric X on the INT8 tensor path, whatever
                                        N (default: only where it is faster, N >= 2048)       */
#define SDPSR_F_REPLICATED_REFINE 256u /* multi-rank: every rank runs the streaming passes on the full
                                        matrix (round-1 behaviour) instead of sharding the partition by
                                        column blocks with a key-table merge (A/B switch)      */

/* which device-resident matrix sdpsr_get_matrix / sdpsr_set_matrix address */
#define SDPSR_MAT_X 0   /* current element X (src/partitions.jl:121)      */
#define SDPSR_MAT_X2 1  /* last product X^2 / XY (src/partitions.jl:122)  */
#define SDPSR_MAT_Q 2   /* eigenvectors of the last sdpsr_eig             */
#define SDPSR_MAT_W 3   /* Q' A Q of the last sdpsr_block_norms           */

/* kernel families for sdpsr_timing_get */
#define SDPSR_K_REFINE 0   /* round + key + first-occurrence relabel pass        */
#define SDPSR_K_GEMM 1     /* FP64 DMMA GEMM                                      */
#define SDPSR_K_FILL 2     /* label -> value gather                               */
#define SDPSR_K_PROJECT 3  /* constraint row dots                                 */
#define SDPSR_K_RANK 4     /* canonical first-occurrence ranking of a key table   */
#define SDPSR_K_EIG 5      /* cuSOLVER syevd (library)                            */
#define SDPSR_K_BASIS 6    /* basis_image reduction                               */
#define SDPSR_K_MISC 7     /* everything else (transpose, norms, ...)             */
#define SDPSR_K_KRYLOV 8   /* label-matrix x vector products of the Krylov path   */
#define SDPSR_K_GEMM_I8 9 /* symmetric square on the tcgen05 INT8 tensor path      */
#define SDPSR_K_COUNT 10

/* ------------------------------------------------------------------ lifetime */
int sdpsr_version(void);
/* ctx may be NULL: returns the message of the last failed sdpsr_create on this thread. */
const char* sdpsr_last_error(const sdpsr_ctx* ctx);
/* Allocates the device state for N x N problems on CUDA device `device`:
 * labels (u32, double-buffered), X, X2 (f64), key tables, scratch.
 * Replaces the allocations of src/partitions.jl:117-122.                               */
int sdpsr_create(sdpsr_ctx** out, int64_t n, int device, uint32_t flags);
int sdpsr_destroy(sdpsr_ctx* ctx);
int sdpsr_device_count(int* count);

/* ------------------------------------------------------------- constraints A
 * The `A` argument of admissible_subspace (src/partitions.jl:112).  The engine
 * derives from it the per-entry constraint pattern ids, the pattern table and a rank-revealing
 * (pivoted Cholesky, extended precision) factorisation of the Gram matrix A A' that replace
 * `qr(A')` (src/partitions.jl:124) in project_colspace! (src/utils.jl:62-66).
 *   dense: A is m x N^2 column-major (a Julia Matrix{Float64}), ld = m.
 *   csr  : row k holds entries rowptr[k]..rowptr[k+1]-1; this is the CSC storage of
 *          A' (N^2 x m), i.e. `SparseMatrixCSC(transpose(A))` in Julia.  The arrays are uploaded as they
 *          are and validated / re-indexed by one kernel (no host pass over the stored entries).
 *   csc  : the SparseMatrixCSC storage of A itself (colptr has N^2+1 entries).
 * index_base is 0 (C / numpy) or 1 (Julia).                                           */
int sdpsr_set_constraints_dense(sdpsr_ctx* ctx, int64_t m, const double* A);
int sdpsr_set_constraints_csr(sdpsr_ctx* ctx, int64_t m, const int64_t* rowptr,
                              const int64_t* colidx, const double* vals, int index_base);
/* the same with 32-bit column indices (scipy.sparse's default index type): no widening copy on the host */
int sdpsr_set_constraints_csr_i32(sdpsr_ctx* ctx, int64_t m, const int64_t* rowptr,
                                  const int32_t* colidx, const double* vals, int index_base);
int sdpsr_set_constraints_csc(sdpsr_ctx* ctx, int64_t m, const int64_t* colptr,
                              const int64_t* rowval, const double* nzval, int index_base);
/* number of distinct non-empty constraint column patterns found */
int sdpsr_constraint_patterns(sdpsr_ctx* ctx, int64_t* npatterns);
/* numerical rank of A: rows that depend on the others (negligible pivot of the pivoted Cholesky of
 * A A') are dropped from the projector, as the rank-revealing sparse `qr(A')` of the reference does
 * (src/partitions.jl:124); the projection onto the row space is unchanged by that.              */
int sdpsr_constraint_rank(sdpsr_ctx* ctx, int64_t* rank);


/* ------------------------------------------------------------- Partition state
 * The context holds one Partition S (src/partitions.jl:6-9): labels 0..dim.         */
/* S = empty partition (all labels 0, dim 0).                                         */
int sdpsr_partition_reset(sdpsr_ctx* ctx);
/* Partition{T}(M::AbstractMatrix{<:Integer}) = copy + __sort_unique!
 * (src/partitions.jl:37-60).  labels: N x N integers of elt_bytes in {1,2,4,8}.     */
int sdpsr_partition_set_labels(sdpsr_ctx* ctx, const void* labels, int elt_bytes,
                               int64_t* dim);
/* P.matrix in canonical first-occurrence numbering.  SDPSR_E_LABEL_OVERFLOW when dim(P) does not
 * fit elt_bytes (InexactError in the reference; the reference may throw earlier, for an intermediate
 * p1 + p2*(dim+1) of refine! that the engine never forms, src/partitions.jl:62-66).   */
int sdpsr_partition_get_labels(sdpsr_ctx* ctx, void* labels, int elt_bytes);
/* sdpsr_partition_get_labels without the wait (the reference hands P.matrix to the caller, src/partitions.jl:84,188):
 * the canonical labels are produced into a staging buffer of their own and travel to `labels` on the context's copy
 * stream while later calls (blockDiagonalize) compute; `labels` should be pinned host memory and must stay valid until
 * sdpsr_partition_labels_wait returns (SDPSR_E_LABEL_OVERFLOW there if a label did not fit elt_bytes).            */
int sdpsr_partition_get_labels_async(sdpsr_ctx* ctx, void* labels, int elt_bytes);
int sdpsr_partition_labels_wait(sdpsr_ctx* ctx);
/* dim(P) (src/partitions.jl:13) */
int sdpsr_partition_dim(sdpsr_ctx* ctx, int64_t* dim);
/* number of entries with label 0 */
int sdpsr_partition_zero_count(sdpsr_ctx* ctx, int64_t* count);
/* _constraints(P) (src/diagonalize.jl:42-50; overridden per back-end, test/partitions_set.jl:92): for every
 * class c = 1..dim the ascending column-major linear indices (plus index_base: 0 for C, 1 for Julia) of its
 * entries, as one CSR: ptr[dim + 1], idx[idx_len] with idx_len = N^2 - sdpsr_partition_zero_count().
 * The generic slow path of the AbstractPartition contract (the host does the counting sort).            */
int sdpsr_partition_constraints(sdpsr_ctx* ctx, int64_t* ptr, uint32_t* idx, int64_t idx_len, int index_base);
/* 1 iff labels[i,j] == labels[j,i] for all i,j */
int sdpsr_partition_is_symmetric(sdpsr_ctx* ctx, int* is_symmetric);
/* S = refine!(S, Partition(M)) with M a host/device N x N Float64 matrix
 * (src/partitions.jl:24-35,62-66).  do_round != 0 applies _clamp_round!(M; atol)
 * first (src/utils.jl:34-53); do_round == 0 keys on the bits of M as they are
 * (needed for Part(CL), whose entries are averaged after rounding, :131-133).
 * On an empty S this is S = Partition(M).                                            */
int sdpsr_refine_values(sdpsr_ctx* ctx, const double* M, double atol, int do_round,
                        int64_t* dim);
/* S = refine!(S, Partition(L)) for an integer label matrix L (src/partitions.jl:62-66). */
int sdpsr_refine_labels(sdpsr_ctx* ctx, const void* labels, int elt_bytes, int64_t* dim);

/* Optional: start the upload of a HOST objective C (vec(C) of admissible_subspace, src/partitions.jl:109-118) on the
 * context's copy stream so that it overlaps sdpsr_set_constraints_*: stage, set constraints, sdpsr_init_partition with
 * the same pointer.  Any other call in between drops the staging (init_partition then copies as usual).  No-op for
 * device-resident C; a sharded context stages its own column block.  Returns without waiting; C must stay valid
 * until init_partition.                                                                                          */
int sdpsr_stage_objective(sdpsr_ctx* ctx, const double* C);


/* ------------------------------------------------- admissible_subspace pieces */
/* The initial partition of src/partitions.jl:129-146, computed on the device:
 *   CL = symmetrize(round(C - proj(C)));  X0 = round(proj(symmetrize(x0)));
 *   S = refine!(Part(CL), Part(X0)),  x0 = min-norm solution of A x = b (Krylov.craig).
 * snap_decimals >= 0 rounds both vectors to that many decimals before the
 * reference's truncation (DESIGN.md "init safeguard"); < 0 disables it.              */
int sdpsr_init_partition(sdpsr_ctx* ctx, const double* C, const double* b, double atol,
                         int snap_decimals, int64_t* dim);
/* X = fill!(X, S; values) (src/partitions.jl:68-75); len must equal dim(S).          */
int sdpsr_fill(sdpsr_ctx* ctx, const double* values, int64_t len);
/* x .-= proj(x); _clamp_round!(x); S = refine!(S, Part(X))   (src/partitions.jl:160-164).
 * Requires constraints and a preceding sdpsr_fill.  X stays on the device.           */
int sdpsr_project_round_refine(sdpsr_ctx* ctx, double atol, int64_t* dim);
/* X2 = X*X; _clamp_round!(X2); S = refine!(S, Part(X2))       (src/partitions.jl:172-174). */
int sdpsr_square_round_refine(sdpsr_ctx* ctx, double atol, int64_t* dim);
/* desymmetrize step (src/partitions.jl:210-214):
 * XY = fill(S,rx) * fill(S,ry); round; S = refine!(S, Part(XY)).                      */
int sdpsr_product_round_refine(sdpsr_ctx* ctx, const double* rx, const double* ry,
                               int64_t len, double atol, int64_t* dim);

/* --------------------------------------------------- blockDiagonalize pieces */
/* A = fill(S, r1); (vals, Q) = eigen(A)  (src/eigen_decomposition.jl:242-254).
 * cuSOLVER Dsyevd; vals ascending (N doubles, host); Q stays on the device.
 * SDPSR_E_NOT_SYMMETRIC when S is not transpose-invariant.                            */
int sdpsr_eig(sdpsr_ctx* ctx, const double* r1, int64_t len, double* vals);
/* A = fill(S, r2); W = Q' A Q; norms[i,j] = max|W[E_i,E_j]| for eigenspaces of equal
 * dimension, else 0 (src/eigen_decomposition.jl:177-204).  ptrs: nptr = ne+1 0-based
 * cluster boundaries; norms: ne x ne column-major (host).                             */
int sdpsr_block_norms(sdpsr_ctx* ctx, const double* r2, int64_t len, const int64_t* ptrs,
                      int64_t nptr, double* norms);
/* irreducible_decomposition + clamptol! (src/eigen_decomposition.jl:295-348,
 * src/diagonalize.jl:39).  kroot[e] = root eigenspace (0-based, smallest member) of
 * the isomorphism class of eigenspace e.  Leaves Qhat (N x sum(blk_sizes)) on the
 * device; blk_sizes must have room for ne entries; *nblk receives the block count.    */
int sdpsr_irreducible(sdpsr_ctx* ctx, const double* r3, int64_t len, const int64_t* ptrs,
                      int64_t nptr, const int64_t* kroot, double atol, int64_t* blk_sizes,
                      int64_t* nblk);
/* Copies Qhat (N x S, column-major) to the host. */
int sdpsr_get_qhat(sdpsr_ctx* ctx, double* qhat, int64_t len);
/* Replaces Qhat (test hook and complex-path helper): N x S column-major. */
int sdpsr_set_qhat(sdpsr_ctx* ctx, const double* qhat, const int64_t* blk_sizes, int64_t nblk);
/* basis_image (src/diagonalize.jl:64-89): out is packed as
 *   for i in 0..dim-1, for k in 0..nblk-1: Qhat_k' * 1[S==i+1] * Qhat_k  (s_k x s_k, col-major)
 * with |entries| < atol clamped to 0.  out_len must be dim * sum(s_k^2).              */
int sdpsr_basis_image(sdpsr_ctx* ctx, double atol, double* out, int64_t out_len);

/* ------------------------------------------ blockDiagonalize, module variant
 * Same results as sdpsr_eig / sdpsr_block_norms / sdpsr_irreducible (i.e. as
 * src/eigen_decomposition.jl:236-348) without the O(N^3) `eigen`: everything the reference takes from the
 * eigendecomposition of the generic element A1 = fill(S, r1) -- one unit vector per eigenspace (first
 * column of the root eigenspace, :311-314) and the projections of A3 times it (:327-336) -- lies in the
 * module M generated by one unit vector per diagonal class, an invariant subspace of dimension
 * D <= 2 dim(S).  The engine builds an orthonormal basis Q of M from columns of the label matrix (closing it
 * under A1 when the partition is a Jordan configuration that is not coherent), and runs the reference's
 * steps on the D x D matrices Q' A_t Q; Qhat = Q Z.  Cost: a few (N x N) x (N x D) products.
 * Self-validating: every entry point returns SDPSR_E_KRYLOV when the module is larger than max_dim, a rank
 * decision is ambiguous, the eigenspace dimensions are not integers adding up to N, or M is not invariant
 * under A2 / A3; the caller then runs the dense entry points with the same r1, r2, r3.
 *
 * sdpsr_eig_krylov: vals[0..ne) = distinct eigenvalues of A1 ascending (clusters of the reference's rule:
 *   a new eigenspace where the gap exceeds atol, :19-40), mult[i] = dim E_i (sum = N); vals and mult need
 *   room for max_dim entries (the cap on D; at most 4096).
 * sdpsr_block_norms_krylov: norms (ne x ne, column-major) = max |U_i' (Q' A2 Q) U_j| for eigenspaces of
 *   equal dimension, else 0 -- block_norms(Q'A2Q, Inf) (:177-204) evaluated inside M.
 * sdpsr_irreducible_krylov: as sdpsr_irreducible, with the eigenspaces of sdpsr_eig_krylov.          */
int sdpsr_eig_krylov(sdpsr_ctx* ctx, const double* r1, int64_t len, int64_t max_dim, double atol,
                     double* vals, int64_t* mult, int64_t* ne);
int sdpsr_block_norms_krylov(sdpsr_ctx* ctx, const double* r2, int64_t len, double* norms);
int sdpsr_irreducible_krylov(sdpsr_ctx* ctx, const double* r3, int64_t len, const int64_t* kroot,
                             double atol, int64_t* blk_sizes, int64_t* nblk);

/* ---------------------------------------------------------------- complex path
 * diagonalize(ComplexF64, P) (src/diagonalize.jl:25-40, src/compat.jl:46-68) for partitions that are
 * not transpose-invariant.  Complex vectors are interleaved (re, im) doubles; `len`/`out_len` count
 * complex numbers.  Call order and host-side steps are those of the real path; the general
 * eigensolver is cuSOLVER Xgeev, eigenvalues sorted by (re, im) like Julia's `eigen`.             */
int sdpsr_eig_complex(sdpsr_ctx* ctx, const double* r1, int64_t len, double* vals /* 2N */);
int sdpsr_block_norms_complex(sdpsr_ctx* ctx, const double* r2, int64_t len, const int64_t* ptrs,
                              int64_t nptr, double* norms);
int sdpsr_irreducible_complex(sdpsr_ctx* ctx, const double* r3, int64_t len, const int64_t* ptrs,
                              int64_t nptr, const int64_t* kroot, double atol, int64_t* blk_sizes,
                              int64_t* nblk);
int sdpsr_get_qhat_complex(sdpsr_ctx* ctx, double* qhat, int64_t len);
int sdpsr_basis_image_complex(sdpsr_ctx* ctx, double atol, double* out, int64_t out_len);

/* ------------------------------------------------------ reduced SDP assembly
 * The step right after the path (README.md:57-60, test/sd_problems.jl:32-37):
 *   newA = A * PMat  (m x dim, column-major),  newC = C' * PMat  (dim),
 * PMat[:, i] = vec(S .== i+1).  Either output may be NULL.  Sums are accumulated with
 * floating-point atomics (order not fixed; exact for the 0/1 data of the reference's problems). */
int sdpsr_reduce_problem(sdpsr_ctx* ctx, const double* C, double* newA, double* newC);

/* -------------------------------------------------------------- plumbing */
int sdpsr_get_matrix(sdpsr_ctx* ctx, int which, double* out);       /* N x N doubles */
int sdpsr_set_matrix(sdpsr_ctx* ctx, int which, const double* in);  /* N x N doubles */
/* C = A * B on the device with the engine's FP64 DMMA GEMM (test / bench hook);
 * a, b, c are SDPSR_MAT_* ids, c != a, c != b. */
int sdpsr_gemm(sdpsr_ctx* ctx, int a, int b, int c);
/* X2 = X * X on the device matrices, the product of `mul!(X2, X, X)` (src/partitions.jl:172) alone
 * (test / bench hook).  method 0: FP64 DMMA GEMM (half GEMM when X is symmetric); method 1: INT8
 * tensor path with `slices` int8 digits per entry (0 = the context's setting), which requires a
 * bit-for-bit symmetric X (SDPSR_E_INVALID otherwise); the digits are 8 bits wide for N <= 16384
 * (up to 7 slices) and 7 bits wide above (up to 8 slices).  method 2 / 3: INT8 with 7-bit / 8-bit
 * digits forced.                                                                                 */
int sdpsr_square(sdpsr_ctx* ctx, int method, int slices);
/* Number of int8 digits per entry used by the INT8 square inside sdpsr_square_round_refine:
 * 0 (default) = 7 digits of 8 bits for N <= 16384, 8 digits of 7 bits above (54 / 55 magnitude bits,
 * the accuracy of an FP64 GEMM); 2..7 = fewer digits; 8 = eight 7-bit digits.                    */
int sdpsr_set_square_slices(sdpsr_ctx* ctx, int slices);
/* Accumulated CUDA-event time, launch count and algorithmic work (bytes for HBM-bound
 * families, flops for SDPSR_K_GEMM, int8 operations for SDPSR_K_GEMM_I8) per kernel family since the last reset.           */
int sdpsr_timing_reset(sdpsr_ctx* ctx);
int sdpsr_timing_get(sdpsr_ctx* ctx, int family, double* total_ms, int64_t* launches,
                     double* work);
/* Run all further work of this context on the caller's CUDA stream (a cudaStream_t passed as
 * void*; NULL = the legacy default stream).  Calls stay blocking.  Lets a host framework order
 * and time the engine's kernels with its own events.                                          */
int sdpsr_set_stream(sdpsr_ctx* ctx, void* cuda_stream);
/* total number of kernels launched by this context since creation */
int sdpsr_launch_count(sdpsr_ctx* ctx, int64_t* launches);

/* -------------------------------------------------------------- multi-GPU
 * One context per process / GPU.  The N x N matrices are sharded by column blocks;
 * every rank keeps the full label matrix.  NCCL is loaded with dlopen at first use.  */
int sdpsr_comm_unique_id(void* id128 /* 128 bytes out */);
int sdpsr_comm_init(sdpsr_ctx* ctx, int nranks, int rank, const void* id128);
int sdpsr_comm_info(sdpsr_ctx* ctx, int* nranks, int* rank);
/* In-process transport: `nranks` contexts of ONE process, each driven by its own host thread, on the
 * same device or on different devices of the process (the sharded path must run with G
 * ranks mapped onto one GPU).  Same sharding, same kernels and the same call sequence as with NCCL; the
 * collectives are pointer exchanges + cudaMemcpy and the barrier is a host barrier, so no kernel waits
 * for another rank's kernel.  Create one group, then every rank's thread calls sdpsr_comm_init_local
 * (collective).  The group is freed when its last context is destroyed (sdpsr_destroy is collective).   */
int sdpsr_comm_local_group(void** group, int nranks);
int sdpsr_comm_init_local(sdpsr_ctx* ctx, void* group, int rank);

/* Test hook (host-only, needs no device): the tile deal of the sharded products exactly as the kernels use it.
 * kind 0: FP64 GEMM (all 128 x 128 tiles), 1: FP64 GEMM lower triangle, 2: INT8 square (128 x 256 tiles),
 * 3: INT8 square on CTA pairs (256 x 256).  out (may be NULL) receives *count (tm, tn) pairs.                  */
int sdpsr_debug_tile_deal(int kind, int64_t n, int nranks, int rank, int32_t* out, int64_t cap, int64_t* count);

/* Test hook (host-only): the work list of the single-CTA INT8 square kernel for `rank` of `nranks` on `grid` CTAs
 * (CTA b walks items b, b + grid, ...): whole tiles first, then the tiles of the partly filled last wave cut along K
 * into equal runs, one per CTA (csrc/gemm_i8.cu, build_schedule).  out (may be NULL) receives 8 int32 per item: tm, tn,
 * kb0, kb1, slot, part, nparts, sem; info (may be NULL) = {nmain, nslots, nsems, k-blocks per tile}.           */
int sdpsr_debug_i8_schedule(int64_t n, int nranks, int rank, int grid, int32_t* out, int64_t cap, int64_t* count,
                            int32_t* info);

#ifdef __cplusplus
}
#endif
#endif /* SDPSR_H */
