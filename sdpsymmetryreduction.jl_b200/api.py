"""Host-side mirror of the reference's Julia API for the Jordan-reduction hot path.

Same names, argument meaning and error behaviour as the reference:

  admissible_subspace(C, A, b; verbose, atol)      src/partitions.jl:77-190
  blockDiagonalize(P, verbose; epsilon, complex)   src/compat.jl:26-68
  diagonalize(T, P; verbose, atol)                 src/diagonalize.jl:25-40
  basis_image(Q, P; atol)                          src/diagonalize.jl:64-89
  desymmetrize(P; verbose, atol) / unSymmetrize    src/partitions.jl:197-223
  Partition(M), dim, refine, randomize             src/partitions.jl:6-75, src/abstract_part.jl:97-110

All array work runs in ``libsdpsr_cuda.so``.  What stays in the host language is
what the reference also keeps as scalar host logic: the driver loops, eigenvalue
clustering, the Otsu threshold and the union-find of eigenspaces
(src/eigen_decomposition.jl:19-40,83-139,163-219).

Random coefficients: the reference draws ``rand(dim)`` from the task-local RNG;
here every driver takes ``rand`` (a callable ``n -> ndarray``) and calls it at
exactly the reference's points with the reference's lengths, so
a caller can feed both sides the same numbers.
"""
from __future__ import annotations

import logging
import math
import os
import time
from collections import namedtuple
from typing import Callable, List, Optional

import numpy as np

from . import binding as B

log = logging.getLogger("sdpsr_b200")
RTOL_DEFAULT = math.sqrt(np.finfo(np.float64).eps)     # Base.rtoldefault(Float64)

BlockDiagonalization = namedtuple("BlockDiagonalization", ["blkSizes", "blks"])


class InvalidDecompositionField(Exception):
    """src/eigen_decomposition.jl:140-150"""


class NumericalInconsistency(Exception):
    """src/eigen_decomposition.jl:152-161"""


class DimensionMismatch(Exception):
    """thrown by check_block_sizes, src/diagonalize.jl:1-23"""


# ----------------------------------------------------------------------------
# context pool: device buffers (and the cuSOLVER workspace) are reused across calls of the same
# shape instead of being cudaMalloc'ed / cudaFree'd per call (like a caching allocator)
# ----------------------------------------------------------------------------
_CTX_POOL: dict = {}
_POOL_PER_KEY = 1


def _acquire_context(n: int, device: int = 0, flags: int = 0) -> B.Context:
    lst = _CTX_POOL.get((int(n), int(device), int(flags)))
    while lst:
        ctx = lst.pop()
        if getattr(ctx, "_h", None):
            ctx.reset()
            ctx._constraints_set = False
            return ctx
    return B.Context(n, device, flags)


def _release_context(ctx: Optional[B.Context]):
    if ctx is None or not getattr(ctx, "_h", None):
        return
    if ctx.comm_info()[0] > 1:           # contexts that carry a communicator belong to their creator
        return
    lst = _CTX_POOL.setdefault((ctx.n, ctx.device, ctx.flags), [])
    if len(lst) < _POOL_PER_KEY:
        lst.append(ctx)
    else:
        ctx.close()


def clear_context_pool():
    """Free the device memory held by pooled contexts."""
    for lst in _CTX_POOL.values():
        for ctx in lst:
            ctx.close()
    _CTX_POOL.clear()


def run_local_ranks(n: int, nranks: int, fn: Callable, *, devices=None, flags: int = 0):
    """Run ``fn(ctx, rank)`` on ``nranks`` ranks of ONE process, one host thread per rank, joined by the
    in-process communicator (``sdpsr_comm_init_local``).  ``devices[r]`` is the CUDA device of rank r
    (default: all on device 0 -- the sharded path with G ranks mapped onto one GPU, SURVEY.md section 4;
    with distinct devices it is the single-process multi-GPU mode).  Returns the list of results; the
    root-cause exception is re-raised after every thread has finished."""
    import threading
    devices = list(devices) if devices is not None else [0] * nranks
    group = B.Context.local_group(nranks)
    results, errors = [None] * nranks, [None] * nranks

    def worker(r):
        ctx = None
        try:
            ctx = B.Context(n, devices[r], flags)
            ctx.comm_init_local(group, r)
            results[r] = fn(ctx, r)
        except BaseException as e:          # noqa: BLE001 -- re-raised below
            errors[r] = e
        finally:
            if ctx is not None:
                ctx.close()                  # collective: every rank detaches

    threads = [threading.Thread(target=worker, args=(r,), name=f"sdpsr-rank{r}") for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    failed = [(r, e) for r, e in enumerate(errors) if e is not None]
    if failed:
        # a rank that fails breaks the communicator and the others then report "a rank did not reach the barrier":
        # re-raise the root cause (the first error that is not that follow-up), naming every failed rank
        def follow_up(e):
            return isinstance(e, B.SdpsrError) and "did not reach the barrier" in str(e)
        r, e = next(((r, e) for r, e in failed if not follow_up(e)), failed[0])
        if len(failed) > 1 and hasattr(e, "add_note"):
            e.add_note("ranks that failed: " + "; ".join(f"rank {q}: {type(x).__name__}: {x}" for q, x in failed))
        raise e
    return results


# SDPSR_COPY_STREAM=0: host transfers on the compute stream, as before the copy stream existed (A/B switch)
_COPY_STREAM = os.environ.get("SDPSR_COPY_STREAM", "1") != "0"


def _default_rand():
    rng = np.random.default_rng()
    return lambda n: rng.random(int(n))


# ----------------------------------------------------------------------------
# Partition
# ----------------------------------------------------------------------------
class Partition:
    """``Partition{T}``: labels 0..nparts in first-occurrence (column-major) order
    (src/partitions.jl:6-42).  ``Partition(M)`` builds it from a matrix of numbers on
    the GPU; ``Partition(nparts, matrix)`` wraps given canonical labels."""

    def __init__(self, *args, device: int = 0, _ctx: Optional[B.Context] = None):
        self._ctx = _ctx
        self.device = device
        self._matrix = None
        self._arrival = None          # callable that waits for an export still in flight (get_labels_async)
        if len(args) == 2:
            self.nparts = int(args[0])
            self.matrix = None if args[1] is None else np.asarray(args[1])   # None: device-resident only
        elif len(args) == 1:
            M = np.asarray(args[0])
            n = M.shape[0]
            assert M.shape == (n, n)
            with B.Context(n, device) as ctx:
                if np.issubdtype(M.dtype, np.integer) or M.dtype == bool:
                    self.nparts = ctx.set_labels(M.astype(np.int64))          # :37-42
                else:
                    self.nparts = ctx.refine_values(M.astype(np.float64), RTOL_DEFAULT, do_round=False)  # :24-35
                self.matrix = ctx.get_labels(np.uint32)
        else:
            raise TypeError("Partition(M) or Partition(nparts, matrix)")

    @property
    def matrix(self):
        """P.matrix (src/partitions.jl:6-9).  When admissible_subspace started the export on the copy stream, the first
        access waits for it (and raises OverflowError if a label did not fit the requested type)."""
        if self._arrival is not None:
            wait, self._arrival = self._arrival, None
            wait()
        return self._matrix

    @matrix.setter
    def matrix(self, value):
        self._arrival = None
        self._matrix = value

    @property
    def shape(self):
        return self.matrix.shape

    def size(self, *a):
        return self.matrix.shape if not a else self.matrix.shape[a[0]]

    def __eq__(self, other):                       # src/partitions.jl:16-17
        return (isinstance(other, Partition) and self.nparts == other.nparts
                and np.array_equal(self.matrix, other.matrix))

    def __repr__(self):
        return f"Partition(nparts={self.nparts}, size={self.matrix.shape})"

    def _context(self, flags: int = 0) -> B.Context:
        """A context whose device partition equals this one."""
        ctx = self._ctx
        if ctx is not None and getattr(ctx, "_h", None):
            return ctx
        n = self.matrix.shape[0]
        ctx = _acquire_context(n, self.device, flags)
        d = ctx.set_labels(self.matrix)
        assert d == self.nparts, (d, self.nparts)
        self._ctx = ctx
        return ctx

    def release(self):
        """Give the device state kept alive for follow-up calls back to the context pool."""
        if self._arrival is not None:
            _ = self.matrix                       # an export still in flight reads a buffer of this context
        if self._ctx is not None:
            _release_context(self._ctx)
            self._ctx = None


def dim(P: Partition) -> int:
    return P.nparts


def refine(P1: Partition, P2: Partition) -> Partition:
    """``coarsestPart(P, Q) = refine!(deepcopy(P), Q)`` (src/compat.jl:12, src/partitions.jl:62-66)."""
    n = P1.matrix.shape[0]
    with B.Context(n, P1.device) as ctx:
        ctx.set_labels(P1.matrix)
        d = ctx.refine_labels(P2.matrix)
        return Partition(d, ctx.get_labels(np.uint32), device=P1.device)


def randomize(P: Partition, rand: Optional[Callable] = None) -> np.ndarray:
    """``randomize(Float64, P)`` (src/abstract_part.jl:97-110)."""
    rand = rand or _default_rand()
    ctx = P._context()
    ctx.fill(rand(P.nparts))
    return ctx.get_matrix(B.MAT_X)


# ----------------------------------------------------------------------------
# admissible_subspace
# ----------------------------------------------------------------------------
def admissible_subspace(C, A, b, *, verbose: bool = False, atol: float = RTOL_DEFAULT,
                        rand: Optional[Callable] = None, snap_decimals: Optional[int] = 12,
                        device: int = 0, flags: int = 0, label_dtype=np.uint32,
                        init_elements=None, trace: Optional[dict] = None,
                        keep_context: bool = True, ctx: Optional[B.Context] = None,
                        labels_out=None, fetch_labels: bool = True) -> Partition:
    """Optimal admissible partition subspace of  min <C,x>, A x = b, Mat(x) psd
    (src/partitions.jl:77-190).

    ``init_elements=(CL, X0)`` lets a host that owns the reference's own ``qr`` and
    ``Krylov.craig`` (the Julia wrapper) supply the two initial elements
    (:129-142); otherwise they are computed on the device.  ``label_dtype=np.uint16``
    returns the reference's default ``Partition{UInt16}`` and raises ``OverflowError`` (InexactError) when the
    FINAL dim does not fit; the reference can also throw mid-loop, when an intermediate
    ``p1 + p2*(dim(P1)+1)`` of ``refine!`` exceeds 65535 (:62-66) -- the engine has no such intermediate, so
    runs that error there for that reason alone succeed here.
    """
    rand = rand or _default_rand()
    Cv = C
    if hasattr(C, "todense"):
        Cv = np.asarray(C.todense()).reshape(-1)
    nn = int(Cv.numel()) if hasattr(Cv, "numel") else int(np.prod(np.shape(Cv)))
    n = math.isqrt(nn)
    if n * n != nn:
        raise AssertionError("n^2 == length(C)")                    # :118
    own = ctx is None
    if own:
        ctx = _acquire_context(n, device, flags)
    t0 = time.perf_counter()
    if init_elements is None:
        if isinstance(Cv, np.ndarray) and _COPY_STREAM:
            # one float64 buffer for both calls: the upload of a host C starts now, on the copy stream, and overlaps the
            # constraint set-up (a sharded context stages its own column block)
            Cv = np.ascontiguousarray(Cv, dtype=np.float64).reshape(-1)
            ctx.stage_objective(Cv)
        ctx.set_constraints(A)
        cur = ctx.init_partition(Cv, b, atol, snap_decimals)         # :124-146
    else:
        ctx.set_constraints(A)
        CL, X0 = init_elements
        ctx.reset()
        ctx.refine_values(CL, atol, do_round=False)                  # S = Part(CL)            :145
        cur = ctx.refine_values(X0, atol, do_round=False)            # refine!(S, Part(X0))    :146
    maxdim = (n * n + n) // 2
    if verbose:
        log.info("Starting the reduction. Dimensions: maximal=%d initial=%d", maxdim, cur)
    if trace is not None:
        trace.update({"init": cur, "iters": [], "t_init": time.perf_counter() - t0})
    it = 0
    while cur < maxdim:                                              # :154
        it += 1
        if verbose:
            log.debug("Iteration %d, Current dimension: %d", it, cur)
        ctx.fill(rand(cur))                                          # randomize!(X, S)        :159
        d_proj = ctx.project_round_refine(atol)                      # :160-164
        if d_proj != cur:                                            # :166-168
            ctx.fill(rand(d_proj))
        d_sq = ctx.square_round_refine(atol)                         # :172-174
        if trace is not None:
            trace["iters"].append((d_proj, d_sq))
        if cur == d_sq:                                              # :180-182
            break
        cur = d_sq
    if verbose:
        log.info("Minimal admissible subspace converged in %d iterations at dimension: final=%d", it, ctx.dim())
    if trace is not None:
        trace["iterations"] = it
        trace["t_total"] = time.perf_counter() - t0
    def overflow(e):
        if e.code == B.E_LABEL_OVERFLOW:
            return OverflowError("InexactError: label does not fit " + str(np.dtype(label_dtype)))
        return e

    arrival = None
    try:
        if not fetch_labels:          # caller keeps the partition on the device
            labels = None
        elif labels_out is not None and keep_context and _COPY_STREAM:
            # a caller-owned (pinned) buffer: the export runs on the copy stream while the caller goes on to
            # blockDiagonalize; P.matrix waits for it on first access
            labels = ctx.get_labels_async(label_dtype, labels_out)

            def arrival(ctx=ctx):
                try:
                    ctx.labels_wait()
                except B.SdpsrError as e:
                    raise overflow(e) from e
        else:
            labels = ctx.get_labels(label_dtype, out=labels_out)
    except B.SdpsrError as e:
        raise overflow(e) from e
    P = Partition(ctx.dim(), labels, device=device, _ctx=ctx if keep_context else None)
    P._arrival = arrival
    if own and not keep_context:
        _release_context(ctx)
    return P


# ----------------------------------------------------------------------------
# desymmetrize
# ----------------------------------------------------------------------------
def desymmetrize(P: Partition, *, verbose: bool = False, atol: float = RTOL_DEFAULT,
                 rand: Optional[Callable] = None) -> Partition:
    """WL-style closure under products X*Y (src/partitions.jl:197-223)."""
    rand = rand or _default_rand()
    n = P.matrix.shape[0]
    ctx = _acquire_context(n, P.device)          # deepcopy(P): never touch P's own context
    cur = ctx.set_labels(P.matrix)
    it = 0
    while True:
        it += 1
        rx = rand(cur)                                               # :210
        ry = rand(cur)                                               # :211
        d = ctx.product_round_refine(rx, ry, atol)                   # :212-214
        if d == cur:
            break
        cur = d
    if verbose:
        log.info("desymmetization converged in %d iterations", it)
    return Partition(ctx.dim(), ctx.get_labels(np.uint32), device=P.device, _ctx=ctx)


def unSymmetrize(P: Partition, **kw) -> Partition:     # src/compat.jl:70
    return desymmetrize(P, **kw)


# ----------------------------------------------------------------------------
# host-side scalar logic of Murota Alg. 4.1 (same split as the reference)
# ----------------------------------------------------------------------------
def eigen_clusters(values: np.ndarray, atol: float) -> np.ndarray:
    """``EigenDecomposition`` ctor (src/eigen_decomposition.jl:19-40): 0-based ptrs."""
    values = np.asarray(values)
    brk = np.flatnonzero(np.abs(values[1:] - values[:-1]) > atol) + 1
    gaps = np.abs(values[brk] - values[brk - 1])
    tiny = gaps < np.spacing(np.maximum(np.abs(values[brk]), np.abs(values[brk - 1])))
    if tiny.any():
        log.warning("Possibly numerically challenging example: no clear spectral gap")   # :34-36
    return np.concatenate([[0], brk, [values.size]]).astype(np.int64)


def otsu_threshold(X: np.ndarray, atol: float) -> float:
    """src/eigen_decomposition.jl:83-139 (16-bin log histogram + Otsu split)."""
    n_bins = max(int(math.ceil(-math.log10(np.finfo(np.float64).eps))), 4)
    a = np.abs(np.asarray(X, dtype=np.float64)).reshape(-1)
    lo, hi = a.min(), a.max()
    if lo < atol:
        lo = atol
    assert lo > 0
    edges = np.exp(np.linspace(math.log(lo), math.log(hi), n_bins + 1))
    if np.all(np.diff(edges) > 0):
        first_gt = np.searchsorted(edges, a, side="right")
    else:
        first_gt = np.array([next((i for i, e in enumerate(edges) if e > x), n_bins) for x in a])
    k = np.clip(first_gt, 1, n_bins)
    counts = np.bincount(k - 1, minlength=n_bins)[:n_bins]
    pdf = counts / counts.sum()
    w = np.cumsum(pdf)
    mu = np.cumsum(np.log(edges[:-1]) * pdf)
    with np.errstate(divide="ignore", invalid="ignore"):
        s2 = (mu[-1] * w - mu) ** 2 / (w * (1 - w))
    cand = s2[:-1]
    nan = np.flatnonzero(np.isnan(cand))
    kk = int(nan[0]) if nan.size else int(np.argmax(cand))
    return float(edges[kk + 1])


class _DisjointSets:
    """DataStructures.IntDisjointSets: union by rank, ties -> first argument's root."""

    def __init__(self, n):
        self.p = list(range(n))
        self.r = [0] * n

    def find(self, x):
        root = x
        while self.p[root] != root:
            root = self.p[root]
        while self.p[x] != root:
            self.p[x], x = root, self.p[x]
        return root

    def union(self, x, y):
        xr, yr = self.find(x), self.find(y)
        if xr == yr:
            return
        if self.r[xr] < self.r[yr]:
            xr, yr = yr, xr
        elif self.r[xr] == self.r[yr]:
            self.r[xr] += 1
        self.p[yr] = xr


def _isomorphism_classes(norms: np.ndarray, atol: float) -> np.ndarray:
    """src/eigen_decomposition.jl:205-219 + __isconsistent (:163-167). Returns kroot."""
    ne = norms.shape[0]
    thr = otsu_threshold(norms, atol)
    K = _DisjointSets(ne)
    ii, jj = np.nonzero(np.triu(norms >= thr, k=1))
    for i, j in zip(ii.tolist(), jj.tolist()):       # row-major order == the reference's loops
        K.union(i, j)
    kpart = [K.find(i) for i in range(ne)]
    first = {}
    for i, r in enumerate(kpart):
        first.setdefault(r, i)
    if not all(r == f for r, f in first.items()):
        raise NumericalInconsistency(
            "eigen_decomposition: the K-partition seems inconsistent with eigenspaces. "
            "Decrease `atol`, or simply try again.")
    return np.asarray(kpart, dtype=np.int64)


# ----------------------------------------------------------------------------
# diagonalize / basis_image / blockDiagonalize
# ----------------------------------------------------------------------------
# Largest dim(P) for which ``eig="auto"`` tries the module variant first (csrc/krylov.cu): its cost is a few
# N x N x D products with D <= 2 dim(P), against an O(N^3) syevd, and a failed attempt falls back to the dense
# path with the same coefficient vectors.
KRYLOV_AUTO_MAX_DIM = 1024
KRYLOV_MAX_MODULE_DIM = None        # cap on the module dimension (default 2 dim(P) + 16); tests lower it
EIG_MODES = ("auto", "syevd", "krylov")


class _Draws:
    """Records the coefficient vectors drawn from ``rand`` so that a second attempt (the dense path
    after an inapplicable Krylov attempt) consumes the SAME vectors and the caller's generator
    advances exactly as in the reference (three draws per ``diagonalize``)."""

    def __init__(self, rand):
        self.rand = rand
        self.vals = []
        self.pos = 0

    def __call__(self, n):
        if self.pos < len(self.vals):
            v = self.vals[self.pos]
            assert v.size == int(n)
        else:
            v = np.asarray(self.rand(n), dtype=np.float64)
            self.vals.append(v)
        self.pos += 1
        return v

    def rewind(self):
        self.pos = 0


def eigen_decomposition(P: Partition, *, atol: float, rand: Callable, ctx: Optional[B.Context] = None):
    """src/eigen_decomposition.jl:236-273. Returns (values, ptrs, kroot); Q stays on the device."""
    ctx = ctx or P._context()
    try:
        vals = ctx.eig(rand(P.nparts))                               # :242-254
    except B.SdpsrError as e:
        if e.code == B.E_NOT_SYMMETRIC:
            raise InvalidDecompositionField(
                "Decomposition over Float64 was requested but eigenvalues of type ComplexF64 were found. "
                "Consider calling `diagonalize` with ComplexF64 as its first argument.") from e
        raise
    ptrs = eigen_clusters(vals, atol)
    norms = ctx.block_norms(rand(P.nparts), ptrs)                    # :259, :203-204
    kroot = _isomorphism_classes(norms, atol)
    return vals, ptrs, kroot


class _KrylovNotApplicable(Exception):
    pass


def _diagonalize_krylov(P: Partition, ctx: B.Context, atol: float, rand: Callable):
    """The same three steps inside the module generated by one unit vector per diagonal class
    (csrc/krylov.cu).  Raises ``_KrylovNotApplicable`` when the device reports that it does not apply."""
    try:
        vals, mult = ctx.eig_krylov(rand(P.nparts), atol,                              # :242-254, clusters :19-40
                                    max_dim=KRYLOV_MAX_MODULE_DIM or 2 * P.nparts + 16)
        ptrs = np.concatenate([[0], np.cumsum(mult)]).astype(np.int64)
        norms = ctx.block_norms_krylov(rand(P.nparts), vals.size)    # :259, :203-204
        kroot = _isomorphism_classes(norms, atol)
        sizes = ctx.irreducible_krylov(rand(P.nparts), kroot, atol)  # :306, clamptol! :39
    except B.SdpsrError as e:
        if e.code == B.E_NOT_SYMMETRIC:
            raise InvalidDecompositionField(
                "Decomposition over Float64 was requested but eigenvalues of type ComplexF64 were found. "
                "Consider calling `diagonalize` with ComplexF64 as its first argument.") from e
        if e.code == B.E_KRYLOV:
            raise _KrylovNotApplicable(e.msg) from e
        raise
    return ptrs, kroot, sizes


def diagonalize(P: Partition, *, verbose: bool = False, atol: Optional[float] = None,
                rand: Optional[Callable] = None, complex: bool = False,
                fetch: bool = True, ctx: Optional[B.Context] = None, eig: str = "auto"):
    """``diagonalize(Float64, P)`` (src/diagonalize.jl:25-40): list of N x s_k matrices Q_hat.

    ``eig``: ``"syevd"`` = the reference's algorithm step by step (dense ``eigen`` through cuSOLVER);
    ``"krylov"`` = the matrix-free variant for partitions with few eigenspaces (same blocks, no dense
    eigendecomposition; raises if it is not applicable); ``"auto"`` (default) tries the Krylov variant
    when dim(P) <= KRYLOV_AUTO_MAX_DIM and falls back to ``syevd`` with the same coefficient vectors.
    ``P._eig_mode`` records the path taken."""
    rand = rand or _default_rand()
    if eig not in EIG_MODES:
        raise ValueError(f"eig must be one of {EIG_MODES}")
    if complex:
        return _diagonalize_complex(P, verbose=verbose, atol=atol, rand=rand, fetch=fetch)
    ctx = ctx or P._context()
    n = ctx.n
    if atol is None:
        atol = 1e-12 * n
    draws = _Draws(rand)
    sizes = None
    if eig == "krylov" or (eig == "auto" and P.nparts <= KRYLOV_AUTO_MAX_DIM):
        t = time.perf_counter()
        try:
            try:
                ptrs, kroot, sizes = _diagonalize_krylov(P, ctx, atol, draws)
            except NumericalInconsistency as e:              # its norms are not the reference's: let the
                if eig == "krylov":                           # reference's own statistic decide
                    raise
                raise _KrylovNotApplicable(str(e)) from e
            if eig == "auto" and sum(int(x) * (int(x) + 1) // 2 for x in sizes) != P.nparts:
                raise _KrylovNotApplicable("block sizes do not add up to dim(P)")
            P._eig_mode = "krylov"
            if verbose:
                log.info("Eigenspaces and algebra-isomorphism inside the cyclic module (%d eigenspaces)... %.3fs",
                         ptrs.size - 1, time.perf_counter() - t)
        except _KrylovNotApplicable as e:
            if eig == "krylov":
                raise NumericalInconsistency(f"Krylov block-diagonalisation not applicable: {e}") from e
            log.debug("Krylov path not applicable (%s); using syevd", e)
            draws.rewind()
            sizes = None
    if sizes is None:
        t = time.perf_counter()
        vals, ptrs, kroot = eigen_decomposition(P, atol=atol, rand=draws, ctx=ctx)
        if verbose:
            log.info("Determining eigen-decomposition over Float64... %.3fs", time.perf_counter() - t)
        t = time.perf_counter()
        sizes = ctx.irreducible(draws(P.nparts), ptrs, kroot, atol)      # :306, clamptol! :39
        if verbose:
            log.info("Determining the algebra-isomorphism... %.3fs", time.perf_counter() - t)
        P._eig_mode = "syevd"
    P._blk_sizes = sizes
    P._ptrs, P._kroot = ptrs, kroot
    if not fetch:
        return sizes
    return ctx.get_qhat(sizes)


def check_block_sizes(sizes, P: Partition, complex: bool = False):
    """src/diagonalize.jl:1-23"""
    final = sum(int(s) ** 2 for s in sizes) if complex else sum(int(s) * (int(s) + 1) // 2 for s in sizes)
    if final != P.nparts:
        log.error("Dimension mismatch: (final_dim=%d, block_sizes=%s) expected_dim=%d", final, list(sizes), P.nparts)
        raise DimensionMismatch(
            "Decomposition failed potentially due to\n"
            "* Rounding error (try different epsilons and/or try again) or\n"
            "* Algebra is not block-diagonalizable over the reals (retry with complex type).")


def basis_image(Qhat: List[np.ndarray], P: Partition, *, atol: Optional[float] = None):
    """``basis_image(Q, P)`` (src/diagonalize.jl:64-89): blks[i][k] = Q_k' 1[P==i+1] Q_k."""
    n = P.matrix.shape[0]
    if atol is None:
        atol = 1e-12 * n
    ctx = P._context()
    ctx.set_qhat(Qhat)
    return ctx.basis_image([q.shape[1] for q in Qhat], atol, dim=P.nparts)


def blockDiagonalize(P: Partition, verbose: bool = True, *, epsilon: float = RTOL_DEFAULT,
                     complex: bool = False, rand: Optional[Callable] = None, eig: str = "auto"):
    """src/compat.jl:26-68.  Returns ``(blkSizes, blks)``; ``blks[i][k]`` is the image of
    the basis element ``P.matrix == i+1`` in block k."""
    rand = rand or _default_rand()
    if complex:
        return _block_diagonalize_complex(P, verbose, epsilon, rand)
    ctx = P._context()
    n = ctx.n
    sizes = diagonalize(P, verbose=verbose, atol=epsilon, rand=rand, fetch=False, ctx=ctx, eig=eig)
    check_block_sizes(sizes, P, False)
    t = time.perf_counter()
    blks = ctx.basis_image(sizes, 1e-12 * n, dim=P.nparts)      # atol default, not epsilon (Appendix C)
    if verbose:
        log.info("Calculating image of the basis of the algebra... %.3fs", time.perf_counter() - t)
    return BlockDiagonalization([int(s) for s in sizes], blks)


# ----------------------------------------------------------------------------
# complex path: src/diagonalize.jl:25-40 with T = ComplexF64
# ----------------------------------------------------------------------------
def _crand(rand):
    """``rand(ComplexF64, k)``: re and im uniform in [0,1), taken pairwise from ``rand``."""
    def f(k):
        z = np.asarray(rand(2 * int(k)), dtype=np.float64)
        return z[0::2] + 1j * z[1::2]
    return f


def _diagonalize_complex(P: Partition, *, verbose=False, atol=None, rand=None, fetch=True):
    n = P.matrix.shape[0]
    if atol is None:
        atol = 1e-12 * n
    Pd = desymmetrize(P, verbose=verbose, rand=rand)                 # default atol        (:26-28)
    ctx = Pd._context()
    crand = _crand(rand)
    vals = ctx.eig_complex(crand(Pd.nparts))                         # src/eigen_decomposition.jl:242-254
    ptrs = eigen_clusters(vals, atol)
    norms = ctx.block_norms_complex(crand(Pd.nparts), ptrs)          # :259, :203-204
    kroot = _isomorphism_classes(norms, atol)
    sizes = ctx.irreducible_complex(crand(Pd.nparts), ptrs, kroot, atol)   # :306, clamptol! src/diagonalize.jl:39
    Pd._blk_sizes, Pd._ptrs, Pd._kroot = sizes, ptrs, kroot
    P._complex_partition = Pd
    if not fetch:
        return sizes
    return ctx.get_qhat_complex(sizes)


def _block_diagonalize_complex(P: Partition, verbose, epsilon, rand):
    """``blockDiagonalize(ComplexF64, P)`` (src/compat.jl:46-68)."""
    sizes = _diagonalize_complex(P, verbose=verbose, atol=epsilon, rand=rand, fetch=False)
    Pd = P._complex_partition
    # the reference desymmetrizes once more, with atol = epsilon (src/compat.jl:54-57); both runs
    # converge to the same canonical partition w.p. 1, and the draws are consumed as there
    P2 = desymmetrize(P, verbose=verbose, atol=epsilon, rand=rand)
    check_block_sizes(sizes, P2, True)
    if P2.nparts != Pd.nparts or not np.array_equal(P2.matrix, Pd.matrix):
        raise NumericalInconsistency("desymmetrize did not reproduce its own partition; try again")
    P2.release()
    n = P.matrix.shape[0]
    blks = Pd._context().basis_image_complex(sizes, 1e-12 * n, dim=Pd.nparts)
    return BlockDiagonalization([int(s) for s in sizes], blks)


# ----------------------------------------------------------------------------
# consumers of the path
# ----------------------------------------------------------------------------
def reduce_problem(P: Partition, C, A, b):
    """The reduced SDP data of README.md:57-60 / test/sd_problems.jl:32-37:
    ``newA = A*PMat``, ``newB = b``, ``newC = C'*PMat`` with ``PMat[:, i] = vec(P.matrix .== i+1)``."""
    ctx = P._context()
    if not getattr(ctx, "_constraints_set", False):
        ctx.set_constraints(A)
    Cv = np.asarray(C.todense()).reshape(-1) if hasattr(C, "todense") else C
    newA, newC = ctx.reduce_problem(Cv, A.shape[0])
    return newA, np.asarray(b, dtype=np.float64).copy(), newC


def _constraints(P: Partition):
    """``_constraints(P)`` (src/diagonalize.jl:42-50): for every class the 0-based column-major linear
    indices of its entries, ascending.  Pure host bookkeeping on the exported label matrix."""
    flat = np.asarray(P.matrix).reshape(-1, order="F")
    order = np.argsort(flat, kind="stable")
    bounds = np.searchsorted(flat[order], np.arange(1, P.nparts + 2))
    return [order[bounds[i]:bounds[i + 1]].astype(np.uint32) for i in range(P.nparts)]
