"""Oracle (TEST INFRASTRUCTURE) for partitions and the Jordan-reduction loop.

Restates, line by line, src/utils.jl:14-81, src/abstract_part.jl:97-110 and
src/partitions.jl:6-223 of the reference in numpy.  See ``oracle/__init__.py``
for the rules about who may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import numpy as np

try:  # scipy is only needed for sparse constraint matrices
    import scipy.sparse as _sp
except Exception:  # pragma: no cover
    _sp = None

RTOL_DEFAULT = math.sqrt(np.finfo(np.float64).eps)  # Base.rtoldefault(Float64)


# ----------------------------------------------------------------------------
# rounding  (src/utils.jl:14-53)
# ----------------------------------------------------------------------------
def clamptol(a: np.ndarray, atol: float = RTOL_DEFAULT) -> np.ndarray:
    """src/utils.jl:14-27: abs(f) < atol -> 0."""
    a = np.asarray(a)
    return np.where(np.abs(a) < atol, 0.0, a)


def sigdigits_of(atol: float) -> int:
    """src/utils.jl:37: floor(Int, -log10(atol))."""
    return int(math.floor(-math.log10(atol)))


def unsafe_round(f: np.ndarray, scale: int) -> np.ndarray:
    """src/utils.jl:49-53: frexp -> trunc(scale*x)/scale -> ldexp.

    One IEEE multiply, truncation toward zero, one IEEE divide.
    """
    x, n = np.frexp(np.asarray(f, dtype=np.float64))
    q = np.trunc(np.float64(scale) * x)          # unsafe_trunc(Int, scale*x)
    y = q / np.float64(scale)                     # Int / Int -> Float64 division
    return np.ldexp(y, n)


def clamp_round(A: np.ndarray, atol: float = RTOL_DEFAULT,
                sigdigits: Optional[int] = None) -> np.ndarray:
    """src/utils.jl:34-47 (_clamp_round!), out of place."""
    A = np.asarray(A, dtype=np.float64)
    if sigdigits is None:
        sigdigits = sigdigits_of(atol)
    r = unsafe_round(A, 10 ** sigdigits)
    return np.where(np.abs(A) < atol, 0.0, r)


def symmetrize(v: np.ndarray, n: int) -> np.ndarray:
    """src/utils.jl:71-81: (M[i,j] + M[j,i]) / 2 (one add, one divide)."""
    M = np.asarray(v, dtype=np.float64).reshape(n, n, order="F")
    S = (M + M.T) / 2
    return S.reshape(-1, order="F")


# ----------------------------------------------------------------------------
# Partition  (src/partitions.jl:6-75)
# ----------------------------------------------------------------------------
@dataclass
class Partition:
    """src/partitions.jl:6-9.  ``matrix`` holds labels 0..nparts (int64 here;
    the reference's element type only matters for its overflow behaviour)."""

    nparts: int
    matrix: np.ndarray

    @property
    def shape(self):
        return self.matrix.shape

    def copy(self) -> "Partition":
        return Partition(self.nparts, self.matrix.copy())

    def __eq__(self, other) -> bool:  # src/partitions.jl:16-17
        return (isinstance(other, Partition) and self.nparts == other.nparts
                and np.array_equal(self.matrix, other.matrix))


def dim(P: Partition) -> int:
    return P.nparts


def _first_occurrence_rank(keys_colmajor: np.ndarray, zero_key) -> Tuple[np.ndarray, int]:
    """Number distinct keys 1,2,... by first occurrence; ``zero_key`` -> 0.

    This is the common core of ``Partition{T}(M)`` (Dict pass, src/partitions.jl:
    24-35) and ``__sort_unique!`` (``unique`` + LUT, :44-60).
    """
    uniq, first, inv = np.unique(keys_colmajor, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # distinct keys by first occurrence
    is_zero = uniq[order] == zero_key
    rank_sorted = np.cumsum(~is_zero)                 # 1,2,... skipping the zero key
    rank_sorted = np.where(is_zero, 0, rank_sorted)
    rank = np.empty(uniq.size, dtype=np.int64)
    rank[order] = rank_sorted
    nparts = int((~is_zero).sum())
    return rank[inv.reshape(-1)], nparts


def partition_from_values(M: np.ndarray) -> Partition:
    """``Partition{T}(M::AbstractMatrix)`` for non-integer M, src/partitions.jl:24-35.

    Keys are compared with ``isequal`` (bitwise: -0.0 != 0.0); the value
    ``zero(eltype(M))`` = +0.0 is pre-seeded with label 0.
    """
    M = np.asarray(M)
    if np.issubdtype(M.dtype, np.integer):
        return partition_from_labels(M)
    M = np.asarray(M, dtype=np.float64)
    bits = M.reshape(-1, order="F").view(np.uint64)
    lab, nparts = _first_occurrence_rank(bits, np.uint64(0))
    return Partition(nparts, lab.reshape(M.shape, order="F"))


def sort_unique(P: Partition) -> Partition:
    """``__sort_unique!``, src/partitions.jl:44-60: renumber by first occurrence
    (column-major), 0 preserved."""
    flat = np.asarray(P.matrix, dtype=np.int64).reshape(-1, order="F")
    assert flat.size == 0 or flat.min() >= 0
    lab, nparts = _first_occurrence_rank(flat, 0)
    return Partition(nparts, lab.reshape(P.matrix.shape, order="F"))


def partition_from_labels(M: np.ndarray) -> Partition:
    """``Partition{T}(M::AbstractMatrix{<:Integer})``, src/partitions.jl:37-42."""
    return sort_unique(Partition(0, np.asarray(M, dtype=np.int64)))


def refine(P1: Partition, P2: Partition) -> Partition:
    """``refine!``, src/partitions.jl:62-66: p1 + p2*(dim(P1)+1), then renumber."""
    assert P1.matrix.shape == P2.matrix.shape
    m = P1.matrix.astype(np.int64) + P2.matrix.astype(np.int64) * (P1.nparts + 1)
    return sort_unique(Partition(0, m))


def fill(P: Partition, values: np.ndarray) -> np.ndarray:
    """``fill!``, src/partitions.jl:68-75: M[idx] = 0 if label==0 else values[label]."""
    values = np.asarray(values, dtype=np.float64)
    assert values.shape == (P.nparts,)
    lut = np.concatenate([[0.0], values])
    return lut[P.matrix]


def randomize(P: Partition, rand: Callable[[int], np.ndarray]) -> np.ndarray:
    """``randomize!``, src/abstract_part.jl:107-110: values = rand(dim(P))."""
    return fill(P, rand(P.nparts))


# ----------------------------------------------------------------------------
# projection onto rowspace(A)  (src/utils.jl:59-69, src/partitions.jl:124-126)
# ----------------------------------------------------------------------------
class RowSpaceProjector:
    """proj(v) = A' * (qr(A') \\ v).

    The least-squares coefficients come from a QR of A' (dense, small problems)
    or from the Gram matrix A A' (large / sparse problems); ``A' * coef`` is
    applied per entry in ascending constraint order with separate multiply and
    add, which is what Julia's sparse ``mul!`` does and makes
    entries with equal constraint pattern bit-equal.
    """

    def __init__(self, A):
        if _sp is not None and _sp.issparse(A):
            self.A = A.tocsr()
            self.A.sort_indices()
            self.sparse = True
        else:
            self.A = np.asarray(A, dtype=np.float64)
            self.sparse = False
        self.m, self.n2 = self.A.shape
        G = self.A @ self.A.T
        self.G = np.asarray(G.todense() if self.sparse else G, dtype=np.float64)
        self._qr = None
        # Rank-deficient A (redundant constraint rows): the reference's sparse path is SuiteSparse's
        # rank-revealing QR, whose projection is A' G^+ A; any least-squares solution of G c = A v gives
        # the same A' c, so the minimum-norm one (lstsq) is used.
        sv = np.linalg.svd(self.G, compute_uv=False)
        self.rank_deficient = bool(sv.size and sv[-1] <= 1e-12 * sv[0])
        if not self.sparse and self.m * self.m * self.n2 <= 2e8 and not self.rank_deficient:
            self._qr = np.linalg.qr(self.A.T)       # (N^2 x m) reduced QR

    def coefficients(self, v: np.ndarray) -> np.ndarray:
        if self._qr is not None:
            Q, R = self._qr
            return np.linalg.solve(R, Q.T @ v)
        return self.solve_gram(np.asarray(self.A @ v).reshape(-1))

    def solve_gram(self, b: np.ndarray) -> np.ndarray:
        if self.rank_deficient:
            return np.linalg.lstsq(self.G, b, rcond=1e-12)[0]
        return np.linalg.solve(self.G, b)

    def apply_transpose(self, coef: np.ndarray) -> np.ndarray:
        t = np.zeros(self.n2)
        if self.sparse:
            ip, ix, vv = self.A.indptr, self.A.indices, self.A.data
            for k in range(self.m):
                sl = slice(ip[k], ip[k + 1])
                t[ix[sl]] = t[ix[sl]] + vv[sl] * coef[k]
        else:
            for k in range(self.m):
                row = self.A[k]
                nz = row != 0
                t[nz] = t[nz] + row[nz] * coef[k]
        return t

    def __call__(self, v: np.ndarray) -> np.ndarray:
        return self.apply_transpose(self.coefficients(v))


def init_elements(C, A, b, atol: float = RTOL_DEFAULT, snap_decimals: Optional[int] = 12,
                  projector: Optional[RowSpaceProjector] = None):
    """The two initial elements CL and X0 of src/partitions.jl:129-142.

    CL = symmetrize(round(C - proj(C)));  X0 = round(proj(symmetrize(x0))) with
    x0 the minimum-norm solution of A x = b (``Krylov.craig``, :137).

    ``snap_decimals``: both vectors are rounded to that many decimals before
    the reference's truncation.  Truncation splits values that sit exactly on a
    grid point (1/16 on esc16j) under +-1 ulp noise; the reference gets
    consistent bits from SPQR which cannot be reproduced without Julia
    (SURVEY.md fact 5).  ``None`` disables the safeguard.
    """
    C = np.asarray(C.todense()).reshape(-1) if (_sp is not None and _sp.issparse(C)) \
        else np.asarray(C, dtype=np.float64).reshape(-1)
    n = math.isqrt(C.size)
    assert n * n == C.size
    proj = projector or RowSpaceProjector(A)
    c = C - proj(C)
    if snap_decimals is not None:
        c = np.round(c, snap_decimals)
    c = clamp_round(c, atol)
    CL = symmetrize(c, n).reshape(n, n, order="F")

    x0 = proj.apply_transpose(proj.solve_gram(np.asarray(b, dtype=np.float64)))
    x = symmetrize(x0, n)
    x = proj(x)
    if snap_decimals is not None:
        x = np.round(x, snap_decimals)
    X0 = clamp_round(x, atol).reshape(n, n, order="F")
    return CL, X0, proj


# ----------------------------------------------------------------------------
# admissible_subspace  (src/partitions.jl:109-190)
# ----------------------------------------------------------------------------
def admissible_subspace_trace(C, A, b, rand: Callable[[int], np.ndarray],
                              atol: float = RTOL_DEFAULT, snap_decimals: Optional[int] = 12,
                              record: bool = False, max_iter: int = 10 ** 6):
    """Run the reference loop; returns (Partition, trace).

    ``trace`` has the initial dim and, per iteration, (dim after projection,
    dim after square); with ``record=True`` it also keeps the label matrix and
    the un-rounded refining matrices of every step (inputs for kernel-level
    parity tests).
    """
    CL, X0, proj = init_elements(C, A, b, atol, snap_decimals)
    n = CL.shape[0]
    S = partition_from_values(CL)                     # :145
    S = refine(S, partition_from_values(X0))          # :146
    maxdim = (n * n + n) // 2
    cur = S.nparts
    trace = {"init": cur, "iters": [], "steps": []}
    it = 0
    while cur < maxdim and it < max_iter:
        it += 1
        r_a = rand(S.nparts)
        X = fill(S, r_a)                              # :159
        x = X.reshape(-1, order="F")
        x = x - proj(x)                               # :161
        if record:
            trace["steps"].append(("project", S.matrix.copy(), r_a.copy(),
                                   x.reshape(n, n, order="F").copy()))
        x = clamp_round(x, atol)                      # :162
        X = x.reshape(n, n, order="F")
        S = refine(S, partition_from_values(X))       # :164
        d_proj = S.nparts
        if cur != S.nparts:                           # :166-168
            r_b = rand(S.nparts)
            X = fill(S, r_b)
        X2 = X @ X                                    # :172
        if record:
            trace["steps"].append(("square", S.matrix.copy(), X.copy(), X2.copy()))
        X2 = clamp_round(X2, atol)                    # :173
        S = refine(S, partition_from_values(X2))      # :174
        trace["iters"].append((d_proj, S.nparts))
        if cur == S.nparts:                           # :180-182
            break
        cur = S.nparts
    trace["iterations"] = it
    return S, trace


def admissible_subspace(C, A, b, rand, atol: float = RTOL_DEFAULT,
                        snap_decimals: Optional[int] = 12) -> Partition:
    return admissible_subspace_trace(C, A, b, rand, atol, snap_decimals)[0]


# ----------------------------------------------------------------------------
# desymmetrize  (src/partitions.jl:197-223)
# ----------------------------------------------------------------------------
def desymmetrize(P: Partition, rand: Callable[[int], np.ndarray],
                 atol: float = RTOL_DEFAULT) -> Partition:
    P = P.copy()
    cur = P.nparts
    while True:
        X = randomize(P, rand)                        # :210
        Y = randomize(P, rand)                        # :211
        XY = clamp_round(X @ Y, atol)                 # :212-213
        P = refine(P, partition_from_values(XY))      # :214
        if cur == P.nparts:
            break
        cur = P.nparts
    return P
