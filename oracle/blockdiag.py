"""Oracle (TEST INFRASTRUCTURE) for block-diagonalization.

Restates src/eigen_decomposition.jl, src/diagonalize.jl and the
``blockDiagonalize`` driver of src/compat.jl:26-68 in numpy.  See
``oracle/__init__.py`` for who may import this module.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Tuple

import numpy as np

from .jordan import RTOL_DEFAULT, Partition, desymmetrize, fill


class InvalidDecompositionField(Exception):
    """src/eigen_decomposition.jl:140-150."""


class NumericalInconsistency(Exception):
    """src/eigen_decomposition.jl:152-161."""


class DimensionMismatch(Exception):
    """Thrown by check_block_sizes, src/diagonalize.jl:1-23."""


# ----------------------------------------------------------------------------
# eigenvalue clustering  (EigenDecomposition ctor, src/eigen_decomposition.jl:19-40)
# ----------------------------------------------------------------------------
def eigen_clusters(values: np.ndarray, atol: float) -> np.ndarray:
    """0-based ``ptrs``: a new cluster starts at i+1 when |v[i+1]-v[i]| > atol
    (``isapprox`` with atol>0 has rtol=0)."""
    values = np.asarray(values)
    if values.size == 0:
        return np.array([0], dtype=np.int64)
    brk = np.flatnonzero(np.abs(values[1:] - values[:-1]) > atol) + 1
    return np.concatenate([[0], brk, [values.size]]).astype(np.int64)


# ----------------------------------------------------------------------------
# Otsu threshold on a 16-bin log histogram  (src/eigen_decomposition.jl:83-139)
# ----------------------------------------------------------------------------
def log_histogram(X: np.ndarray, num_bins: int, atol: float):
    a = np.abs(np.asarray(X, dtype=np.float64)).reshape(-1)
    min_val, max_val = a.min(), a.max()
    if min_val < atol:
        min_val = atol
    assert min_val > 0
    edges = np.exp(np.linspace(math.log(min_val), math.log(max_val), num_bins + 1))
    # k = (first edge index (1-based) with edge > x, or num_bins+1) - 1, clamped to 1..num_bins
    first_gt = np.searchsorted(edges, a, side="right")      # 0-based index of first edge > x
    # edges need not be strictly monotone if min==max; emulate findfirst literally then
    if not np.all(np.diff(edges) > 0):
        first_gt = np.array([next((i for i, e in enumerate(edges) if e > x), num_bins)
                             for x in a])
    k = np.clip(first_gt, 1, num_bins)                       # 1-based bin index
    counts = np.bincount(k - 1, minlength=num_bins)[:num_bins]
    return counts, edges


def otsu_threshold(X: np.ndarray, atol: float) -> float:
    n_bins = max(int(math.ceil(-math.log10(np.finfo(np.float64).eps))), 4)   # 16
    counts, edges = log_histogram(X, n_bins, atol)
    pdf = counts / counts.sum()
    w = np.cumsum(pdf)
    mu = np.cumsum(np.log(edges[:-1]) * pdf)
    muT = mu[-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        s2 = (muT * w - mu) ** 2 / (w * (1 - w))
    cand = s2[:-1]
    # Julia argmax: NaN is maximal, first one wins
    nan = np.flatnonzero(np.isnan(cand))
    k = int(nan[0]) if nan.size else int(np.argmax(cand))
    return float(edges[k + 1])


# ----------------------------------------------------------------------------
# IntDisjointSets (DataStructures.jl 0.18): union by rank, ties -> first root
# ----------------------------------------------------------------------------
class IntDisjointSets:
    def __init__(self, n: int):
        self.parents = list(range(n))
        self.ranks = [0] * n

    def find_root(self, x: int) -> int:
        p = self.parents
        root = x
        while p[root] != root:
            root = p[root]
        while p[x] != root:          # full path compression
            p[x], x = root, p[x]
        return root

    def union(self, x: int, y: int) -> int:
        xr, yr = self.find_root(x), self.find_root(y)
        if xr == yr:
            return xr
        if self.ranks[xr] < self.ranks[yr]:
            xr, yr = yr, xr
        elif self.ranks[xr] == self.ranks[yr]:
            self.ranks[xr] += 1
        self.parents[yr] = xr
        return xr

    def __len__(self):
        return len(self.parents)


def is_consistent(K: IntDisjointSets) -> bool:
    """src/eigen_decomposition.jl:163-167: every root is the smallest index of its class."""
    part = [K.find_root(i) for i in range(len(K))]
    seen = {}
    for i, r in enumerate(part):
        seen.setdefault(r, i)
    return all(r == first for r, first in seen.items())


# ----------------------------------------------------------------------------
# Murota et al. Alg. 4.1
# ----------------------------------------------------------------------------
def block_norms_inf(W: np.ndarray, ptrs: np.ndarray) -> np.ndarray:
    """src/eigen_decomposition.jl:177-193 with p = Inf (max |entry| of each block,
    zero for eigenspaces of different dimension)."""
    ne = len(ptrs) - 1
    dims = np.diff(ptrs)
    out = np.zeros((ne, ne))
    aW = np.abs(W)
    # reduce rows then columns by cluster
    rowmax = np.maximum.reduceat(aW, ptrs[:-1], axis=0)
    blk = np.maximum.reduceat(rowmax, ptrs[:-1], axis=1)
    for i in range(ne):
        for j in range(i, ne):
            if dims[i] == dims[j]:
                out[i, j] = out[j, i] = blk[i, j]
    return out


def isomorphism_partition(Q: np.ndarray, ptrs: np.ndarray, A: np.ndarray, atol: float):
    """src/eigen_decomposition.jl:201-219."""
    W = Q.T @ A @ Q
    norms = block_norms_inf(W, ptrs)
    thr = otsu_threshold(norms, atol)
    ne = len(ptrs) - 1
    K = IntDisjointSets(ne)
    for i in range(ne):
        for j in range(i + 1, ne):
            if norms[i, j] >= thr:
                K.union(i, j)
    return K, norms, thr


def eigen_decomposition(P: Partition, rand: Callable[[int], np.ndarray], atol: float):
    """src/eigen_decomposition.jl:236-273.  Returns (values, Q, ptrs, K)."""
    A = fill(P, rand(P.nparts))                         # :242
    if not np.array_equal(A, A.T):
        # Julia's ``eigen`` of a non-symmetric real matrix returns complex
        # eigenvalues (generically) -> convert fails -> InvalidDecompositionField
        w = np.linalg.eigvals(A)
        if np.abs(w.imag).max() > 0:
            raise InvalidDecompositionField("Float64 requested, complex eigenvalues found")
        raise InvalidDecompositionField("non-symmetric generic element")
    vals, Q = np.linalg.eigh(A)                         # :246 (ascending)
    ptrs = eigen_clusters(vals, atol)                   # :254
    A = fill(P, rand(P.nparts))                         # :259
    K, norms, thr = isomorphism_partition(Q, ptrs, A, atol)
    if not is_consistent(K):
        raise NumericalInconsistency("eigen_decomposition: K-partition inconsistent")
    return vals, Q, ptrs, K


def irreducible_decomposition(Q: np.ndarray, ptrs: np.ndarray, K: IntDisjointSets,
                              P: Partition, rand: Callable[[int], np.ndarray]) -> List[np.ndarray]:
    """src/eigen_decomposition.jl:295-348."""
    ne = len(ptrs) - 1
    kpart = [K.find_root(i) for i in range(ne)]
    roots = list(dict.fromkeys(kpart))
    A = fill(P, rand(P.nparts))                         # :306
    out = []
    for i in roots:
        Ki = [t for t in range(ne) if kpart[t] == i]
        assert Ki[0] == i
        Qi = Q[:, ptrs[i]:ptrs[i + 1]]
        if len(Ki) == 1:
            out.append(Qi[:, :1].copy())
            continue
        cols = [Qi[:, 0].copy()]
        for j in Ki[1:]:
            Qj = Q[:, ptrs[j]:ptrs[j + 1]]
            Pblk = (Qi.T @ A @ Qj).T                   # block(A, Ei, Ej)'
            Pblk = Pblk / np.linalg.norm(Pblk[0, :])
            cols.append(Qj @ Pblk[:, 0])
        out.append(np.stack(cols, axis=1))
    return out


def diagonalize(P: Partition, rand: Callable[[int], np.ndarray], atol: Optional[float] = None,
                complex_: bool = False) -> List[np.ndarray]:
    """src/diagonalize.jl:25-40 (real path; the complex path is in
    ``diagonalize_complex``)."""
    if complex_:
        return diagonalize_complex(P, rand, atol)
    n = P.matrix.shape[0]
    if atol is None:
        atol = 1e-12 * n
    vals, Q, ptrs, K = eigen_decomposition(P, rand, atol)
    Qhat = irreducible_decomposition(Q, ptrs, K, P, rand)
    return [np.where(np.abs(q) < atol, 0.0, q) for q in Qhat]


def check_block_sizes(Qhat: List[np.ndarray], P: Partition, complex_: bool = False):
    """src/diagonalize.jl:1-23."""
    sizes = [q.shape[1] for q in Qhat]
    final = sum(s * s for s in sizes) if complex_ else sum(s * (s + 1) // 2 for s in sizes)
    if final != P.nparts:
        raise DimensionMismatch(f"final_dim={final} block_sizes={sizes} expected={P.nparts}")


def basis_image(Qhat: List[np.ndarray], P: Partition, atol: Optional[float] = None):
    """src/diagonalize.jl:64-89: blks[i][k] = Qk' * 1[P==i+1] * Qk, clamped at 1e-12*N."""
    n = P.matrix.shape[0]
    if atol is None:
        atol = 1e-12 * n
    out = []
    for i in range(1, P.nparts + 1):
        B = (P.matrix == i).astype(np.float64)
        row = []
        for q in Qhat:
            m = q.conj().T @ (B @ q)
            m = np.where(np.abs(m) < atol, 0.0, m)
            row.append(m)
        out.append(row)
    return out


def blockDiagonalize(P: Partition, rand: Callable[[int], np.ndarray],
                     epsilon: float = RTOL_DEFAULT, complex: bool = False):
    """src/compat.jl:26-68.  Returns (blkSizes, blks)."""
    Pc = P.copy()
    Qhat = diagonalize(Pc, rand, atol=epsilon, complex_=complex)
    if complex:
        P = desymmetrize(P.copy(), rand, atol=epsilon)      # src/compat.jl:54-57
    check_block_sizes(Qhat, P, complex)
    blks = basis_image(Qhat, P)                             # atol = 1e-12*N (Appendix C)
    return [q.shape[1] for q in Qhat], blks


# ----------------------------------------------------------------------------
# complex path (next-tier row (f)1; src/diagonalize.jl:26-28, generic ``eigen``)
# ----------------------------------------------------------------------------
def diagonalize_complex(P: Partition, rand, atol: Optional[float] = None) -> List[np.ndarray]:
    n = P.matrix.shape[0]
    if atol is None:
        atol = 1e-12 * n
    P = desymmetrize(P, rand)                               # default atol (:27)

    def crand(k):                                           # rand(ComplexF64, k)
        z = rand(2 * k)
        return z[0::2] + 1j * z[1::2]

    lut = lambda r: np.concatenate([[0.0], r])[P.matrix]
    A = lut(crand(P.nparts))
    w, V = np.linalg.eig(A)
    order = np.lexsort((w.imag, w.real))                    # Julia sorts by (re, im)
    w, V = w[order], V[:, order]
    ptrs = eigen_clusters(w, atol)
    A = lut(crand(P.nparts))
    ne = len(ptrs) - 1
    W = V.conj().T @ A @ V
    norms = block_norms_inf(W, ptrs)
    thr = otsu_threshold(norms, atol)
    K = IntDisjointSets(ne)
    for i in range(ne):
        for j in range(i + 1, ne):
            if norms[i, j] >= thr:
                K.union(i, j)
    if not is_consistent(K):
        raise NumericalInconsistency("eigen_decomposition: K-partition inconsistent")
    kpart = [K.find_root(i) for i in range(ne)]
    roots = list(dict.fromkeys(kpart))
    A = lut(crand(P.nparts))
    out = []
    for i in roots:
        Ki = [t for t in range(ne) if kpart[t] == i]
        Qi = V[:, ptrs[i]:ptrs[i + 1]]
        if len(Ki) == 1:
            out.append(Qi[:, :1].copy())
            continue
        cols = [Qi[:, 0].copy()]
        for j in Ki[1:]:
            Qj = V[:, ptrs[j]:ptrs[j + 1]]
            Pblk = (Qi.conj().T @ A @ Qj).conj().T
            Pblk = Pblk / np.linalg.norm(Pblk[0, :])
            cols.append(Qj @ Pblk[:, 0])
        out.append(np.stack(cols, axis=1))
    return [np.where(np.abs(q) < atol, 0.0, q) for q in out]
