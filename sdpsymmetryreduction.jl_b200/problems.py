"""SDP problem generators ``(C, A, b)`` for the Jordan-reduction hot path.

These restate the *consumers* of the reference (layer L6 in SURVEY.md): the
problem builders that live in the reference's tests and docs, not in its
package.  They are shipped here because the benchmark and the parity tests need
inputs of the shapes BASELINE.json names.

Conventions (identical to the reference): all matrices are column-major,
``vec(M)[i + N*j] = M[i, j]``; ``C`` has length N^2; ``A`` is m x N^2 (dense
``ndarray`` or ``scipy.sparse.csr_matrix``); ``b`` has length m.

Reference sites:
  * Theta' of the Erdos-Renyi polarity graph ER(q): test/sd_problems.jl:16-27
  * QAP relaxation: test/sd_problems.jl:63-105, reader test/qap.jl:3-11
  * Theta' form for an arbitrary graph: README.md:45-48
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import Optional

import numpy as np
import scipy.sparse as sp


@dataclass
class SDPProblem:
    """A symmetric SDP  min <C,x>  s.t.  A x = b, Mat(x) psd  (README.md:31-38)."""

    name: str
    C: np.ndarray            # (N^2,) float64
    A: object                # (m, N^2) ndarray or scipy.sparse matrix
    b: np.ndarray            # (m,) float64
    n: int                   # matrix order N
    expected_dim: Optional[int] = None          # known dim of the final partition
    expected_blocks: Optional[list] = None      # known sorted block sizes
    expected_mult: Optional[list] = None        # known multiplicities (sorted)
    meta: Optional[dict] = None

    def __iter__(self):  # ``admissible_subspace(*problem)`` like ``CAb...``
        return iter((self.C, self.A, self.b))


# ----------------------------------------------------------------------------
# graphs
# ----------------------------------------------------------------------------
def er_graph_adjacency(q: int) -> np.ndarray:
    """Adjacency of the Erdos-Renyi polarity graph ER(q), q prime.

    Points of PG(2,q) in the order of test/sd_problems.jl:17-19, adjacent iff
    orthogonal mod q and distinct (test/sd_problems.jl:20).
    """
    pts = [(0, 0, 1)]
    pts += [(0, 1, b) for b in range(q)]
    pts += [(1, a, b) for a in range(q) for b in range(q)]
    P = np.array(pts, dtype=np.int64)
    G = (P @ P.T) % q == 0
    np.fill_diagonal(G, False)     # x != y  (vectors are pairwise distinct)
    return G


def kneser_adjacency(n: int, k: int) -> np.ndarray:
    """Kneser graph K(n,k): k-subsets of range(n), adjacent iff disjoint."""
    masks = np.array(
        [sum(1 << e for e in c) for c in itertools.combinations(range(n), k)],
        dtype=np.int64,
    )
    return (masks[:, None] & masks[None, :]) == 0


def kneser_intersection_sizes(n: int, k: int) -> np.ndarray:
    """|u & v| for all pairs of k-subsets (the Johnson-scheme relation)."""
    masks = np.array(
        [sum(1 << e for e in c) for c in itertools.combinations(range(n), k)],
        dtype=np.int64,
    )
    x = masks[:, None] & masks[None, :]
    return _popcount64(x)


def _popcount64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)
    if hasattr(np, "bitwise_count"):
        return np.bitwise_count(x).astype(np.int64)
    c = np.zeros(x.shape, dtype=np.int64)
    while True:
        nz = x != 0
        if not nz.any():
            return c
        c += (x & np.uint64(1)).astype(np.int64)
        x = x >> np.uint64(1)


def hamming_distance_matrix(d: int, q: int, dtype=np.int8) -> np.ndarray:
    """Hamming distance between all pairs of words of ``range(q)^d``.

    Vertex order = ``itertools.product(range(q), repeat=d)`` (first digit most
    significant), as in SURVEY.md section 8(d) cfg 3.
    """
    N = q ** d
    idx = np.arange(N, dtype=np.int64)
    D = np.zeros((N, N), dtype=dtype)
    for pos in range(d):
        dig = ((idx // (q ** pos)) % q).astype(np.int16)
        D += (dig[:, None] != dig[None, :]).astype(dtype)
    return D


def hamming_adjacency(d: int, q: int) -> np.ndarray:
    return hamming_distance_matrix(d, q) == 1


# ----------------------------------------------------------------------------
# Theta' SDP of a graph
# ----------------------------------------------------------------------------
def theta_prime(adj: np.ndarray, name: str = "theta'", sparse: Optional[bool] = None,
                **expect) -> SDPProblem:
    """Theta' SDP: C = ones, A = [vec(Adj)'; vec(I)'], b = [0, 1].

    Dense ``A`` mirrors test/sd_problems.jl:23-25; for large N pass
    ``sparse=True`` to get the same matrix as CSR (what a Julia user would pass
    as ``SparseMatrixCSC``).
    """
    adj = np.asarray(adj, dtype=bool)
    N = adj.shape[0]
    assert adj.shape == (N, N)
    if sparse is None:
        sparse = N > 512
    C = np.ones(N * N, dtype=np.float64)
    b = np.array([0.0, 1.0])
    if not sparse:
        A = np.empty((2, N * N), dtype=np.float64)
        A[0] = adj.reshape(-1, order="F")
        A[1] = np.eye(N).reshape(-1, order="F")
    else:
        e = np.flatnonzero(adj.reshape(-1, order="F")).astype(np.int64)
        dg = np.arange(N, dtype=np.int64) * (N + 1)
        indptr = np.array([0, e.size, e.size + N], dtype=np.int64)
        A = sp.csr_matrix(
            (np.ones(e.size + N), np.concatenate([e, dg]), indptr), shape=(2, N * N)
        )
    return SDPProblem(name=name, C=C, A=A, b=b, n=N, **expect)


def lovasz_er(q: int) -> SDPProblem:
    """test/sd_problems.jl:16-27 ; pins from test/lovasz.jl:6-8,22-24,38-40."""
    pins = {
        3: (12, [2, 2, 3], [1, 2, 3]),
        5: (15, [2, 2, 2, 3], [1, 4, 5, 5]),
        7: (18, [2, 2, 2, 2, 3], [1, 6, 6, 7, 8]),
    }
    dim, blocks, mult = pins.get(q, (None, None, None))
    return theta_prime(er_graph_adjacency(q), name=f"theta'-ER({q})", sparse=False,
                       expected_dim=dim, expected_blocks=blocks, expected_mult=mult)


def petersen() -> SDPProblem:
    """Petersen = K(5,2), N = 10 (BASELINE.json configs[0])."""
    return theta_prime(kneser_adjacency(5, 2), name="theta'-Petersen", sparse=False,
                       expected_dim=3, expected_blocks=[1, 1, 1], expected_mult=[1, 4, 5])


def kneser(n: int, k: int, sparse: Optional[bool] = None) -> SDPProblem:
    from math import comb
    # Johnson scheme J(n,k): min(k, n-k)+1 classes, all blocks 1x1
    ncls = min(k, n - k) + 1
    mult = sorted(comb(n, j) - (comb(n, j - 1) if j else 0) for j in range(ncls))
    return theta_prime(kneser_adjacency(n, k), name=f"theta'-K({n},{k})", sparse=sparse,
                       expected_dim=ncls, expected_blocks=[1] * ncls, expected_mult=mult)


def hamming(d: int, q: int, sparse: Optional[bool] = None) -> SDPProblem:
    from math import comb
    mult = sorted(comb(d, j) * (q - 1) ** j for j in range(d + 1))
    return theta_prime(hamming_adjacency(d, q), name=f"theta'-H({d},{q})", sparse=sparse,
                       expected_dim=d + 1, expected_blocks=[1] * (d + 1), expected_mult=mult,
                       meta={"d": d, "q": q})


def krawtchouk(d: int, q: int) -> np.ndarray:
    """K[i, j] = K_i(j): eigenvalue of the distance-i graph of H(d,q) on the
    j-th eigenspace (closed form used as a value pin, SURVEY.md 8(d) cfg 3)."""
    from math import comb
    K = np.zeros((d + 1, d + 1))
    for i in range(d + 1):
        for j in range(d + 1):
            K[i, j] = sum((-1) ** h * (q - 1) ** (i - h) * comb(j, h) * comb(d - j, i - h)
                          for h in range(0, i + 1))
    return K


def eberlein(v: int, k: int) -> np.ndarray:
    """E[i, j] = E_i(j): eigenvalue of the relation |A & B| = k - i of the Johnson scheme J(v,k) on the
    j-th eigenspace (Eberlein polynomial; value pin for the Kneser configs, SURVEY.md 8(d) cfg 4).
    Row k is the Kneser graph K(v,k): (-1)^j C(v-k-j, k-j)."""
    from math import comb
    d = min(k, v - k)
    E = np.zeros((d + 1, d + 1))
    for i in range(d + 1):
        for j in range(d + 1):
            E[i, j] = sum((-1) ** h * comb(j, h) * comb(k - j, i - h) * comb(v - k - j, i - h)
                          for h in range(0, i + 1) if i - h <= k - j and i - h <= v - k - j)
    return E


# ----------------------------------------------------------------------------
# QAP relaxation
# ----------------------------------------------------------------------------
def read_qapdata(path: str):
    """QAPLIB file: n, then n rows of the first matrix, n rows of the second
    (test/qap.jl:3-11)."""
    with open(path) as fh:
        tok = fh.read().split()
    n = int(tok[0])
    vals = np.array(tok[1:1 + 2 * n * n], dtype=np.float64).reshape(2 * n, n)
    return vals[:n].copy(), vals[n:].copy()


def qap_constraints(n: int):
    """The (2n+1) x n^4 constraint matrix of test/sd_problems.jl:63-93 as CSR."""
    In = np.eye(n)
    Jn = np.ones((n, n))
    rows, b = [], []
    for j in range(n):
        E = np.zeros((n, n))
        E[j, j] = 1.0
        rows.append(np.kron(In, E).reshape(-1, order="F"))
        b.append(1.0)
        if j < n - 1:       # the last one is linearly dependent on the others
            rows.append(np.kron(E, In).reshape(-1, order="F"))
            b.append(1.0)
    rows.append((np.kron(In, Jn - In) + np.kron(Jn - In, In)).reshape(-1, order="F"))
    b.append(0.0)
    rows.append(np.ones(n ** 4))
    b.append(float(n * n))
    A = sp.csr_matrix(np.vstack(rows))
    A.sort_indices()
    return A, np.array(b)


def qap(flowA: np.ndarray, flowB: np.ndarray, name: str = "qap", **expect) -> SDPProblem:
    """QuadraticAssignment(flowA, flowB) of test/sd_problems.jl:95-105."""
    n = flowA.shape[0]
    assert flowA.shape == (n, n) and flowB.shape == (n, n)
    A, b = qap_constraints(n)
    C = np.kron(flowA, flowB)
    if not np.array_equal(C, C.T):
        C = (C + C.T) / 2
    return SDPProblem(name=name, C=C.reshape(-1, order="F").copy(), A=A, b=b, n=n * n, **expect)


def qap_esc16j(npz_path: str) -> SDPProblem:
    """esc16j from the committed fixture (tests/golden/esc16j.npz, generated
    from test/qapdata/esc16j.dat by tools/make_golden.py)."""
    z = np.load(npz_path)
    return qap(z["flowA"], z["flowB"], name="qap-esc16j", expected_dim=150,
               expected_blocks=[1] * 10 + [7] * 5,
               expected_mult=sorted([1, 1, 4, 4, 6, 8, 8, 32, 32, 48, 1, 1, 4, 4, 6]))


# ----------------------------------------------------------------------------
# synthetic permutation-symmetric SDP (BASELINE.json configs[4])
# ----------------------------------------------------------------------------
def synthetic_product_scheme(fields: int = 3, bits: int = 5, m: int = 64, seed: int = 1234,
                             sparse: bool = True, keep_orbitals: bool = True,
                             chunk_cols: int = 2048) -> SDPProblem:
    """Vertices Z_2^(fields*bits); orbital of (u,v) = per-field popcounts of u^v,
    conjugated by a random vertex permutation (SURVEY.md 8(d) cfg 5).

    The closure must recover exactly the (bits+1)^fields orbitals; all blocks
    are 1x1.  Built in column chunks so that N = 32768 needs ~20 GB of host memory
    instead of ~60.
    """
    from math import comb
    nb = fields * bits
    N = 1 << nb
    rng = np.random.default_rng(seed)
    perm = rng.permutation(N)
    norb = (bits + 1) ** fields
    cvals = rng.integers(1, 10, size=norb).astype(np.float64)
    u = perm.astype(np.int64)

    def orbital_block(c0, c1):
        """orbital ids of rows 0..N-1 x columns c0..c1-1, shape (N, c1-c0)"""
        x = u[:, None] ^ u[None, c0:c1]
        o = np.zeros(x.shape, dtype=np.int32)
        for f in range(fields):
            fld = (x >> (f * bits)) & ((1 << bits) - 1)
            o = o * (bits + 1) + _popcount64(fld).astype(np.int32)
        return o

    # orbital sizes: every row of the orbital matrix has the same histogram (vertex-transitive)
    sizes = np.bincount(orbital_block(0, 1).reshape(-1), minlength=norb).astype(np.int64) * N
    order = np.argsort(sizes, kind="stable")
    assert order[0] == 0
    m = min(m, norb)
    chosen = order[:m]
    sel = np.full(norb, -1, dtype=np.int64)
    sel[chosen] = np.arange(m)

    C = np.empty(N * N, dtype=np.float64)
    orb_full = np.empty((N, N), dtype=np.int32, order="F") if keep_orbitals else None
    per_row = [[] for _ in range(m)]
    for c0 in range(0, N, chunk_cols):
        c1 = min(N, c0 + chunk_cols)
        o = orbital_block(c0, c1)
        oF = o.reshape(-1, order="F")                        # column-major inside the chunk
        C[c0 * N:c1 * N] = cvals[oF]
        if keep_orbitals:
            orb_full[:, c0:c1] = o
        rowid = sel[oF]
        nz = np.flatnonzero(rowid >= 0)
        rid = rowid[nz]
        oo = np.argsort(rid, kind="stable")
        nz, rid = nz[oo], rid[oo]
        bounds = np.searchsorted(rid, np.arange(m + 1))
        for k in range(m):
            if bounds[k + 1] > bounds[k]:
                per_row[k].append(nz[bounds[k]:bounds[k + 1]].astype(np.int64) + c0 * N)
    cols = [np.concatenate(r) if r else np.zeros(0, dtype=np.int64) for r in per_row]
    counts = np.array([c.size for c in cols], dtype=np.int64)
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    allcols = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
    A = sp.csr_matrix((np.ones(allcols.size), allcols, indptr), shape=(m, N * N))
    b = np.zeros(m)
    b[0] = 1.0
    if not sparse:
        A = A.toarray()
    mult = sorted(int(np.prod([comb(bits, j) for j in js]))
                  for js in itertools.product(range(bits + 1), repeat=fields))
    return SDPProblem(name=f"synthetic-{fields}xH({bits},2)-m{m}", C=C, A=A, b=b, n=N,
                      expected_dim=norb, expected_blocks=[1] * norb, expected_mult=mult,
                      meta={"orbitals": orb_full})


# ----------------------------------------------------------------------------
# Jordan (not coherent) partitions: symmetrised regular representation of a non-abelian group
# ----------------------------------------------------------------------------
def symmetrized_group_partition(kind: str = "S3") -> np.ndarray:
    """Label matrix (1-based classes, N x N, symmetric) of the partition  L[x, y] = class of {x^-1 y, y^-1 x}
    for a small non-abelian group acting on itself.  Its span is the Jordan algebra of SYMMETRIC elements of
    the regular representation: closed under squaring, but NOT under products (it is not a coherent
    configuration), so the orbit of a unit vector under the basis is not an invariant subspace -- the case
    that exercises the closure step of the module path (csrc/krylov.cu).  kinds: "S3", "S4", "D<n>" (dihedral
    of order 2n), "Q8"."""
    if kind == "S3" or kind == "S4":
        k = int(kind[1])
        elems = list(itertools.permutations(range(k)))
        mul = lambda a, b: tuple(a[b[i]] for i in range(k))
        inv = lambda a: tuple(sorted(range(k), key=lambda i: a[i]))
    elif kind.startswith("D"):
        n = int(kind[1:])
        elems = [(r, s) for s in (0, 1) for r in range(n)]          # r^a s^b
        mul = lambda a, b: ((a[0] + (b[0] if a[1] == 0 else -b[0])) % n, (a[1] + b[1]) % 2)
        inv = lambda a: ((-a[0]) % n, 0) if a[1] == 0 else a
    elif kind == "Q8":
        # quaternion units as (sign, axis): axis 0 = 1, 1 = i, 2 = j, 3 = k
        elems = [(s, a) for s in (1, -1) for a in range(4)]
        tab = {(1, 2): (1, 3), (2, 3): (1, 1), (3, 1): (1, 2), (2, 1): (-1, 3), (3, 2): (-1, 1), (1, 3): (-1, 2)}

        def mul(a, b):
            if a[1] == 0:
                return (a[0] * b[0], b[1])
            if b[1] == 0:
                return (a[0] * b[0], a[1])
            if a[1] == b[1]:
                return (-a[0] * b[0], 0)
            s, ax = tab[(a[1], b[1])]
            return (a[0] * b[0] * s, ax)
        inv = lambda a: a if a[1] == 0 else (-a[0], a[1])
    else:
        raise ValueError(kind)
    index = {g: i for i, g in enumerate(elems)}
    N = len(elems)
    cls = {}
    L = np.zeros((N, N), dtype=np.int64)
    for y in range(N):                      # column-major first occurrence: columns outer
        for x in range(N):
            g = mul(inv(elems[x]), elems[y])
            key = min(index[g], index[inv(g)])
            L[x, y] = cls.setdefault(key, len(cls) + 1)
    assert np.array_equal(L, L.T)
    return L
