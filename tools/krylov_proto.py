"""Prototype (CPU, numpy; development aid, not product code): block-diagonalization without a dense
eigendecomposition.  The generic element A1 = fill(P, r1) of a Jordan-closed partition algebra has only
ne = sum(s_k) distinct eigenvalues, so Lanczos (full reorthogonalisation) from a generic start vector
breaks down after ne steps and its Ritz vectors are exact eigenvectors, one per eigenspace.
Multiplicities = traces of the spectral projectors, read from Lanczos runs started at one unit vector per
diagonal class.  Checked here against the oracle's dense path (same coefficient vectors -> same blocks).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import oracle as O                                          # noqa: E402
from oracle import blockdiag as OB                          # noqa: E402
import sdpsr_b200.problems as pr                            # noqa: E402


def lanczos(matvec, v0, kmax, tol, anorm=None):
    """Full-reorth Lanczos.  Returns (alpha, beta, V) with breakdown when beta <= tol*anorm."""
    n = v0.size
    V = np.zeros((n, 0))
    al, be = [], []
    v = v0 / np.linalg.norm(v0)
    scale = 0.0
    for k in range(kmax):
        V = np.concatenate([V, v[:, None]], axis=1)
        w = matvec(v)
        a = float(v @ w)
        al.append(a)
        for _ in range(2):                       # CGS2 against all previous vectors
            w = w - V @ (V.T @ w)
        b = float(np.linalg.norm(w))
        scale = max(scale, float(np.sqrt(a * a + b * b + (be[-1] ** 2 if be else 0.0))))
        if b <= tol * (anorm if anorm is not None else scale):
            return np.array(al), np.array(be), V, True
        be.append(b)
        v = w / b
    return np.array(al), np.array(be), V, False


def tri_eig(al, be):
    T = np.diag(al) + np.diag(be, 1) + np.diag(be, -1)
    return np.linalg.eigh(T)


class NotApplicable(Exception):
    pass


def krylov_diagonalize(P, rand, atol, tol=1e-10, kmax=64):
    n = P.matrix.shape[0]
    lab = P.matrix
    r1 = rand(P.nparts)
    A1 = O.fill(P, r1)
    rng = np.random.default_rng(12345)
    v0 = rng.random(n) - 0.5
    al, be, V, ok = lanczos(lambda v: A1 @ v, v0, min(n, P.nparts, kmax), tol)
    if not ok:
        raise NotApplicable("no breakdown")
    th, S = tri_eig(al, be)
    ne = th.size
    anorm = float(np.abs(th).max())
    Y = V @ S                                   # one unit eigenvector per eigenspace, ascending eigenvalue
    assert np.all(np.diff(th) > atol), "clusters closer than atol"
    # multiplicities from the diagonal classes
    dlab = np.diag(lab)
    mult = np.zeros(ne)
    for c in np.unique(dlab):
        rows = np.flatnonzero(dlab == c)
        e = np.zeros(n)
        e[rows[0]] = 1.0
        a2, b2, V2, ok2 = lanczos(lambda v: A1 @ v, e, ne + 1, tol, anorm)
        if not ok2:
            raise NotApplicable("multiplicity run: no breakdown")
        t2, S2 = tri_eig(a2, b2)
        wts = S2[0, :] ** 2
        for t, w in zip(t2, wts):
            i = int(np.argmin(np.abs(th - t)))
            assert abs(th[i] - t) < 1e-8 * max(1, np.abs(th).max()), (th[i], t)
            mult[i] += w * rows.size
    m = np.rint(mult).astype(np.int64)
    assert np.abs(mult - m).max() < 1e-6 and m.sum() == n and m.min() >= 1, (mult, m.sum())
    ptrs = np.concatenate([[0], np.cumsum(m)])
    # isomorphism test
    r2 = rand(P.nparts)
    A2 = O.fill(P, r2)
    U = A2 @ Y
    G = np.zeros((ne, ne))
    for i in range(ne):
        a4, b4, V4, ok4 = lanczos(lambda v: A1 @ v, U[:, i], ne + 1, tol, anorm)
        if not ok4:
            raise NotApplicable("isomorphism run: no breakdown")
        t4, S4 = tri_eig(a4, b4)
        un = np.linalg.norm(U[:, i])
        for t, w in zip(t4, S4[0, :]):
            j = int(np.argmin(np.abs(th - t)))
            G[i, j] = max(G[i, j], abs(w) * un)
    G = np.maximum(G, G.T)
    norms = np.where(m[:, None] == m[None, :], G, 0.0)
    thr = OB.otsu_threshold(norms, atol)
    K = OB.IntDisjointSets(ne)
    for i in range(ne):
        for j in range(i + 1, ne):
            if norms[i, j] >= thr:
                K.union(i, j)
    if not OB.is_consistent(K):
        raise OB.NumericalInconsistency("inconsistent")
    kpart = [K.find_root(i) for i in range(ne)]
    roots = list(dict.fromkeys(kpart))
    r3 = rand(P.nparts)
    A3 = O.fill(P, r3)
    out = []
    for i in roots:
        Ki = [t for t in range(ne) if kpart[t] == i]
        if len(Ki) == 1:
            out.append(Y[:, i:i + 1].copy())
            continue
        u = A3 @ Y[:, i]
        a3, b3, V3, ok3 = lanczos(lambda v: A1 @ v, u, len(Ki) + 1, tol, anorm)
        if not (ok3 and a3.size == len(Ki)):
            raise NotApplicable(f"class Krylov dimension {a3.size} != {len(Ki)}")
        t3, S3 = tri_eig(a3, b3)
        R = V3 @ S3
        cols = [Y[:, i].copy()]
        for j in Ki[1:]:
            k = int(np.argmin(np.abs(t3 - th[j])))
            assert abs(t3[k] - th[j]) < 1e-8 * max(1, np.abs(th).max())
            c = R[:, k]
            if c @ u < 0:
                c = -c
            cols.append(c)
        out.append(np.stack(cols, axis=1))
    Qhat = [np.where(np.abs(q) < atol, 0.0, q) for q in out]
    return Qhat, m, th


class Coeffs:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def __call__(self, n):
        return self.rng.random(int(n))


def check(name, P, atol=O.jordan.RTOL_DEFAULT if hasattr(O, "jordan") else 1.4901161193847656e-8):
    try:
        Qk, m, th = krylov_diagonalize(P, Coeffs(5), atol)
    except NotApplicable as e:
        print(f"{name:28s} N={P.matrix.shape[0]:5d} dim={P.nparts:5d} not applicable: {e}")
        return
    Qo = O.diagonalize(P, Coeffs(5), atol=atol)
    bk = O.basis_image(Qk, P)
    bo = O.basis_image(Qo, P)
    sk, so = [q.shape[1] for q in Qk], [q.shape[1] for q in Qo]
    err = max(float(np.abs(bk[i][k] - bo[i][k]).max()) for i in range(P.nparts) for k in range(len(so))) if sk == so else None
    print(f"{name:28s} N={P.matrix.shape[0]:5d} dim={P.nparts:5d} ne={th.size:4d} sizes_equal={sk == so} "
          f"blocks={sorted(sk)[-3:]} max|blk-oracle|={err}")


if __name__ == "__main__":
    import oracle.jordan as OJ
    atol = OJ.RTOL_DEFAULT
    golden = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
    probs = [pr.petersen(), pr.lovasz_er(3), pr.lovasz_er(5), pr.lovasz_er(7), pr.kneser(8, 3), pr.hamming(3, 8),
             pr.qap_esc16j(os.path.join(golden, "esc16j.npz")), pr.synthetic_product_scheme(3, 3, 16),
             pr.synthetic_product_scheme(3, 2, 8), pr.lovasz_er(11), pr.lovasz_er(13), pr.kneser(10, 4), pr.kneser(12, 5),
             pr.hamming(5, 4)]
    for p in probs:
        P = O.admissible_subspace(*p, Coeffs(1))
        check(p.name, P, atol)
    Pm = np.load(os.path.join(golden, "numerical_issues_P.npy"))
    P = O.Partition(int(Pm.max()), Pm) if hasattr(O.Partition, "__init__") else None
    check("numerical_issues", P, 1e-7)
