"""Pins the CPU oracle against every known-answer test the reference's own
test-suite holds for the hot path.  CPU only."""
import json
import os

import numpy as np
import pytest

import oracle as O
from sdpsr_b200 import problems as pr

from conftest import GOLDEN, Coeffs


@pytest.fixture(scope="module")
def vec():
    with open(os.path.join(GOLDEN, "runtests_vectors.json")) as fh:
        return json.load(fh)


# --- test/runtests.jl:11 ----------------------------------------------------------
def test_round_to_zero():
    assert O.clamptol(np.array([1e-10]))[0] == 0.0


# --- test/runtests.jl:13-20 -------------------------------------------------------
@pytest.mark.parametrize("seed", range(5))
def test_partition_ctor_dims(seed):
    rng = np.random.default_rng(seed)
    M = rng.integers(1, 11, size=(10, 10))
    M[0, 0] = 0
    nun = len(np.unique(M))
    assert O.partition_from_values(M).nparts == nun - 1
    assert O.partition_from_values(M.astype(np.float64)).nparts == nun - 1
    M = rng.integers(1, 11, size=(10, 10))
    nun = len(np.unique(M))
    assert O.partition_from_values(M).nparts == nun
    assert O.partition_from_values(M.astype(np.float64)).nparts == nun


# --- test/runtests.jl:22-27 -------------------------------------------------------
def test_refine_identity(vec):
    P1 = O.partition_from_values(np.array(vec["P1"]))
    P2 = O.partition_from_values(np.array(vec["P2"]))
    P3 = O.partition_from_values(np.array(vec["P3_coarsest_P1_P2"]))
    assert O.refine(P1.copy(), P2) == P3
    assert np.array_equal(O.refine(P1.copy(), P2).matrix, np.array(vec["P3_coarsest_P1_P2"]))


def test_randomize_roundtrip(vec):
    P1 = O.partition_from_values(np.array(vec["P1"]))
    assert O.partition_from_values(O.randomize(P1, Coeffs(3))) == P1


def test_first_occurrence_is_column_major():
    M = np.array([[5.0, 7.0], [7.5, 5.0]])
    # column-major scan: 5.0, 7.5, 7.0, 5.0
    assert np.array_equal(O.partition_from_values(M).matrix, [[1, 3], [2, 1]])
    assert np.array_equal(O.partition_from_values(np.array([[0.0, -0.0], [1.0, 0.0]])).matrix,
                          [[0, 2], [1, 0]])   # isequal: -0.0 is not the zero key


# --- test/runtests.jl:40 ----------------------------------------------------------
def test_desymmetrize_identity(vec):
    P1 = O.partition_from_values(np.array(vec["P1"]))
    want = vec["unsymmetrize_P1"]
    got = O.desymmetrize(P1, Coeffs(7))
    assert got.nparts == want["nparts"]
    assert np.array_equal(got.matrix, np.array(want["matrix"]))


# --- test/runtests.jl:43-57 -------------------------------------------------------
def test_complex_path_sizes(vec):
    c4 = vec["circulant4"]
    P = O.Partition(c4["nparts"], np.array(c4["matrix"]))
    X = O.randomize(P, Coeffs(1))
    assert np.array_equal(X, X.T)
    sizes, _ = O.blockDiagonalize(P, Coeffs(2), complex=True)
    assert sizes == c4["complex_blkSizes"]
    P3 = O.partition_from_values(np.array(vec["C3"]["matrix"]))
    with pytest.raises(O.InvalidDecompositionField):
        O.blockDiagonalize(P3, Coeffs(3))
    sizes, _ = O.blockDiagonalize(P3, Coeffs(4), complex=True)
    assert sizes == vec["C3"]["complex_blkSizes"]


# --- rounding spec (src/utils.jl:34-53) -------------------------------
def test_rounding_is_truncation():
    a = np.array([0.123456789, -0.123456789, 1e-9, 3.0, 1 / 16, 2.0 ** -30 * 0.75])
    r = O.clamp_round(a)
    assert r[2] == 0.0
    # 0.123456789 = 0.987654312 * 2^-3 -> trunc 7 digits of the fraction
    assert r[0] == np.ldexp(9876543 / 1e7, -3) and r[1] == -r[0]
    assert r[3] == 3.0 and r[4] == 1 / 16
    assert not np.signbit(O.clamp_round(np.array([-1e-12]))[0])      # never -0.0


# --- test/lovasz.jl:6,8,22,24,38,40 and Appendix B trajectories -----------------------
@pytest.mark.parametrize("q,traj", [
    (3, [(2, 6), (6, 12), (12, 12)]),
    (5, [(2, 6), (6, 14), (14, 15), (15, 15)]),
    (7, [(2, 6), (6, 14), (14, 17), (17, 18), (18, 18)]),
])
def test_lovasz_er(q, traj):
    prob = pr.lovasz_er(q)
    for seed in (1, 2):
        c = Coeffs(seed)
        P, tr = O.admissible_subspace_trace(*prob, c)
        assert P.nparts == prob.expected_dim
        assert tr["init"] == 2 and tr["iters"] == traj
        Qhat = O.diagonalize(P, c)                         # atol = 1e-12*N, as in the test
        assert sorted(q.shape[1] for q in Qhat) == prob.expected_blocks
        sizes, blks = O.blockDiagonalize(P, c)
        assert sorted(sizes) == prob.expected_blocks
        assert len(blks) == P.nparts


def test_petersen():
    prob = pr.petersen()
    c = Coeffs(5)
    P, tr = O.admissible_subspace_trace(*prob, c)
    assert P.nparts == 3 and tr["iters"] == [(2, 3), (3, 3)]
    sizes, blks = O.blockDiagonalize(P, c)
    assert sizes == [1, 1, 1]
    # rows of the eigenmatrix of the Johnson scheme J(5,2): I, J(5,2) (intersecting), Petersen
    vals = sorted(tuple(round(float(blks[i][k][0, 0]), 9) for i in range(3)) for k in range(3))
    assert vals == sorted([(1, 6, 3), (1, 1, -2), (1, -2, 1)])


# --- test/qap.jl:20,23 ------------------------------------------------------------------
def test_qap_esc16j():
    prob = pr.qap_esc16j(os.path.join(GOLDEN, "esc16j.npz"))
    assert prob.n == 256 and prob.A.shape == (33, 65536)
    c = Coeffs(11)
    P, tr = O.admissible_subspace_trace(*prob, c)
    assert P.nparts == 150
    assert tr["init"] == 7 and tr["iters"] == [(7, 117), (117, 150), (150, 150)]
    assert (P.matrix == 0).sum() == 0
    assert list(P.matrix[:20, 0]) == [1, 2, 2, 3, 2, 3, 3, 4, 2, 3, 3, 4, 3, 4, 4, 5, 6, 7, 7, 8]
    Qhat = O.diagonalize(P, c)
    assert sorted(q.shape[1] for q in Qhat) == [1] * 10 + [7] * 5


def test_pattern_hash_init_is_not_the_reference():
    """SURVEY.md fact 6 is about the *initial partition*; here we only check the
    init really has 7 classes (not the 263 column patterns)."""
    prob = pr.qap_esc16j(os.path.join(GOLDEN, "esc16j.npz"))
    CL, X0, _ = O.init_elements(*prob)
    S = O.refine(O.partition_from_values(CL), O.partition_from_values(X0))
    assert S.nparts == 7


# --- test/numerical_issues.jl:91-94 (bounded: 300 trials instead of 10 000) ---------------
def test_numerical_issues_fixture():
    Pm = np.load(os.path.join(GOLDEN, "numerical_issues_P.npy"))
    part = O.partition_from_values(Pm)
    assert part.nparts == 1312
    c = Coeffs(99)
    for _ in range(300):
        vals, Q, ptrs, K = O.eigen_decomposition(part, c, atol=1e-7)
    roots = [K.find_root(i) for i in range(len(K))]
    sizes = sorted(np.bincount(roots)[np.unique(roots)])
    assert len(ptrs) - 1 == 64 and sizes == [16, 48]


# --- scheme graphs: closed forms (SURVEY.md 8d cfg 3) ------------------------------------
def test_hamming_closed_form():
    prob = pr.hamming(3, 4)
    c = Coeffs(3)
    P = O.admissible_subspace(*prob, c)
    assert P.nparts == 4
    D = pr.hamming_distance_matrix(3, 4)
    assert np.array_equal(P.matrix, D + 1)             # labels 1..4 <-> distance 0..3
    sizes, blks = O.blockDiagonalize(P, c)
    assert sizes == [1, 1, 1, 1]
    K = pr.krawtchouk(3, 4)
    got = np.array([[blks[i][k][0, 0] for k in range(4)] for i in range(4)])
    # column k of `got` is some column j of the eigenmatrix
    for k in range(4):
        assert min(np.abs(K - got[:, [k]]).max(axis=0)) < 1e-10


def test_spectrum_invariant():
    """eig(sum x_i B_i) == union over blocks of eig(sum x_i blks[i][k]) x multiplicity."""
    prob = pr.lovasz_er(5)
    c = Coeffs(8)
    P = O.admissible_subspace(*prob, c)
    vals, Q, ptrs, K = O.eigen_decomposition(P, Coeffs(8), atol=O.jordan.RTOL_DEFAULT)
    sizes, blks = O.blockDiagonalize(P, Coeffs(8))
    x = np.random.default_rng(0).random(P.nparts)
    big = np.linalg.eigvalsh(O.fill(P, x))
    roots = [K.find_root(i) for i in range(len(K))]
    mult = [int(ptrs[r + 1] - ptrs[r]) for r in dict.fromkeys(roots)]
    small = []
    for k, s in enumerate(sizes):
        Mk = sum(x[i] * blks[i][k] for i in range(P.nparts))
        small += list(np.linalg.eigvalsh(Mk)) * mult[k]
    assert np.allclose(sorted(small), big, atol=1e-10)
