"""test/runtests.jl:29-37 transliterated: the small dense helpers of src/compat.jl (host utilities in the reference
and here; the library is imported but no device call is made)."""
import numpy as np

import sdpsr_b200 as S


def test_round_to_zero():
    assert S.compat.roundToZero(1e-10) == 0.0                                   # test/runtests.jl:11
    assert S.compat.roundToZero(0.5) == 0.5


def test_round_mat():
    rng = np.random.default_rng(1)
    M = rng.random((10, 10))
    R = S.compat.roundMat(M.copy())
    assert np.allclose(R, M, atol=1e-4)                                         # :30
    assert np.array_equal(R, np.array([[float(f"{x:.5g}") for x in row] for row in M]))   # sigdigits = 5


def test_project_and_round():
    rng = np.random.default_rng(2)
    A = rng.random((9, 3))
    M = rng.random((3, 3))
    P = S.compat.projectAndRound(M, A, round=False)
    x, *_ = np.linalg.lstsq(A, P.reshape(-1, order="F"), rcond=None)
    assert np.abs(x).max() < 1e-10                                              # :34
    T = (M - P).reshape(-1, order="F")
    x, *_ = np.linalg.lstsq(A, T, rcond=None)
    assert np.allclose(A @ x - T, 0, atol=1e-8)                                 # :36-37
    assert np.allclose(S.compat.orthProject(A, A[:, 0]), A[:, 0])


def test_partition_matrix_waits_for_a_pending_export():
    """`Partition.matrix` is lazy only while an export started by admissible_subspace is in flight: the first access
    joins it (and surfaces its error), later accesses and plainly constructed partitions never wait (no device call
    here: the arrival is a stand-in)."""
    M = np.arange(9, dtype=np.uint16).reshape(3, 3)
    P = S.Partition(8, M)
    assert P._arrival is None and P.matrix is not None and P.shape == (3, 3)
    calls = []
    P._arrival = lambda: calls.append("joined")
    assert P.nparts == 8 and calls == []              # nothing but .matrix needs the host copy
    assert np.array_equal(P.matrix, M) and calls == ["joined"]
    assert np.array_equal(P.matrix, M) and calls == ["joined"]

    def overflow():
        raise OverflowError("InexactError: label does not fit uint8")
    Q = S.Partition(8, M)
    Q._arrival = overflow
    try:
        Q.matrix
    except OverflowError:
        pass
    else:
        raise AssertionError("the export's error must surface at the first access")
    assert Q._arrival is None                          # reported once
    Q.matrix = M + 1                                   # assigning a matrix cancels any pending arrival
    assert np.array_equal(Q.matrix, M + 1)
