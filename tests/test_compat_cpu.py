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
