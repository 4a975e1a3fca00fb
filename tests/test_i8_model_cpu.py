"""CPU checks of the host model of the INT8 square (tests/i8_model.py): the model the GPU tests
hold csrc/gemm_i8.cu to must itself be an FP64-grade X*X."""
import numpy as np
import pytest

from i8_model import exact_square, slices_of


@pytest.mark.parametrize("n", [1, 7, 64, 200])
def test_model_is_fp64_grade_and_symmetric(n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)) * np.exp(rng.uniform(-4, 2, size=(n, n)))
    X = np.triu(A) + np.triu(A, 1).T
    ref = X @ X
    for bits, smax in ((7, 8), (8, 7)):
        prev = None
        for S in range(smax, 1, -1):
            got = exact_square(X, S, bits)
            assert np.array_equal(got, got.T)
            err = np.abs(got - ref).max() / np.abs(ref).max()
            if S == smax:
                assert err < 2e-14
            if prev is not None:
                assert err >= prev * 0.5          # fewer digits never help
            prev = err


def test_digits_reconstruct_the_quantised_entries():
    rng = np.random.default_rng(3)
    X = rng.random((50, 50)) - 0.3
    for bits in (7, 8):
        for S in (2, 5, 7):
            D, e = slices_of(X, S, bits)
            q = sum(D[s] * 2.0 ** (bits * (S - 1 - s)) for s in range(S))
            assert np.abs(D[0]).max() <= 64 and all(np.abs(d).max() <= 2 ** (bits - 1) for d in D)
            unit = 2.0 ** (e - (bits * (S - 1) + 6))
            assert np.abs(q * unit - X).max() <= unit / 2          # half a unit of the last digit
