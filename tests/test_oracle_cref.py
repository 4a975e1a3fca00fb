"""The C restatement (oracle/partition_ref.c) must agree with the numpy oracle."""
import numpy as np
import pytest

import oracle as O
from oracle import cref


@pytest.mark.parametrize("seed", range(4))
def test_cref_matches_numpy_oracle(seed):
    rng = np.random.default_rng(seed)
    n = 60
    M = rng.integers(0, 9, size=(n, n)) * 0.173 + rng.integers(0, 2, size=(n, n)) * 1e-9
    Mr = O.clamp_round(M)
    assert np.array_equal(cref.clamp_round(M, O.jordan.RTOL_DEFAULT), Mr)
    P = O.partition_from_values(Mr)
    d, lab = cref.part_from_values(Mr)
    assert d == P.nparts and np.array_equal(lab, P.matrix)
    M2 = rng.integers(0, 5, size=(n, n)).astype(np.float64)
    P2 = O.refine(P, O.partition_from_values(O.clamp_round(M2)))
    d2, lab2 = cref.round_refine(lab, d, M2, O.jordan.RTOL_DEFAULT)
    assert d2 == P2.nparts and np.array_equal(lab2, P2.matrix)


def test_cref_runtests_identity():
    P1 = np.array([[1, 2, 2], [2, 3, 3], [2, 3, 3]])
    P2 = np.array([[1, 1, 2], [1, 1, 2], [1, 1, 3]])
    d, lab = cref.refine(P1, 3, P2)
    assert d == 6 and np.array_equal(lab, [[1, 2, 4], [2, 3, 5], [2, 3, 6]])
