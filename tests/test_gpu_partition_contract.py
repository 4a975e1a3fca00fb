"""The `AbstractPartition` contract (src/abstract_part.jl:1-18) on the device back-end: a partition class that offers
only what the reference asks of a back-end -- constructor from a matrix of numbers, dim, size, fill!, refine! -- is
handed to the contract-only driver of tests/test_partitions_set_cpu.py (src/partitions.jl:109-190 restated against the
contract), exactly as test/partitions_set.jl:102-105 hands the reference's drivers a second back-end.  Every partition
operation runs through the C ABI; the labels must equal the oracle's bit for bit."""
import numpy as np
import pytest

import oracle as O
from sdpsr_b200 import binding as B
from sdpsr_b200 import problems as pr

from conftest import Coeffs
from test_partitions_set_cpu import generic_admissible_subspace

pytestmark = pytest.mark.gpu
ATOL = O.jordan.RTOL_DEFAULT


class DevicePartition:
    """What integration/julia/SDPSRCuda.jl's CuPartition provides, in the host language of this image."""

    def __init__(self, M):                                    # (P::AbstractPartition)(M::AbstractMatrix)
        M = np.asfortranarray(np.asarray(M, dtype=np.float64))
        self.size = M.shape
        self.ctx = B.Context(M.shape[0])
        self.d = self.ctx.refine_values(M, ATOL, do_round=False)        # on the empty partition: Partition(M)

    def dim(self):
        return self.d

    def refine(self, other):                                  # refine!(p, q)
        self.d = self.ctx.refine_labels(other.matrix())
        other.close()
        return self

    def fill(self, values):                                   # fill!(M, p; values)
        self.ctx.fill(np.asarray(values, dtype=np.float64))
        return self.ctx.get_matrix(B.MAT_X)

    def matrix(self):
        return self.ctx.get_labels(np.uint32)

    def close(self):
        self.ctx.close()


@pytest.mark.parametrize("q,expect_dim", [(3, 12), (5, 15)])
def test_contract_only_driver_on_the_device_backend(q, expect_dim):
    prob = pr.lovasz_er(q)
    P = generic_admissible_subspace(DevicePartition, *prob, Coeffs(21))
    Po = O.admissible_subspace(*prob, Coeffs(21))
    assert P.dim() == Po.nparts == expect_dim                 # test/partitions_set.jl:106, test/lovasz.jl:6,22
    assert np.array_equal(P.matrix(), Po.matrix)
    P.close()
