"""test/partitions_set.jl transliterated (CPU only): the reference runs its drivers on a SECOND back-end of the
`AbstractPartition` contract (`PartitionSet.Partition{BitSet}`: a partition kept as sets of linear indices, classes
numbered in the order refine! creates them) and pins the same dims and block sizes as for the label-matrix back-end
(:106-108, :129-132).  Here that back-end is restated on the host, the generic driver of src/partitions.jl:109-190 is
written against the contract only (constructor from a matrix of numbers, dim, size, fill!, refine!), and the result
is held against the oracle's label-matrix partition: same dims, same block sizes, and -- stronger than the reference
asserts -- the same classes up to their numbering.  This is the contract the GPU back-end plugs into
(src/abstract_part.jl:1-18; integration/julia/SDPSRCuda.jl supplies the same methods plus _constraints / deepcopy)."""
import os

import numpy as np
import pytest

import oracle as O
from oracle import blockdiag as OB
from oracle import jordan as OJ
from sdpsr_b200 import problems as pr

from conftest import GOLDEN, Coeffs


class PartitionSet:
    """test/partitions_set.jl:6-93.  `sets[k]` = column-major linear indices (0-based) of class k+1, `zero_set` the
    indices outside every class.  An `owner` array (index -> position in `sets`, -1 for the zero set) stands in for the
    reference's `findfirst(S -> s in S, P.sets)` (:49) -- same answer, O(1) instead of O(dim)."""

    def __init__(self, M):                                   # Partition{S}(M)  :16-34
        M = np.asarray(M)
        self.size = M.shape
        flat = M.reshape(-1, order="F")
        keys = flat.view(np.uint64) if flat.dtype == np.float64 else flat
        zero = np.uint64(0) if flat.dtype == np.float64 else 0
        d = {zero: 0}                                        # zero(eltype(M)) => 0
        sets = [set()]
        for idx, v in enumerate(keys.tolist()):
            k = d.setdefault(v, len(sets))
            if k == len(sets):
                sets.append({idx})
            else:
                sets[k].add(idx)
        self.zero_set = sets.pop(0)
        self.sets = sets
        self._reindex()

    def _reindex(self):
        self.owner = np.full(self.size[0] * self.size[1], -1, dtype=np.int64)
        for k, s in enumerate(self.sets):
            self.owner[list(s)] = k

    def dim(self):                                           # :12
        return len(self.sets)

    def refine(self, other):                                 # refine!(P1, P2)  :36-68
        assert self.size == other.size
        P2_sets = [set(s) for s in other.sets]
        z2 = set(other.zero_set)
        if self.zero_set != z2:                              # :39-44
            P2_sets.append(self.zero_set - z2)
            self.zero_set &= z2
            z2 -= self.zero_set
            P2_sets.append(z2)
        for S2 in P2_sets:                                   # :45-66
            splits = {}
            for s2 in S2:
                idx = int(self.owner[s2])
                src = self.zero_set if idx < 0 else self.sets[idx]
                src.discard(s2)
                splits.setdefault(idx, set()).add(s2)
            for i in list(splits):                           # an emptied class takes its split back (:59-64)
                if i >= 0 and not self.sets[i]:
                    self.sets[i] |= splits.pop(i)
            for piece in splits.values():                    # the others become new classes at the end (:65)
                self.sets.append(piece)
                self.owner[list(piece)] = len(self.sets) - 1
        # (a piece split off the zero set by a non-zero class of P2 has key -1 above and is appended like the rest)
        return self

    def fill(self, values):                                  # fill!(M, P; values)  :70-81
        assert len(values) == len(self.sets)
        out = np.zeros(self.size[0] * self.size[1])
        for s, v in zip(self.sets, values):
            out[list(s)] = v
        return out.reshape(self.size, order="F")

    def matrix(self):                                        # __matrix  :83-91
        lab = np.zeros(self.size[0] * self.size[1], dtype=np.int64)
        for k, s in enumerate(self.sets):
            lab[list(s)] = k + 1
        return lab.reshape(self.size, order="F")

    def constraints(self):                                   # SR._constraints(P) = collect.(P.sets)  :93
        return [np.array(sorted(s), dtype=np.int64) for s in self.sets]


def generic_admissible_subspace(Part, C, A, b, rand, atol=OJ.RTOL_DEFAULT):
    """src/partitions.jl:109-190 against the AbstractPartition contract only."""
    CL, X0, proj = OJ.init_elements(C, A, b, atol, 12)
    n = CL.shape[0]
    S = Part(CL)                                             # :145
    S = S.refine(Part(X0))                                   # :146
    maxdim = (n * n + n) // 2
    cur = S.dim()
    while cur < maxdim:                                      # :154
        X = S.fill(rand(S.dim()))                            # :159
        x = X.reshape(-1, order="F")
        x = OJ.clamp_round(x - proj(x), atol)                # :160-162
        S = S.refine(Part(x.reshape(n, n, order="F")))       # :164
        if cur != S.dim():                                   # :166-168
            X = S.fill(rand(S.dim()))
        else:
            X = x.reshape(n, n, order="F")
        S = S.refine(Part(OJ.clamp_round(X @ X, atol)))      # :172-174
        if cur == S.dim():                                   # :180-182
            break
        cur = S.dim()
    return S


def same_classes(lab_a, lab_b):
    """equal as partitions (zero set included), whatever the numbering"""
    a, b = lab_a.reshape(-1), lab_b.reshape(-1)
    if not np.array_equal(a == 0, b == 0):
        return False
    pairs = np.unique(np.stack([a, b]), axis=1)
    return pairs.shape[1] == np.unique(a).size == np.unique(b).size


CASES = [
    (lambda: pr.lovasz_er(3), 12, [2, 2, 3]),                                              # :98-108
    (lambda: pr.qap_esc16j(os.path.join(GOLDEN, "esc16j.npz")), 150, [1] * 10 + [7] * 5),   # :119-132
]


@pytest.mark.parametrize("make,expect_dim,expect_sizes", CASES, ids=["ER(3)", "esc16j"])
def test_partition_set_backend(make, expect_dim, expect_sizes):
    prob = make()
    P = generic_admissible_subspace(PartitionSet, *prob, Coeffs(21))
    assert P.dim() == expect_dim                                                          # :106 / :129
    # the label-matrix back-end finds the same classes (its numbering is by first occurrence, the set back-end's by
    # creation order)
    Po = O.admissible_subspace(*prob, Coeffs(22))
    assert Po.nparts == expect_dim and same_classes(P.matrix(), Po.matrix)
    # _constraints describes the same classes again
    cons = P.constraints()
    assert len(cons) == expect_dim and sum(c.size for c in cons) + len(P.zero_set) == prob.n ** 2
    lab = np.zeros(prob.n ** 2, dtype=np.int64)
    for k, c in enumerate(cons):
        lab[c] = k + 1
    assert np.array_equal(lab.reshape(prob.n, prob.n, order="F"), P.matrix())
    # diagonalize(Float64, P) through the generic path: block sizes (:107-108 / :131-132)
    Qhat = OB.diagonalize(OJ.Partition(P.dim(), P.matrix()), Coeffs(23))
    assert sorted(q.shape[1] for q in Qhat) == sorted(expect_sizes)
    OB.check_block_sizes(Qhat, OJ.Partition(P.dim(), P.matrix()))


def test_partition_set_semantics_match_the_label_matrix_backend():
    """Constructor, refine! (zero sets included) and fill! of the set back-end against the oracle's label matrices on
    random small matrices with zeros."""
    rng = np.random.default_rng(5)
    for trial in range(20):
        n = int(rng.integers(2, 7))
        M1 = rng.integers(0, 4, size=(n, n)).astype(np.float64)
        M2 = rng.integers(0, 3, size=(n, n)).astype(np.float64)
        A, B = PartitionSet(M1), PartitionSet(M2)
        Ao, Bo = OJ.partition_from_values(M1), OJ.partition_from_values(M2)
        assert A.dim() == Ao.nparts and np.array_equal(A.matrix(), Ao.matrix)            # both number by first occurrence
        R = A.refine(B)
        Ro = OJ.refine(Ao, Bo)
        assert R.dim() == Ro.nparts and same_classes(R.matrix(), Ro.matrix)
        vals = rng.random(R.dim())
        F = R.fill(vals)
        assert np.array_equal(F == 0.0, R.matrix() == 0)
        for k in range(R.dim()):
            assert np.all(F[R.matrix() == k + 1] == vals[k])
