"""CPU oracle: a restatement of the reference's Jordan-reduction hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import, call, link or execute anything in this directory.
The product (``sdpsymmetryreduction.jl_b200``) never imports it and has no CPU
fallback: it raises if ``libsdpsr_cuda.so`` is missing.

The reference (DanielBrosch/SDPSymmetryReduction.jl v0.2.1) is pure Julia and
neither this container nor the GPU box has a Julia runtime, so the reference
itself cannot be executed (``oracle/_ref`` does not exist).  The oracle is pinned
instead against every known-answer test the reference's own test-suite holds
for this path (see ``tests/test_oracle_pins.py``):

  * test/runtests.jl:11-27,40        label / refinement / desymmetrize identities
  * test/lovasz.jl:6,8,22,24,38,40   ER(3/5/7): dims 12/15/18, blocks [2,2,3]...
  * test/qap.jl:20,23                esc16j: dim 150, blocks 10x[1] + 5x[7]
  * test/numerical_issues.jl:91-94   64x64 / 1312-class fixture never fails
  * test/runtests.jl:43-57           complex path block sizes, real path throws

Parity status: PINNED for partitions (labels, dims), block sizes and
multiplicities.  Block *values* are not pinned by the reference at all; they
are pinned here by the uniqueness argument of SURVEY.md A.6 plus closed forms
(Krawtchouk / Eberlein eigenmatrices) and the spectrum invariant.  The bits of
the third-party pieces the reference delegates to (SPQR ``qr``, ``Krylov.craig``,
OpenBLAS, LAPACK) are "parity unpinned" -- they cannot be reproduced without
Julia; the partition is insensitive to them except at the initial step, which
is guarded by the documented ``snap_decimals`` safeguard.

Every function takes its random coefficients from an explicit ``rand(n)``
callable so that the oracle and the GPU path consume identical values in the
reference's draw order.
"""
from .jordan import (  # noqa: F401
    Partition,
    admissible_subspace,
    admissible_subspace_trace,
    clamp_round,
    clamptol,
    desymmetrize,
    fill,
    init_elements,
    partition_from_values,
    randomize,
    refine,
    sort_unique,
    unsafe_round,
)
from .blockdiag import (  # noqa: F401
    DimensionMismatch,
    InvalidDecompositionField,
    NumericalInconsistency,
    basis_image,
    blockDiagonalize,
    check_block_sizes,
    diagonalize,
    eigen_clusters,
    eigen_decomposition,
    irreducible_decomposition,
    isomorphism_partition,
    otsu_threshold,
)
