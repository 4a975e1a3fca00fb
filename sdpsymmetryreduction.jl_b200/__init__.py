"""sdpsr_b200 -- B200-native engine for the Jordan-reduction hot path of
SDPSymmetryReduction.jl (``admissible_subspace`` -> ``Partition``;
``blockDiagonalize`` -> ``(blkSizes, blks)``).

The compute lives in ``csrc/`` (hand-written sm_100a CUDA behind the C ABI of
``include/sdpsr.h``, built into ``libsdpsr_cuda.so``).  This Python package is the
host-side mirror of the reference's Julia API for that path; it has no CPU
fallback and raises ``LibraryNotBuilt`` when the shared library is missing.
"""
from . import compat, problems  # noqa: F401
from .api import (  # noqa: F401
    DimensionMismatch,
    InvalidDecompositionField,
    NumericalInconsistency,
    Partition,
    admissible_subspace,
    basis_image,
    blockDiagonalize,
    clear_context_pool,
    desymmetrize,
    diagonalize,
    dim,
    randomize,
    reduce_problem,
    refine,
    run_local_ranks,
    unSymmetrize,
)
from .binding import LibraryNotBuilt, SdpsrError, lib_path, load_library  # noqa: F401

__version__ = "0.1.0"
