"""End-to-end CPU port of the reference's hot path built from the FAST restatements (TEST
INFRASTRUCTURE -- see oracle/__init__.py for who may import this).

Same algorithm, same draw order and -- checked in tests/test_oracle_cref.py -- the same labels as
``oracle.jordan.admissible_subspace``, but the partition primitives run in ``partition_ref.c`` (one
thread, an open-addressing table: what Julia's ``Dict`` pass costs) instead of numpy's sort-based
``np.unique``, and the dense products go to the BLAS numpy links (OpenBLAS, all threads: what Julia's
``mul!`` / ``eigen`` cost).  This is the closest stand-in for the Julia reference this image can run;
``bench.py --impl reference`` times it end to end to calibrate its bounded-sample cost model.

Reference sites: src/partitions.jl:109-190 (loop), :24-66 (partition steps), src/utils.jl:34-53.
"""
from __future__ import annotations

import time
from typing import Callable, Optional

import numpy as np

from . import cref
from .jordan import RTOL_DEFAULT, Partition, init_elements


def admissible_subspace_fast(C, A, b, rand: Callable[[int], np.ndarray], atol: float = RTOL_DEFAULT,
                             snap_decimals: Optional[int] = 12, phases: Optional[dict] = None) -> Partition:
    """src/partitions.jl:109-190 with per-phase wall times accumulated into ``phases``."""
    ph = phases if phases is not None else {}
    for k in ("init", "fill", "project", "refine", "gemm"):
        ph.setdefault(k, 0.0)
    t0 = time.perf_counter()
    CL, X0, proj = init_elements(C, A, b, atol, snap_decimals)                 # :124-142
    n = CL.shape[0]
    d, lab32 = cref.part_from_values(CL)                                      # S = Part(CL)          :145
    labels = lab32.astype(np.uint64)
    d2, p2 = cref.part_from_values(X0)
    d, labels = cref.refine(labels, d, p2)                                    # refine!(S, Part(X0))  :146
    ph["init"] += time.perf_counter() - t0
    maxdim = (n * n + n) // 2
    cur = d
    iters = []
    while cur < maxdim:                                                       # :154
        t0 = time.perf_counter()
        lut = np.concatenate([[0.0], np.asarray(rand(cur), dtype=np.float64)])
        X = lut[labels]                                                       # randomize!(X, S)      :159
        ph["fill"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        x = X.reshape(-1, order="F")
        x = x - proj(x)                                                       # :160-161
        ph["project"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        X = cref.clamp_round(x.reshape(n, n, order="F"), atol)                # :162
        dX, pX = cref.part_from_values(X)
        d_proj, labels = cref.refine(labels, cur, pX)                         # :164
        ph["refine"] += time.perf_counter() - t0
        if d_proj != cur:                                                     # :166-168
            t0 = time.perf_counter()
            lut = np.concatenate([[0.0], np.asarray(rand(d_proj), dtype=np.float64)])
            X = lut[labels]
            ph["fill"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        X2 = X @ X                                                            # mul!(X2, X, X)        :172
        ph["gemm"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        d_sq, labels = cref.round_refine(labels, d_proj, X2, atol)            # :173-174
        ph["refine"] += time.perf_counter() - t0
        iters.append((d_proj, d_sq))
        if cur == d_sq:                                                       # :180-182
            break
        cur = d_sq
    ph["iterations"] = len(iters)
    ph["iters"] = iters
    return Partition(int(cur if not iters else iters[-1][1]), labels.astype(np.int64))
