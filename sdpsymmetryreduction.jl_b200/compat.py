"""The legacy names of the reference (src/compat.jl:1-21), for callers written against the pre-0.2 API.

``part`` / ``coarsestPart`` / ``rndPart`` are the device-backed ``Partition`` operations of ``api.py``.  The three
small dense helpers (``roundMat``, ``orthProject``, ``projectAndRound``) are host utilities in the reference too
(``qr(A) \\ v`` on a caller-supplied dense matrix; they are not on the path ``admissible_subspace`` takes -- that
projection runs on the device, csrc/project.cu) and stay host utilities here."""
from __future__ import annotations

import numpy as np

from . import api


def roundToZero(f, atol: float = api.RTOL_DEFAULT):
    """``clamptol`` (src/utils.jl:14-32, src/compat.jl:1-2)."""
    a = np.asarray(f, dtype=np.float64)
    out = np.where(np.abs(a) < atol, 0.0, a)
    return float(out) if out.ndim == 0 else out


def part(M) -> api.Partition:                                   # src/compat.jl:11
    return api.Partition(M)


def coarsestPart(P: api.Partition, Q: api.Partition) -> api.Partition:   # src/compat.jl:12
    return api.refine(P, Q)


def rndPart(P: api.Partition, rand=None) -> np.ndarray:         # src/compat.jl:13
    return api.randomize(P, rand)


def roundMat(M: np.ndarray) -> np.ndarray:
    """``M .= clamptol.(round.(M, sigdigits=5))`` in place (src/compat.jl:15)."""
    M = np.asarray(M)
    with np.errstate(divide="ignore", invalid="ignore"):
        mag = np.where(M == 0, 1.0, 10.0 ** (4 - np.floor(np.log10(np.abs(np.where(M == 0, 1.0, M))))))
    M[...] = roundToZero(np.round(M * mag) / mag)
    return M


def orthProject(A: np.ndarray, v: np.ndarray) -> np.ndarray:
    """Orthogonal projection of v onto the column space of A: ``A * (qr(A) \\ v)`` (src/compat.jl:4-6,
    src/utils.jl:59-69)."""
    A = np.asarray(A, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    x, *_ = np.linalg.lstsq(A, v, rcond=None)
    return A @ x


def projectAndRound(M: np.ndarray, A: np.ndarray, round: bool = True) -> np.ndarray:   # noqa: A002 (reference keyword)
    """src/compat.jl:17-24: ``v = vec(M); v .-= orthProject(A, v); round && roundMat(v); reshape``."""
    M = np.asarray(M, dtype=np.float64)
    v = M.reshape(-1, order="F").copy()
    v -= orthProject(A, v)
    if round:
        roundMat(v)
    return v.reshape(M.shape, order="F")
