"""Multi-GPU sharding of the path: pure host logic, mirrored by
``owned_tiles`` in csrc/gemm_f64.cu and CPU-tested with gloo (tests/test_sharding_gloo.py).

Only the GEMMs shard (they are >98 % of the kernel time): the 128-column TILE-COLUMNS of a
product C = A*B are dealt round-robin to the ranks (``tn % nranks == rank``).  For a symmetric
product only tiles on or below the diagonal are computed; the round-robin deal balances the
triangle.  Each owned tile-column is one contiguous slab of the column-major matrix, so the
exchange is one grouped set of broadcasts, one per tile-column, rooted at its owner; the upper
triangle is then mirrored locally.  The cheap streaming passes (fill, projection, refine) are
replicated on every rank: they see bit-identical X / X^2, so every observable (dims, canonical
labels, X) is identical on all ranks with no further communication.
"""
from __future__ import annotations

TILE = 128


def num_tiles(n: int, tile: int = TILE) -> int:
    return (n + tile - 1) // tile


def owner_of_tile_column(tn: int, nranks: int) -> int:
    return tn % nranks


def owned_tiles(tiles_m: int, tiles_n: int, lower: bool, nranks: int, rank: int):
    """(tm, tn) pairs this rank computes -- same order as csrc/gemm_f64.cu::owned_tiles."""
    out = []
    for tn in range(rank, tiles_n, nranks):
        for tm in range(tn if lower else 0, tiles_m):
            out.append((tm, tn))
    return out


def work_balance(n: int, lower: bool, nranks: int):
    t = num_tiles(n)
    counts = [len(owned_tiles(t, t, lower, nranks, r)) for r in range(nranks)]
    return counts
