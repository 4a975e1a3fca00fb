"""A host-side MODEL of the tile deal of the sharded products, used by the CPU tests only.

The deal itself is compiled into libsdpsr_cuda.so (``owned_tiles`` in csrc/gemm_f64.cu, ``build_tiles`` in
csrc/gemm_i8.cu) and exported host-side as ``sdpsr_debug_tile_deal``; tests/test_sharding_gloo.py holds this model
against it tile for tile, then proves coverage and balance on the model.  Tile-COLUMNS of a product are dealt
round-robin to the ranks (``tn % nranks == rank``); for a symmetric product only tiles on or below the diagonal
are computed and the round-robin deal balances the triangle.  Nothing in the product imports this module; the
sharding of the partition itself (column blocks + key-table merge) lives in csrc/shard.cu.
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


# ---- INT8 symmetric square (csrc/gemm_i8.cu) ------------------------------------------------------
I8_TILE_M = 128      # tile rows    (UMMA M)
I8_TILE_N = 256      # tile columns (UMMA N)
I8_GROUP = 12        # tile rows walked together, so that the tiles in flight share operand panels


def owner_of_tile_column_snake(tn: int, nranks: int) -> int:
    """Snake deal of the INT8 square (csrc/sdpsr_internal.cuh::sdpsr_tilecol_owner_snake): 0..G-1, G-1..0, ...  Column
    tn of the lower triangle holds fewer tiles the larger tn is; two consecutive rounds pair up to equal weight."""
    q, p = divmod(tn, nranks)
    return nranks - 1 - p if q & 1 else p


def owned_tiles_i8(n: int, nranks: int, rank: int):
    """(tm, tn) tiles of the lower triangle this rank computes -- same order as
    csrc/gemm_i8.cu::build_tiles: 256-column tile-columns dealt in snake order, and a tile is kept when
    it holds at least one entry on or below the diagonal."""
    tiles_m = (n + I8_TILE_M - 1) // I8_TILE_M
    tiles_n = (n + I8_TILE_N - 1) // I8_TILE_N
    out = []
    for g0 in range(0, tiles_m, I8_GROUP):
        g1 = min(tiles_m, g0 + I8_GROUP)
        for tn in range(tiles_n):
            if owner_of_tile_column_snake(tn, nranks) != rank:
                continue
            for tm in range(g0, g1):
                if (tm + 1) * I8_TILE_M - 1 >= tn * I8_TILE_N:
                    out.append((tm, tn))
    return out


def i8_product_path(c0: int):
    """The operand-tile stream and the products of one accumulator pair (c0, c0+1) of the INT8 square,
    as csrc/gemm_i8.cu walks them for one k-block: tiles T_0 = B_{c0+1}, T_1 = A_0, T_2 = B_{c0}, T_3 = A_1,
    ..., and product j multiplies the consecutive tiles (T_j, T_{j+1}).  Returns (tiles, products) with
    tiles = [("B"|"A", slice)] and products = [(s, t, accumulator)] (accumulator 0 holds c0, 1 holds c0+1).
    c0 = -1 (odd number of slices) degenerates to the single product (A_0, B_0)."""
    nch = c0 + 2
    tiles = []
    for i in range(nch):
        tiles.append(("B", c0 + 1 - i))
        tiles.append(("A", i))
    products = []
    for j in range(2 * c0 + 3):
        ta, tb = tiles[j], tiles[j + 1]
        a, b = (ta, tb) if ta[0] == "A" else (tb, ta)
        assert a[0] == "A" and b[0] == "B"
        products.append((a[1], b[1], 0 if j & 1 else 1))
    return tiles, products


def i8_schedule(slices: int):
    """All products of one output tile and k-block: accumulator pairs from the smallest weight up."""
    out = []
    c0 = slices - 2
    while c0 >= -1:
        out.append((c0, *i8_product_path(c0)))
        c0 -= 2
    return out
