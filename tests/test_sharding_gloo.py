"""CPU tests of the multi-GPU host logic (world_size 2, gloo): the tile deal covers every tile
exactly once and is balanced, and the exchange protocol (owner broadcasts each tile-column,
everyone mirrors the lower triangle) reassembles the full product on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdpsr_b200 import binding as B
from sdpsr_b200 import sharding as sh


@pytest.mark.parametrize("n", [1, 100, 257, 4096, 15504, 16384, 32768])
@pytest.mark.parametrize("nranks", [1, 2, 3, 8])
def test_python_model_equals_the_library_deal(n, nranks):
    """sharding.py is only a MODEL; the deal the kernels use is compiled into libsdpsr_cuda.so and exported
    host-side (sdpsr_debug_tile_deal).  Both must agree tile for tile, in order -- so the coverage / balance
    properties asserted on the model below are properties of the product code."""
    t = sh.num_tiles(n)
    for r in range(nranks):
        assert B.tile_deal(0, n, nranks, r) == sh.owned_tiles(t, t, False, nranks, r)
        assert B.tile_deal(1, n, nranks, r) == sh.owned_tiles(t, t, True, nranks, r)
        assert B.tile_deal(2, n, nranks, r) == sh.owned_tiles_i8(n, nranks, r)
    # CTA-pair deal (256 x 256 tiles): every lower-triangle tile exactly once
    t2 = -(-n // 256)
    seen = set()
    for r in range(nranks):
        for tm, tn in B.tile_deal(3, n, nranks, r):
            assert (tm, tn) not in seen and sh.owner_of_tile_column_snake(tn, nranks) == r and tm >= tn
            seen.add((tm, tn))
    assert len(seen) == t2 * (t2 + 1) // 2


@pytest.mark.parametrize("n", [100, 128, 1000, 4096, 16384, 15504])
@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
@pytest.mark.parametrize("lower", [False, True])
def test_tile_deal_partitions_the_tiles(n, nranks, lower):
    t = sh.num_tiles(n)
    seen = set()
    for r in range(nranks):
        for tile in sh.owned_tiles(t, t, lower, nranks, r):
            assert tile not in seen
            assert sh.owner_of_tile_column(tile[1], nranks) == r
            seen.add(tile)
    want = {(tm, tn) for tn in range(t) for tm in range(tn if lower else 0, t)}
    assert seen == want
    counts = sh.work_balance(n, lower, nranks)
    if t >= 8 * nranks:                       # enough tile-columns: within 15 % of perfect balance
        assert max(counts) <= 1.15 * sum(counts) / nranks + 1


@pytest.mark.parametrize("n", [1, 100, 257, 4096, 15504, 16384, 32768])
@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
def test_int8_tile_deal_covers_the_lower_triangle(n, nranks):
    """csrc/gemm_i8.cu::build_tiles: every entry on or below the diagonal lies in exactly one owned tile,
    the owner of a tile is the (snake-order) owner of its 256-column tile-column, and the deal is balanced."""
    seen = {}
    for r in range(nranks):
        for tm, tn in sh.owned_tiles_i8(n, nranks, r):
            assert (tm, tn) not in seen and sh.owner_of_tile_column_snake(tn, nranks) == r
            seen[(tm, tn)] = r
    for i, j in [(0, 0), (n - 1, 0), (n - 1, n - 1), (n // 2, n // 3), (min(n - 1, 255), min(n - 1, 255)),
                 (min(n - 1, 256), min(n - 1, 255)), (min(n - 1, 127), min(n - 1, 100))]:
        if i >= j:
            assert (i // sh.I8_TILE_M, j // sh.I8_TILE_N) in seen
    for tm, tn in seen:                                   # no tile lies wholly above the diagonal
        assert (tm + 1) * sh.I8_TILE_M - 1 >= tn * sh.I8_TILE_N
    tiles_m, tiles_n = -(-n // sh.I8_TILE_M), -(-n // sh.I8_TILE_N)
    want = sum(1 for tn in range(tiles_n) for tm in range(tiles_m) if (tm + 1) * sh.I8_TILE_M - 1 >= tn * sh.I8_TILE_N)
    assert len(seen) == want
    counts = [sum(1 for v in seen.values() if v == r) for r in range(nranks)]
    if tiles_n >= 8 * nranks:
        assert max(counts) <= 1.15 * sum(counts) / nranks + 1
    if tiles_n % (2 * nranks) == 0:       # whole snake rounds: every rank holds exactly the same number of tiles
        assert max(counts) == min(counts)


@pytest.mark.parametrize("n,nranks,grid", [(16384, 1, 148), (16384, 2, 148), (16384, 4, 148), (16384, 8, 148), (15504, 1, 148),
                                            (15504, 3, 132), (4096, 1, 148), (2048, 1, 5), (2048, 2, 7), (1000, 1, 3),
                                            (384, 1, 2), (200, 1, 3), (130, 1, 148)])
def test_int8_work_list_covers_every_k_block_once(n, nranks, grid):
    """csrc/gemm_i8.cu::build_schedule (exported as sdpsr_debug_i8_schedule): whole tiles in full waves, then the tiles
    of a last wave that is at most half full cut along K.  Every k-block of every owned tile is covered exactly once, the parts of a split tile are
    numbered 0..nparts-1 with consecutive scratch slots, unsplit tiles carry no slot, no two tiles share a slot or a
    semaphore, and no CTA carries more than one k-block above the ideal share."""
    for rank in range(nranks):
        items, info = B.i8_schedule(n, nranks, rank, grid)
        tiles = B.tile_deal(2, n, nranks, rank)
        KB, g = info["kblocks"], min(grid, max(1, len(tiles)))
        assert KB == -(-n // 128)
        cover, load = {}, np.zeros(g)
        for i, row in enumerate(items):
            tm, tn, k0, k1, slot, part, nparts, sem = (int(x) for x in row)
            if k1 <= k0:
                assert i >= info["nmain"]                     # padding only in the tail
                continue
            if i < info["nmain"]:
                assert (k0, k1, slot, nparts) == (0, KB, -1, 1)
            cover.setdefault((tm, tn), []).append((k0, k1, slot, part, nparts, sem))
            load[i % g] += k1 - k0
        assert sorted(cover) == sorted(tiles)
        slots, sems = set(), {}
        for t, parts in cover.items():
            parts.sort()
            assert parts[0][0] == 0 and parts[-1][1] == KB
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert all(p[4] == len(parts) for p in parts)
            if len(parts) == 1:
                assert parts[0][2] == -1
                continue
            assert sorted(p[3] for p in parts) == list(range(len(parts)))
            base = {p[2] - p[3] for p in parts}
            assert len(base) == 1 and min(base) >= 0
            for p in parts:
                assert p[2] not in slots and p[2] < info["nslots"]
                slots.add(p[2])
            assert len({p[5] for p in parts}) == 1 and parts[0][5] < info["nsems"]
            assert sems.setdefault(parts[0][5], t) == t
        assert len(slots) == info["nslots"]
        r = len(tiles) % g
        if len(tiles) >= g and KB >= 2 and 0 < 2 * r <= g:      # split: equal runs of at least a quarter tile
            assert info["nslots"] > 0 and info["nmain"] == len(tiles) - r
            assert load.max() <= info["nmain"] / g * KB + max(-(-r * KB // g), -(-KB // 4))      # one run on top
        else:
            assert info["nslots"] == 0 and info["nmain"] == len(tiles)
        assert max(len(v) for v in cover.values()) <= 5


@pytest.mark.parametrize("slices", [2, 3, 4, 5, 6, 7, 8])
def test_int8_product_schedule_covers_every_pair_once(slices):
    """csrc/gemm_i8.cu's zig-zag schedule: every ordered pair (s, t) with s + t < S is multiplied exactly
    once into the accumulator of its weight, each product after the first of a path needs exactly one new
    operand tile, and the paths have even length (B and A tiles alternate in the shared-memory ring)."""
    seen = {}
    loads = 0
    for c0, tiles, products in sh.i8_schedule(slices):
        assert len(tiles) == 2 * (c0 + 2) and len(products) == 2 * c0 + 3
        assert [t[0] for t in tiles] == ["B", "A"] * (c0 + 2)
        assert all(0 <= sl < slices for _, sl in tiles)
        loads += len(tiles)
        for s, t, acc in products:
            assert (s, t) not in seen
            assert s + t == c0 + acc                       # accumulator 0 <-> c0, accumulator 1 <-> c0 + 1
            seen[(s, t)] = c0 + acc
    want = {(s, t) for s in range(slices) for t in range(slices) if s + t < slices}
    assert set(seen) == want
    assert len(want) == slices * (slices + 1) // 2
    # one new tile per product except the first of each path: loads = products + number of paths
    assert loads == len(want) + len(sh.i8_schedule(slices))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, tile, lower, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    L = rng.integers(0, 5, size=(n, n))
    if lower:
        L = np.minimum(L, L.T)
    X = np.array([0.0, 0.3, 0.7, 0.11, 0.9])[L]
    t = (n + tile - 1) // tile
    C = np.zeros((n, n), order="F")
    for tm, tn in sh.owned_tiles(t, t, lower, world, rank):          # this rank's share of X*X
        r0, c0 = tm * tile, tn * tile
        C[r0:r0 + tile, c0:c0 + tile] = X[r0:r0 + tile, :] @ X[:, c0:c0 + tile]
    for tn in range(t):                                              # grouped broadcasts, one per slab
        slab = torch.from_numpy(np.ascontiguousarray(C[:, tn * tile:(tn + 1) * tile]))
        dist.broadcast(slab, src=sh.owner_of_tile_column(tn, world))
        C[:, tn * tile:(tn + 1) * tile] = slab.numpy()
    if lower:                                                        # mirror_lower_kernel
        iu = np.triu_indices(n, 1)
        C[iu] = C.T[iu]
    ref = X @ X
    ok = np.allclose(C, ref, rtol=1e-13, atol=1e-13)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("lower", [False, True])
def test_exchange_protocol_gloo_world2(lower):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 150, 32, lower, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == 1


# ---------------------------------------------------------------------------------------------------------
# the sharded partition (csrc/shard.cu): local pass + key-table merge, over gloo with two ranks
# ---------------------------------------------------------------------------------------------------------
def _merge_worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    old = rng.integers(0, 4, size=(n, n))                    # current labels (canonical ids, replicated)
    val = rng.integers(0, 5, size=(n, n))                    # rounded-value codes of the refining matrix
    key = (val.astype(np.int64) << 8) | old                  # the pass's 64-bit key (0 = the zero class)
    flat = key.reshape(-1, order="F")
    # 1. local pass over this rank's column block: key -> first (column-major) index inside the block
    c0, c1 = n * rank // world, n * (rank + 1) // world
    local = {}
    for idx in range(c0 * n, c1 * n):
        k = int(flat[idx])
        if k and k not in local:
            local[k] = idx
    # 2. all-gather of the (key, first index) pairs, identical merge on every rank
    gathered = [None] * world
    dist.all_gather_object(gathered, sorted(local.items()))
    merged = {}
    for pairs in gathered:
        for k, idx in pairs:
            merged[k] = min(idx, merged.get(k, idx))
    order = sorted(merged, key=lambda k: merged[k])          # canonical numbering = rank of the first indices
    canon = {k: i + 1 for i, k in enumerate(order)}
    # 3. local relabel of the block
    mine = np.array([canon.get(int(k), 0) for k in flat[c0 * n:c1 * n]], dtype=np.int64)
    # 4. lazy gather of the blocks (only to check the result here)
    blocks = [None] * world
    dist.all_gather_object(blocks, mine)
    got = np.concatenate(blocks)
    # the reference: Partition(M) numbering by first occurrence over the WHOLE matrix (src/partitions.jl:24-35)
    want, seen = np.zeros(n * n, dtype=np.int64), {}
    for idx, k in enumerate(flat):
        if k:
            want[idx] = seen.setdefault(int(k), len(seen) + 1)
    flag = torch.tensor([1 if np.array_equal(got, want) and len(order) == len(seen) else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(flag.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 24])
def test_sharded_refine_merge_protocol_gloo_world2(n):
    """C1/C2 of SURVEY.md 8(e): per-rank key tables over column blocks, merged by smallest first index, give the
    reference's first-occurrence numbering on every rank (n = 7: ragged blocks)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_merge_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == 1
