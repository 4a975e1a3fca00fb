"""Multi-rank parity: the sharded path must give, on EVERY rank, the oracle's labels and blocks.

Two transports (csrc/comm.cu): the in-process one (G ranks = G host threads, all mapped onto ONE GPU --
this is what runs on a single-GPU box) and NCCL under torchrun (one process per GPU;
needs >= 2 GPUs, otherwise the same check runs in-process so that the test never skips)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as O
import sdpsr_b200 as S
from sdpsr_b200 import binding as B
from sdpsr_b200 import problems as pr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
os.environ.setdefault("SDPSR_LOCAL_BARRIER_TIMEOUT_S", "120")


class Coeffs:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def __call__(self, n):
        return self.rng.random(int(n))


def _problems():
    return [pr.lovasz_er(7), pr.qap_esc16j(os.path.join(GOLDEN, "esc16j.npz")), pr.kneser(10, 4), pr.hamming(3, 8),
            pr.synthetic_product_scheme(3, 3, 16), pr.hamming(5, 4)]


def _check_on_ranks(prob, nranks, flags, eig="auto"):
    Po = O.admissible_subspace(*prob, Coeffs(11))
    so, bo = O.blockDiagonalize(Po, Coeffs(12))

    def job(ctx, rank):
        assert ctx.comm_info() == (nranks, rank)
        tr = {}
        Pg = S.admissible_subspace(*prob, rand=Coeffs(11), ctx=ctx, trace=tr)
        bd = S.blockDiagonalize(Pg, False, rand=Coeffs(12), eig=eig)
        return Pg.nparts, Pg.matrix.copy(), list(bd.blkSizes), bd.blks, tr["iters"], ctx.timing()

    res = S.run_local_ranks(prob.n, nranks, job, flags=flags | B.F_TIMING)
    for rank, (dim, labels, sizes, blks, iters, tim) in enumerate(res):
        assert dim == Po.nparts == prob.expected_dim, (rank, dim)
        assert np.array_equal(labels, Po.matrix), f"rank {rank}: labels differ from the oracle"
        assert sizes == list(so), (rank, sizes, list(so))
        err = max(float(np.abs(blks[i][k] - bo[i][k]).max() / max(1.0, np.abs(bo[i][k]).max()))
                  for i in range(Po.nparts) for k in range(len(so)))
        assert err < 1e-8, (rank, err)
        assert iters == res[0][4]
    return res


@pytest.mark.parametrize("nranks", [2, 3, 4])
@pytest.mark.parametrize("prob", _problems(), ids=lambda p: p.name)
def test_sharded_path_on_one_device(prob, nranks):
    """Default kernel choice, G ranks sharing GPU 0."""
    _check_on_ranks(prob, nranks, 0)


@pytest.mark.parametrize("flags", [B.F_FORCE_I8, B.F_NCCL_EXCHANGE, B.F_FORCE_I8 | B.F_NCCL_EXCHANGE,
                                   B.F_TINY_TABLE | B.F_FORCE_BITMAP_RANK],
                         ids=["int8-peer-stores", "slab-exchange", "int8-slab-exchange", "tiny-table"])
@pytest.mark.parametrize("prob", [pr.qap_esc16j(os.path.join(GOLDEN, "esc16j.npz")), pr.hamming(5, 4)],
                         ids=lambda p: p.name)
def test_sharded_path_kernel_variants(prob, flags):
    res = _check_on_ranks(prob, 2, flags)
    if flags & B.F_FORCE_I8:
        assert all(r[5]["gemm_i8"]["launches"] >= 1 for r in res), "the sharded INT8 square did not run"


@pytest.mark.parametrize("grid", [3, 4])
def test_sharded_int8_square_with_split_tail(grid, monkeypatch):
    """The K-split tail of the INT8 square under sharding: the part that draws the last ticket stores the finished
    tile into every rank's X2 (peer stores) -- same partition and blocks as the oracle (N = 1024: 8 x 4 tile grid)."""
    monkeypatch.setenv("SDPSR_I8_GRID", str(grid))
    res = _check_on_ranks(pr.hamming(5, 4), 2, B.F_FORCE_I8)
    assert all(r[5]["gemm_i8"]["launches"] >= 1 for r in res)


def test_sharded_dense_eigen_path():
    """blockDiagonalize through syevd: rank 0 factorises, Q is broadcast."""
    _check_on_ranks(pr.hamming(5, 4), 2, 0, eig="syevd")


def test_sharded_path_matches_oracle_under_torchrun():
    """One process per GPU over NCCL.  On a single-GPU box NCCL cannot host two ranks on one device; the same
    problems then run through the in-process transport (identical sharding and kernels)."""
    if B.device_count() < 2:
        for prob in _problems()[:3]:
            _check_on_ranks(prob, 2, 0)
        return
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "mgpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"] and res["world"] == 2
