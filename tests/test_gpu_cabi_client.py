"""A non-Python client of the C ABI: tests/cabi_client.c (plain C, include/sdpsr.h only) replays the call sequence of
integration/julia/SDPSRCuda.jl -- admissible_subspace with host-supplied initial elements, blockDiagonalize through
the module path with the dense fallback, `_constraints` -- and its results are held against the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle as O
from sdpsr_b200 import problems as pr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CLIENT = os.path.join(ROOT, "tests", "_build", "cabi_client")
ATOL = float(np.sqrt(np.finfo(np.float64).eps))


class Recorder:
    """Coefficient source that records every draw (the client replays them in order)."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.draws = []

    def __call__(self, n):
        v = self.rng.random(int(n))
        self.draws.append(v)
        return v


def _build_client():
    if os.path.exists(CLIENT) and os.path.getmtime(CLIENT) >= os.path.getmtime(os.path.join(ROOT, "tests", "cabi_client.c")):
        return
    import __graft_entry__ as G
    G.build_cabi_client()


@pytest.mark.parametrize("prob", [pr.lovasz_er(5), pr.lovasz_er(7), pr.qap_esc16j(os.path.join(GOLDEN, "esc16j.npz")),
                                  pr.hamming(3, 4)], ids=lambda p: p.name)
def test_c_client_matches_oracle(prob, tmp_path):
    import scipy.sparse as sp
    _build_client()
    rec = Recorder(77)
    CL, X0, _ = O.init_elements(*prob)
    Po = O.admissible_subspace(*prob, rec)
    so, bo = O.blockDiagonalize(Po, rec)
    A = sp.csr_matrix(prob.A)
    A.sort_indices()
    n = prob.n
    with open(tmp_path / "in.bin", "wb") as f:
        np.array([n, A.shape[0], A.nnz], dtype=np.int64).tofile(f)
        A.indptr.astype(np.int64).tofile(f)
        A.indices.astype(np.int64).tofile(f)
        A.data.astype(np.float64).tofile(f)
        np.asfortranarray(CL.reshape(n, n, order="F"), dtype=np.float64).reshape(-1, order="F").tofile(f)
        np.asfortranarray(X0.reshape(n, n, order="F"), dtype=np.float64).reshape(-1, order="F").tofile(f)
        np.array([ATOL, ATOL], dtype=np.float64).tofile(f)
        np.array([len(rec.draws)], dtype=np.int64).tofile(f)
        for v in rec.draws:
            np.array([v.size], dtype=np.int64).tofile(f)
            v.astype(np.float64).tofile(f)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "sdpsymmetryreduction.jl_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([CLIENT, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(tmp_path / "out.bin", "rb").read()
    off = 0

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a
    d, iters = (int(x) for x in take(np.int64, 2))
    labels = take(np.uint32, n * n).reshape(n, n, order="F")
    mode, nblk = (int(x) for x in take(np.int64, 2))
    sizes = [int(x) for x in take(np.int64, nblk)]
    sq = sum(s * s for s in sizes)
    blocks = take(np.float64, d * sq)
    ncons = int(take(np.int64, 1)[0])
    cptr = take(np.int64, d + 1)
    cidx = take(np.uint32, ncons)
    assert off == len(raw)
    assert d == Po.nparts == prob.expected_dim
    assert np.array_equal(labels, Po.matrix), "labels differ from the oracle"
    assert sizes == list(so) and mode == 0
    o = 0
    for i in range(d):
        for k, s in enumerate(sizes):
            got = blocks[o:o + s * s].reshape(s, s, order="F")
            o += s * s
            assert np.abs(got - bo[i][k]).max() < 1e-8 * max(1.0, np.abs(bo[i][k]).max()), (i, k)
    # _constraints: ascending 1-based linear indices of every class (src/diagonalize.jl:42-50)
    flat = Po.matrix.reshape(-1, order="F")
    for i in range(d):
        want = np.flatnonzero(flat == i + 1) + 1
        assert np.array_equal(cidx[cptr[i]:cptr[i + 1]], want.astype(np.uint32))
