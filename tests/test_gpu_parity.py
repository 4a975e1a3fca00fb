"""GPU parity tests: the CUDA path (through the C ABI, via ctypes) against the CPU
oracle on identical inputs and identical random coefficients.  Integer results
(labels, dims, block sizes) must be bit-exact; block values within 1e-8 relative
(the tolerance BASELINE.json's north_star states)."""
import json
import os

import numpy as np
import pytest

import oracle as O
import sdpsr_b200 as S
from sdpsr_b200 import binding as B
from sdpsr_b200 import problems as pr

from conftest import GOLDEN, Coeffs

pytestmark = pytest.mark.gpu

ATOL = O.jordan.RTOL_DEFAULT
FLAG_SETS = [0, B.F_TINY_TABLE, B.F_FORCE_BITMAP_RANK, B.F_NO_SMEM_CACHE,
             B.F_TINY_TABLE | B.F_FORCE_BITMAP_RANK | B.F_NO_SMEM_CACHE]


@pytest.fixture(scope="module")
def vec():
    with open(os.path.join(GOLDEN, "runtests_vectors.json")) as fh:
        return json.load(fh)


def oracle_refine_values(P, M, do_round=True):
    Mr = O.clamp_round(M, ATOL) if do_round else M
    P2 = O.partition_from_values(Mr)
    return P2 if P is None else O.refine(P, P2)


# ---------------------------------------------------------------------------------
# refine / labels / fill
# ---------------------------------------------------------------------------------
def test_library_reports_device():
    assert B.device_count() >= 1
    assert B.load_library().sdpsr_version() >= 100


@pytest.mark.parametrize("flags", FLAG_SETS)
def test_runtests_label_vectors(vec, flags):
    """test/runtests.jl:22-25 through set_labels / refine_labels / get_labels."""
    P1, P2, P3 = (np.array(vec[k]) for k in ("P1", "P2", "P3_coarsest_P1_P2"))
    with B.Context(3, 0, flags) as ctx:
        assert ctx.set_labels(P1) == 3
        assert np.array_equal(ctx.get_labels(), P1)
        assert ctx.refine_labels(P2) == 6
        assert np.array_equal(ctx.get_labels(), P3)
        for dt in (np.uint8, np.uint16, np.uint32, np.uint64):
            assert np.array_equal(ctx.get_labels(dt), P3)
        assert ctx.zero_count() == 0


@pytest.mark.parametrize("flags", FLAG_SETS)
@pytest.mark.parametrize("n,nvals,seed", [(10, 10, 0), (10, 10, 1), (37, 5, 2), (64, 3000, 3), (129, 40, 4)])
def test_partition_ctor_matches_oracle(n, nvals, seed, flags):
    """test/runtests.jl:13-20: integer and float constructors, with and without zeros."""
    rng = np.random.default_rng(seed)
    M = rng.integers(0 if seed % 2 == 0 else 1, nvals + 1, size=(n, n))
    want = O.partition_from_values(M)
    with B.Context(n, 0, flags) as ctx:
        assert ctx.set_labels(M) == want.nparts
        assert np.array_equal(ctx.get_labels(), want.matrix)
        ctx.reset()
        assert ctx.refine_values(M.astype(np.float64), ATOL, do_round=False) == want.nparts
        assert np.array_equal(ctx.get_labels(), want.matrix)
        assert ctx.zero_count() == int((M == 0).sum())


def adversarial_values(rng, n):
    """Values that stress the truncation grid: powers of two, fractions one
    ulp below 1, values around atol, both signs, wide exponent range, exact decimals."""
    base = np.concatenate([
        2.0 ** rng.integers(-40, 40, size=50),
        np.nextafter(2.0 ** rng.integers(-20, 20, size=50), 0),
        np.nextafter(2.0 ** rng.integers(-20, 20, size=50), np.inf),
        np.array([ATOL, np.nextafter(ATOL, 0), np.nextafter(ATOL, 1), -ATOL, 1e-9, -1e-9, 0.0, 1 / 16, 0.3, 0.1, 1 / 3,
                  0.99999995, 0.9999999, 0.5, 0.50000005, 1e300, -1e300, 1e-300, 123456.7, 1234567.8]),
        rng.standard_normal(100) * 10.0 ** rng.integers(-6, 6, size=100),
        np.round(rng.random(100), 7),
    ])
    base = np.concatenate([base, -base])
    return rng.choice(base, size=(n, n))


@pytest.mark.parametrize("flags", FLAG_SETS)
@pytest.mark.parametrize("n,seed", [(16, 0), (33, 1), (100, 2)])
def test_round_refine_chain_matches_oracle(n, seed, flags):
    """refine!(S, Part(_clamp_round!(M))) repeatedly, with adversarial values."""
    rng = np.random.default_rng(seed)
    P = None
    with B.Context(n, 0, flags) as ctx:
        for step in range(4):
            if step % 2 == 0:
                M = adversarial_values(rng, n)
            else:
                M = rng.integers(0, 4, size=(n, n)) * 0.125 + rng.integers(0, 2, size=(n, n)) * 1e-9
            d = ctx.refine_values(M, ATOL, do_round=True)
            P = oracle_refine_values(P, M, True)
            assert d == P.nparts, (step, d, P.nparts)
            assert np.array_equal(ctx.get_labels(), P.matrix), f"labels differ at step {step}"


@pytest.mark.parametrize("atol", [1e-3, 1e-5, 1.4901161193847656e-8, 1e-10, 1e-13])
def test_round_refine_other_tolerances(atol):
    """sigdigits = floor(-log10(atol)) other than 7, incl. grids too wide for the fused key."""
    rng = np.random.default_rng(5)
    n = 40
    with B.Context(n, 0, 0) as ctx:
        P = None
        for step in range(3):
            M = np.round(rng.random((n, n)), 3) + rng.integers(0, 3, size=(n, n)) * 1e-6
            d = ctx.refine_values(M, atol, True)
            Mr = O.clamp_round(M, atol)
            P2 = O.partition_from_values(Mr)
            P = P2 if P is None else O.refine(P, P2)
            assert d == P.nparts
            assert np.array_equal(ctx.get_labels(), P.matrix)


@pytest.mark.parametrize("flags", [0, B.F_TINY_TABLE], ids=["default", "tiny-table"])
def test_round_refine_chain_with_many_classes(flags):
    """dim > 4096: the per-CTA key cache is bypassed and the batched global-table probes of refine_fast_kernel
    run (four probes of a thread issued together); labels stay bit-identical to the oracle."""
    rng = np.random.default_rng(17)
    n = 160
    with B.Context(n, 0, flags) as ctx:
        P = None
        for step in range(4):
            M = rng.integers(0, 3000, size=(n, n)) * 0.37 + 0.11
            M[rng.random((n, n)) < 0.05] = 0.0
            d = ctx.refine_values(M, ATOL, do_round=True)
            P = oracle_refine_values(P, M, True)
            assert d == P.nparts, (step, d, P.nparts)
            assert np.array_equal(ctx.get_labels(), P.matrix), f"labels differ at step {step}"
        assert P.nparts > 4096


def test_negative_zero_and_raw_bits():
    M = np.array([[0.0, -0.0, 1.0], [1.0, 0.0, -0.0], [2.0, 2.0, 0.0]])
    want = O.partition_from_values(M)
    with B.Context(3) as ctx:
        assert ctx.refine_values(M, ATOL, do_round=False) == want.nparts == 3
        assert np.array_equal(ctx.get_labels(), want.matrix)


@pytest.mark.parametrize("n", [3, 10, 31, 64, 100])
def test_fill_matches_oracle(n):
    rng = np.random.default_rng(n)
    M = rng.integers(0, 7, size=(n, n))
    P = O.partition_from_values(M)
    r = rng.random(P.nparts)
    with B.Context(n) as ctx:
        ctx.set_labels(M)
        ctx.fill(r)
        X = ctx.get_matrix(B.MAT_X)
    assert np.array_equal(X, O.fill(P, r))           # bit-exact gather


def test_randomize_roundtrip(vec):
    """test/runtests.jl:27: part(rndPart(P1)) == P1."""
    P1 = S.Partition(np.array(vec["P1"]))
    assert P1.nparts == 3
    X = S.randomize(P1, Coeffs(3))
    assert S.Partition(X) == P1


def test_fill_length_is_checked():
    with B.Context(4) as ctx:
        ctx.set_labels(np.arange(16).reshape(4, 4) % 3)
        with pytest.raises(B.SdpsrError) as e:
            ctx.fill(np.ones(5))
        assert e.value.code == B.E_INVALID


def test_uint16_overflow_like_reference():
    """labels that do not fit the requested width raise."""
    n = 300
    M = np.arange(n * n).reshape(n, n) + 1            # 90 000 classes > 65 535
    with B.Context(n) as ctx:
        assert ctx.set_labels(M) == n * n
        with pytest.raises(B.SdpsrError) as e:
            ctx.get_labels(np.uint16)
        assert e.value.code == B.E_LABEL_OVERFLOW
        L = ctx.get_labels(np.uint32)
        assert np.array_equal(L, O.partition_from_values(M).matrix)


# ---------------------------------------------------------------------------------
# GEMM
# ---------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [3, 16, 31, 57, 128, 130, 256, 300, 515])
def test_gemm_matches_fp64_reference(n):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    Bm = rng.standard_normal((n, n))
    with B.Context(n) as ctx:
        ctx.set_matrix(B.MAT_X, A)
        ctx.set_matrix(B.MAT_Q, Bm)
        ctx.gemm(B.MAT_X, B.MAT_Q, B.MAT_X2)
        Cm = ctx.get_matrix(B.MAT_X2)
    ref = A @ Bm
    scale = np.abs(A) @ np.abs(Bm)
    assert np.max(np.abs(Cm - ref) / scale) < 1e-14      # fp64 tolerance: a few ulp of the sum of |terms|


# ---------------------------------------------------------------------------------
# admissible_subspace
# ---------------------------------------------------------------------------------
def small_problems():
    return [pr.petersen(), pr.lovasz_er(3), pr.lovasz_er(5), pr.lovasz_er(7),
            pr.qap_esc16j(os.path.join(GOLDEN, "esc16j.npz")), pr.hamming(2, 8), pr.hamming(3, 4),
            pr.kneser(8, 3), pr.kneser(10, 4), pr.synthetic_product_scheme(3, 2, 8),
            pr.synthetic_product_scheme(3, 3, 16)]


@pytest.mark.parametrize("prob", small_problems(), ids=lambda p: p.name)
@pytest.mark.parametrize("flags", [0, B.F_TINY_TABLE | B.F_FORCE_BITMAP_RANK, B.F_FORCE_I8])
def test_admissible_subspace_labels_identical(prob, flags):
    """Final partition identical (bit-exact canonical labels) and identical dim
    trajectory, given the same random coefficients."""
    co, cg = Coeffs(11), Coeffs(11)
    Po, tr_o = O.admissible_subspace_trace(*prob, co)
    tr_g = {}
    Pg = S.admissible_subspace(*prob, rand=cg, flags=flags, trace=tr_g, keep_context=False)
    assert tr_g["init"] == tr_o["init"]
    assert tr_g["iters"] == tr_o["iters"]
    assert cg.draws == co.draws                        # same rand() call sequence
    assert Pg.nparts == Po.nparts == prob.expected_dim
    assert np.array_equal(Pg.matrix, Po.matrix)


@pytest.mark.parametrize("prob", [pr.lovasz_er(5), pr.qap_esc16j(os.path.join(GOLDEN, "esc16j.npz"))],
                         ids=lambda p: p.name)
def test_host_supplied_init_elements(prob):
    """The Julia-wrapper route: CL and X0 computed by the host (the reference's own qr/craig)."""
    CL, X0, _ = O.init_elements(*prob)
    co, cg = Coeffs(5), Coeffs(5)
    Po = O.admissible_subspace(*prob, co)
    Pg = S.admissible_subspace(*prob, rand=cg, init_elements=(CL, X0), keep_context=False)
    assert np.array_equal(Pg.matrix, Po.matrix)


def test_dense_and_sparse_constraints_agree():
    prob = pr.lovasz_er(5)
    import scipy.sparse as sp
    Pd = S.admissible_subspace(prob.C, prob.A, prob.b, rand=Coeffs(1), keep_context=False)
    Ps = S.admissible_subspace(prob.C, sp.csr_matrix(prob.A), prob.b, rand=Coeffs(1), keep_context=False)
    assert Pd == Ps
    with B.Context(prob.n) as ctx:
        csc = sp.csc_matrix(prob.A)
        ctx.set_constraints_csc(2, csc.indptr, csc.indices, csc.data, 0)
        assert ctx.constraint_patterns() == 2
        assert ctx.init_partition(prob.C, prob.b, ATOL) == 2


def test_staged_objective_upload():
    """sdpsr_stage_objective starts the upload of a host C on the copy stream (it overlaps the constraint set-up);
    init_partition consumes it only for the same buffer with nothing but set_constraints in between, and copies as
    usual otherwise -- the partition is the same in every case."""
    prob = pr.qap_esc16j(os.path.join(GOLDEN, "esc16j.npz"))
    Cv = np.ascontiguousarray(np.asarray(prob.C, dtype=np.float64).reshape(-1))
    junk = np.random.default_rng(1).random(Cv.size)
    with B.Context(prob.n) as ctx:
        ctx.set_constraints(prob.A)
        want_d = ctx.init_partition(Cv, prob.b, ATOL)
        want = ctx.get_labels()
    assert want_d > 1

    def run(steps):
        with B.Context(prob.n) as ctx:
            for st in steps:
                st(ctx)
            d = ctx.init_partition(Cv, prob.b, ATOL)
            return d, ctx.get_labels()

    cases = {
        "staged": [lambda c: c.stage_objective(Cv), lambda c: c.set_constraints(prob.A)],
        "staged after the constraints": [lambda c: c.set_constraints(prob.A), lambda c: c.stage_objective(Cv)],
        "another buffer staged": [lambda c: c.stage_objective(junk), lambda c: c.set_constraints(prob.A)],
        "dropped by a call in between": [lambda c: c.stage_objective(junk), lambda c: c.set_constraints(prob.A),
                                         lambda c: c.dim()],
        "staged twice": [lambda c: c.stage_objective(junk), lambda c: c.stage_objective(Cv),
                         lambda c: c.set_constraints(prob.A)],
    }
    for name, steps in cases.items():
        d, lab = run(steps)
        assert d == want_d and np.array_equal(lab, want), name


def test_async_label_export():
    """With a caller-owned label buffer admissible_subspace starts the export on the copy stream and returns;
    blockDiagonalize runs without the host copy, P.matrix waits for it on first access."""
    prob = pr.lovasz_er(5)
    Po = O.admissible_subspace(*prob, Coeffs(7))
    so, bo = O.blockDiagonalize(Po, Coeffs(8))
    out = np.zeros((prob.n, prob.n), dtype=np.uint16, order="F")
    P = S.admissible_subspace(*prob, rand=Coeffs(7), labels_out=out, label_dtype=np.uint16)
    assert P._arrival is not None
    bd = S.blockDiagonalize(P, False, rand=Coeffs(8))
    assert P._arrival is not None                      # nothing on that path needed the host copy
    M = P.matrix
    assert P._arrival is None and np.shares_memory(M, out)
    assert np.array_equal(M, Po.matrix) and list(bd.blkSizes) == list(so)
    P.release()
    # a second export while one is pending, and release() with one pending, are both safe
    out2 = np.zeros((prob.n, prob.n), dtype=np.uint32, order="F")
    P2 = S.admissible_subspace(*prob, rand=Coeffs(7), labels_out=out2, label_dtype=np.uint32)
    P2._ctx.get_labels_async(np.uint16, out)
    P2.release()
    assert np.array_equal(out2, Po.matrix) and np.array_equal(out, Po.matrix)
    # a label type that cannot hold dim(P) is refused before anything is started
    rng = np.random.default_rng(3)
    Mbig = rng.integers(1, 400, size=(40, 40))
    with B.Context(40) as ctx:
        assert ctx.set_labels(Mbig.astype(np.int64)) > 255
        with pytest.raises(B.SdpsrError) as ei:
            ctx.get_labels_async(np.uint8, np.zeros((40, 40), dtype=np.uint8, order="F"))
        assert ei.value.code == B.E_LABEL_OVERFLOW
        ctx.labels_wait()                                # nothing pending: a no-op


def test_copy_stream_with_page_locked_buffers():
    """Page-locked host buffers take the kernel-driven transfers (pcie_copy_kernel: zero-copy loads of C, zero-copy
    stores of the labels); pageable ones take cudaMemcpyAsync.  Same partition either way."""
    torch = pytest.importorskip("torch")
    prob = pr.hamming(5, 3, sparse=True)                       # N = 243 (ld = 256 != N: staged through the 2-D copy)
    prob2 = pr.hamming(4, 4, sparse=True)                      # N = 256 (ld == N: the kernel path)
    for q in (prob, prob2):
        Po = O.admissible_subspace(*q, Coeffs(5))
        Cpin = torch.from_numpy(np.ascontiguousarray(np.asarray(q.C, dtype=np.float64).reshape(-1))).pin_memory()
        Lpin = torch.zeros(q.n * q.n, dtype=torch.int16).pin_memory()
        out = Lpin.numpy().view(np.uint16).reshape(q.n, q.n, order="F")
        P = S.admissible_subspace(Cpin.numpy(), q.A, q.b, rand=Coeffs(5), labels_out=out, label_dtype=np.uint16)
        assert P._arrival is not None
        bd = S.blockDiagonalize(P, False, rand=Coeffs(6))
        assert np.array_equal(P.matrix, Po.matrix) and np.shares_memory(P.matrix, out)
        so, _ = O.blockDiagonalize(Po, Coeffs(6))
        assert list(bd.blkSizes) == list(so)
        P.release()


def test_uint16_default_label_type():
    prob = pr.lovasz_er(3)
    P = S.admissible_subspace(*prob, rand=Coeffs(2), label_dtype=np.uint16, keep_context=False)
    assert P.matrix.dtype == np.uint16 and P.nparts == 12


def test_medium_hamming_closed_form():
    """H(3,8), N = 512: labels must equal Hamming distance + 1."""
    prob = pr.hamming(3, 8)
    P = S.admissible_subspace(*prob, rand=Coeffs(3))
    assert P.nparts == 4
    assert np.array_equal(P.matrix, pr.hamming_distance_matrix(3, 8).astype(np.uint32) + 1)
    bd = S.blockDiagonalize(P, False, rand=Coeffs(4))
    assert bd.blkSizes == [1, 1, 1, 1]
    K = pr.krawtchouk(3, 8)
    got = np.array([[bd.blks[i][k][0, 0] for k in range(4)] for i in range(4)])
    for k in range(4):
        assert min(np.abs(K - got[:, [k]]).max(axis=0)) < 1e-8 * np.abs(K).max()
    P.release()


# ---------------------------------------------------------------------------------
# blockDiagonalize
# ---------------------------------------------------------------------------------
@pytest.mark.parametrize("eig", ["syevd", "auto"])
@pytest.mark.parametrize("prob", small_problems(), ids=lambda p: p.name)
def test_block_diagonalize_matches_oracle(prob, eig):
    """Same coefficient vectors -> same blocks as the oracle's dense restatement, on the reference's own
    algorithm (``syevd``) and on the default path (``auto``: Krylov variant where it applies)."""
    Po = O.admissible_subspace(*prob, Coeffs(21))
    Pg = S.Partition(Po.nparts, Po.matrix.astype(np.uint32))
    so, bo = O.blockDiagonalize(Po, Coeffs(22))
    cg = Coeffs(22)
    bd = S.blockDiagonalize(Pg, False, rand=cg, eig=eig)
    assert cg.draws == [Po.nparts] * 3                  # three draws of dim(P), whatever the path (A.5)
    assert Pg._eig_mode == "syevd" if eig == "syevd" else Pg._eig_mode in ("krylov", "syevd")
    assert list(bd.blkSizes) == list(so)
    assert sorted(bd.blkSizes) == prob.expected_blocks
    for i in range(Po.nparts):
        for k in range(len(so)):
            ref = bo[i][k]
            tol = 1e-8 * max(1.0, np.abs(ref).max())
            assert np.abs(bd.blks[i][k] - ref).max() < tol, (i, k)
    Pg.release()


KRYLOV_PROBLEMS = [pr.petersen(), pr.lovasz_er(3), pr.lovasz_er(5), pr.lovasz_er(7), pr.kneser(8, 3),
                   pr.kneser(10, 4), pr.hamming(3, 8), pr.hamming(5, 4), pr.qap_esc16j(os.path.join(GOLDEN, "esc16j.npz")),
                   pr.synthetic_product_scheme(3, 2, 8), pr.synthetic_product_scheme(3, 3, 16)]


@pytest.mark.parametrize("flags", [0, B.F_TINY_TABLE], ids=["single-pass-basis", "list-basis"])
@pytest.mark.parametrize("prob", KRYLOV_PROBLEMS, ids=lambda p: p.name)
def test_krylov_block_diagonalize_matches_oracle_and_dense(prob, flags):
    """The matrix-free variant (csrc/krylov.cu) must be applicable on these few-eigenspace problems and
    reproduce the oracle's blocks, block order, multiplicities and the dense path's results.  With few
    classes basis_image is the single-pass kernel; the TINY_TABLE hook keeps the general kernel
    (per-class lists) under test."""
    Po = O.admissible_subspace(*prob, Coeffs(31))
    so, bo = O.blockDiagonalize(Po, Coeffs(32))
    Pk = S.Partition(Po.nparts, Po.matrix.astype(np.uint32))
    Pk._context(flags)
    Pd = S.Partition(Po.nparts, Po.matrix.astype(np.uint32))
    bk = S.blockDiagonalize(Pk, False, rand=Coeffs(32), eig="krylov")
    bdn = S.blockDiagonalize(Pd, False, rand=Coeffs(32), eig="syevd")
    assert Pk._eig_mode == "krylov" and Pd._eig_mode == "syevd"
    assert list(bk.blkSizes) == list(bdn.blkSizes) == list(so)
    assert np.array_equal(Pk._ptrs, Pd._ptrs) and np.array_equal(Pk._kroot, Pd._kroot)
    for i in range(Po.nparts):
        for k in range(len(so)):
            tol = 1e-8 * max(1.0, np.abs(bo[i][k]).max())
            assert np.abs(bk.blks[i][k] - bo[i][k]).max() < tol, (i, k)
            assert np.abs(bk.blks[i][k] - bdn.blks[i][k]).max() < tol, (i, k)
    Qk = S.diagonalize(Pk, rand=Coeffs(33), eig="krylov", atol=1e-9)
    for q in Qk:                                        # orthonormal columns
        assert np.allclose(q.T @ q, np.eye(q.shape[1]), atol=1e-9)
    Pk.release()
    Pd.release()


@pytest.mark.parametrize("kind", ["S3", "D4", "D5", "Q8", "D7", "S4", "D16"])
def test_module_path_on_jordan_partitions_that_are_not_coherent(kind):
    """Symmetrised regular representation of a non-abelian group: a Jordan algebra that is not closed under
    products, so the orbit of a unit vector is not invariant and the module has to be closed under the generic
    element (csrc/krylov.cu).  Same blocks as the oracle's dense restatement, same as the device's dense path."""
    L = pr.symmetrized_group_partition(kind)
    Po = O.Partition(int(L.max()), L.astype(np.uint32))
    so, bo = O.blockDiagonalize(Po, Coeffs(52))
    Pk = S.Partition(Po.nparts, Po.matrix.astype(np.uint32))
    Pd = S.Partition(Po.nparts, Po.matrix.astype(np.uint32))
    bk = S.blockDiagonalize(Pk, False, rand=Coeffs(52), eig="krylov")
    bdn = S.blockDiagonalize(Pd, False, rand=Coeffs(52), eig="syevd")
    assert list(bk.blkSizes) == list(bdn.blkSizes) == list(so)
    assert np.array_equal(Pk._ptrs, Pd._ptrs) and np.array_equal(Pk._kroot, Pd._kroot)
    for i in range(Po.nparts):
        for k in range(len(so)):
            tol = 1e-8 * max(1.0, np.abs(bo[i][k]).max())
            assert np.abs(bk.blks[i][k] - bo[i][k]).max() < tol, (i, k)
            assert np.abs(bk.blks[i][k] - bdn.blks[i][k]).max() < tol, (i, k)
    Pk.release()
    Pd.release()


def test_krylov_not_applicable_falls_back_to_dense():
    """A module larger than the cap: ``eig="krylov"`` reports it, ``auto`` lands on the dense path with the
    same coefficient vectors and the same blocks."""
    import sdpsr_b200.api as api
    prob = pr.qap_esc16j(os.path.join(GOLDEN, "esc16j.npz"))
    Po = O.admissible_subspace(*prob, Coeffs(41))
    so, bo = O.blockDiagonalize(Po, Coeffs(42))
    P = S.Partition(Po.nparts, Po.matrix.astype(np.uint32))
    old = api.KRYLOV_MAX_MODULE_DIM
    api.KRYLOV_MAX_MODULE_DIM = 20          # the module of esc16j has dimension 45
    try:
        with pytest.raises(S.NumericalInconsistency):
            S.blockDiagonalize(P, False, rand=Coeffs(42), eig="krylov")
        cg = Coeffs(42)
        bd = S.blockDiagonalize(P, False, rand=cg, eig="auto")
    finally:
        api.KRYLOV_MAX_MODULE_DIM = old
    assert P._eig_mode == "syevd" and cg.draws == [150] * 3
    assert list(bd.blkSizes) == list(so)
    for i in range(0, Po.nparts, 7):
        for k in range(len(so)):
            assert np.abs(bd.blks[i][k] - bo[i][k]).max() < 1e-8 * max(1.0, np.abs(bo[i][k]).max())
    P.release()


def test_krylov_rejects_nonsymmetric(vec):
    P3 = S.Partition(np.array(vec["C3"]["matrix"]))
    with pytest.raises(S.InvalidDecompositionField):
        S.blockDiagonalize(P3, False, rand=Coeffs(1), eig="krylov")


def test_diagonalize_default_atol_and_qhat(vec):
    """diagonalize(Float64, P) as called in test/lovasz.jl:7 (atol = 1e-12*N)."""
    prob = pr.lovasz_er(7)
    P = S.admissible_subspace(*prob, rand=Coeffs(1))
    Qhat = S.diagonalize(P, rand=Coeffs(2))
    assert sorted(q.shape[1] for q in Qhat) == [2, 2, 2, 2, 3]
    for q in Qhat:                                      # orthonormal columns
        assert np.allclose(q.T @ q, np.eye(q.shape[1]), atol=1e-10)
    blks = S.basis_image(Qhat, P)
    ref = O.basis_image(Qhat, O.Partition(P.nparts, P.matrix.astype(np.int64)))
    for i in range(P.nparts):
        for k in range(len(Qhat)):
            assert np.abs(blks[i][k] - ref[i][k]).max() < 1e-10
    P.release()


def test_real_path_rejects_nonsymmetric(vec):
    """test/runtests.jl:50-56: cyclic C3 over the reals throws InvalidDecompositionField."""
    P3 = S.Partition(np.array(vec["C3"]["matrix"]))
    with pytest.raises(S.InvalidDecompositionField):
        S.blockDiagonalize(P3, False, rand=Coeffs(1))


def test_desymmetrize_identity(vec):
    """test/runtests.jl:40."""
    P1 = S.Partition(np.array(vec["P1"]))
    got = S.unSymmetrize(P1, rand=Coeffs(7))
    want = vec["unsymmetrize_P1"]
    assert got.nparts == want["nparts"]
    assert np.array_equal(got.matrix, np.array(want["matrix"]))
    assert got == S.Partition(O.desymmetrize(O.partition_from_values(np.array(vec["P1"])), Coeffs(7)).matrix)


def test_numerical_issues_fixture():
    """test/numerical_issues.jl:91-94 (bounded to 100 trials): never inconsistent at atol 1e-7."""
    Pm = np.load(os.path.join(GOLDEN, "numerical_issues_P.npy"))
    P = S.Partition(Pm)
    assert P.nparts == 1312
    c = Coeffs(99)
    from sdpsr_b200.api import eigen_decomposition
    for _ in range(100):
        vals, ptrs, kroot = eigen_decomposition(P, atol=1e-7, rand=c)
    assert len(ptrs) - 1 == 64
    assert sorted(np.bincount(kroot)[np.unique(kroot)]) == [16, 48]
    P.release()


def test_spectrum_invariant_gpu():
    prob = pr.lovasz_er(7)
    P = S.admissible_subspace(*prob, rand=Coeffs(8))
    bd = S.blockDiagonalize(P, False, rand=Coeffs(9))
    x = np.random.default_rng(0).random(P.nparts)
    big = np.linalg.eigvalsh(np.concatenate([[0.0], x])[P.matrix])
    mult = [int(P._ptrs[r + 1] - P._ptrs[r]) for r in dict.fromkeys(P._kroot.tolist())]
    small = []
    for k, s in enumerate(bd.blkSizes):
        Mk = sum(x[i] * bd.blks[i][k] for i in range(P.nparts))
        small += list(np.linalg.eigvalsh(Mk)) * mult[k]
    assert np.allclose(sorted(small), big, atol=1e-9)
    assert sorted(mult) == prob.expected_mult
    P.release()


# ---------------------------------------------------------------------------------
# half-GEMM for symmetric X (SYRK-style) and stream hand-over
# ---------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [57, 130, 300, 515])
def test_symmetric_square_uses_half_gemm_and_matches(n):
    rng = np.random.default_rng(n)
    L = rng.integers(0, 9, size=(n, n))
    L = np.minimum(L, L.T)
    P = O.partition_from_values(L)
    r = rng.random(P.nparts)
    X = O.fill(P, r)
    ref = X @ X
    outs = []
    for flags in (0, B.F_NO_SYRK, B.F_FORCE_I8):
        with B.Context(n, 0, flags | B.F_TIMING) as ctx:
            ctx.set_labels(L)
            ctx.fill(r)
            d = ctx.square_round_refine(ATOL)
            X2 = ctx.get_matrix(B.MAT_X2)
            outs.append((d, ctx.get_labels(), X2))
            assert (ctx.timing()["gemm_i8"]["launches"] == 1) == (flags == B.F_FORCE_I8)
    assert np.max(np.abs(outs[0][2] - ref)) < 1e-12 * np.abs(ref).max()
    assert np.max(np.abs(outs[2][2] - ref)) < 1e-12 * np.abs(ref).max()     # INT8 path: FP64-grade
    assert np.array_equal(outs[0][2], outs[0][2].T)                  # mirrored: exactly symmetric
    assert np.array_equal(outs[2][2], outs[2][2].T)
    want = O.refine(P, O.partition_from_values(O.clamp_round(ref, ATOL)))
    for d, lab, _ in outs:
        assert d == want.nparts and np.array_equal(lab, want.matrix)


def _sym_matrix(n, rng, kind):
    if kind == "lut":          # what the closure loop squares: few distinct uniform values, symmetric pattern
        lab = rng.integers(0, 8, size=(n, n))
        lab = np.triu(lab) + np.triu(lab, 1).T
        return np.asfortranarray(np.concatenate([[0.0], rng.random(7)])[lab])
    A = rng.standard_normal((n, n)) * np.exp(rng.uniform(-6, 2, size=(n, n)))      # both signs, wide range
    return np.asfortranarray(np.triu(A) + np.triu(A, 1).T)


@pytest.mark.parametrize("kind", ["lut", "wide"])
@pytest.mark.parametrize("n", [1, 15, 130, 257, 384])
def test_int8_square_bit_exact_against_host_model(n, kind):
    """csrc/gemm_i8.cu is integer-exact: the device result equals the host model (tests/i8_model.py)
    bit for bit, is exactly symmetric, and at 8 digits agrees with the FP64 product to dgemm accuracy."""
    from i8_model import exact_square
    rng = np.random.default_rng(1000 + n)
    X = _sym_matrix(n, rng, kind)
    ref = X @ X
    with B.Context(n) as ctx:
        ctx.set_matrix(B.MAT_X, X)
        for method, bits, S_ in ((2, 7, 8), (2, 7, 7), (2, 7, 4), (2, 7, 2), (3, 8, 7), (3, 8, 5), (3, 8, 2), (1, 8, 0)):
            ctx.square(method, S_)
            got = ctx.get_matrix(B.MAT_X2)
            assert np.array_equal(got, exact_square(X, S_ or 7, bits)), (n, kind, bits, S_)
            assert np.array_equal(got, got.T)
            if S_ in (8, 0) or (bits, S_) == (8, 7):
                assert np.max(np.abs(got - ref)) <= 1e-13 * np.abs(ref).max()
        ctx.square(0)                                            # DMMA on the same X
        assert np.max(np.abs(ctx.get_matrix(B.MAT_X2) - ref)) <= 1e-13 * max(np.abs(ref).max(), 1e-300)


@pytest.mark.parametrize("n", [200, 300, 515])
def test_int8_square_with_k_segments(n, monkeypatch):
    """Long K is cut into segments that fit the int32 accumulators (16384 for 8-bit digits); the test hook
    SDPSR_I8_SEGBLOCKS makes the segments 128 long so that the same code runs at small N."""
    from i8_model import exact_square
    rng = np.random.default_rng(n)
    X = _sym_matrix(n, rng, "wide")
    monkeypatch.setenv("SDPSR_I8_SEGBLOCKS", "1")
    with B.Context(n) as ctx:
        ctx.set_matrix(B.MAT_X, X)
        for method, bits, S_ in ((3, 8, 7), (2, 7, 8), (3, 8, 3)):
            ctx.square(method, S_)
            got = ctx.get_matrix(B.MAT_X2)
            assert np.array_equal(got, exact_square(X, S_, bits, kseg=128)), (n, bits, S_)
            assert np.array_equal(got, got.T)
    ref = X @ X
    assert np.max(np.abs(exact_square(X, 7, 8, kseg=128) - ref)) <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("grid", [2, 3, 5, 7, 11])
@pytest.mark.parametrize("n", [515, 700, 1100])
def test_int8_square_split_tail(n, grid, monkeypatch):
    """The tiles of the last, partly filled wave are cut along K over all CTAs (int32 parts in scratch slots, the
    part that draws the last ticket adds them and folds): same bits as the unsplit schedule and the host model.
    SDPSR_I8_GRID caps the number of CTAs so that small problems have waves and a tail; SDPSR_I8_TAIL=0 is the
    unsplit schedule."""
    from i8_model import exact_square
    rng = np.random.default_rng(7 * n + grid)
    X = _sym_matrix(n, rng, "wide")
    items, info = B.i8_schedule(n, 1, 0, grid)
    monkeypatch.setenv("SDPSR_I8_GRID", str(grid))
    with B.Context(n) as ctx:
        ctx.set_matrix(B.MAT_X, X)
        for method, bits, S_ in ((3, 8, 7), (2, 7, 8), (3, 8, 4)):
            want = exact_square(X, S_, bits)
            monkeypatch.setenv("SDPSR_I8_TAIL", "1")
            ctx.square(method, S_)
            got = ctx.get_matrix(B.MAT_X2)
            assert np.array_equal(got, want), (n, grid, bits, S_, info)
            assert np.array_equal(got, got.T)
            monkeypatch.setenv("SDPSR_I8_TAIL", "0")
            ctx.square(method, S_)
            assert np.array_equal(ctx.get_matrix(B.MAT_X2), want)
    # the parametrisation must really exercise the split path somewhere
    if (n, grid) in ((700, 5), (1100, 11), (515, 7)):
        assert info["nslots"] > 0, info


@pytest.mark.parametrize("n", [1, 130, 257, 515, 700])
def test_int8_square_cta_pair_kernel(n, monkeypatch):
    """The cta_group::2 variant (256 x 256 tiles on CTA pairs, default for N > 16384) computes the same
    bits; SDPSR_I8_PAIR=1 selects it at any N, SDPSR_I8_SEGBLOCKS=2 adds K segments."""
    from i8_model import exact_square
    rng = np.random.default_rng(50 + n)
    X = _sym_matrix(n, rng, "wide")
    monkeypatch.setenv("SDPSR_I8_PAIR", "1")
    with B.Context(n, 0, B.F_TIMING) as ctx:
        ctx.set_matrix(B.MAT_X, X)
        for method, bits, S_ in ((3, 8, 7), (2, 7, 8), (3, 8, 2)):
            ctx.square(method, S_)
            got = ctx.get_matrix(B.MAT_X2)
            assert np.array_equal(got, exact_square(X, S_, bits)), (n, bits, S_)
            assert np.array_equal(got, got.T)
    monkeypatch.setenv("SDPSR_I8_SEGBLOCKS", "2")
    with B.Context(n) as ctx:
        ctx.set_matrix(B.MAT_X, X)
        ctx.square(3, 7)
        assert np.array_equal(ctx.get_matrix(B.MAT_X2), exact_square(X, 7, 8, kseg=256))


def test_int8_square_degenerate_inputs():
    n = 40
    with B.Context(n) as ctx:
        ctx.set_matrix(B.MAT_X, np.zeros((n, n)))
        ctx.square(1, 8)
        assert not ctx.get_matrix(B.MAT_X2).any()
        X = np.arange(n * n, dtype=np.float64).reshape(n, n)                  # not symmetric
        ctx.set_matrix(B.MAT_X, X)
        with pytest.raises(B.SdpsrError) as ei:
            ctx.square(1, 8)
        assert ei.value.code == B.E_INVALID
        Xs = X + X.T
        Xs[3, 5] = Xs[5, 3] = np.inf                                           # Inf: outside the slicing range
        ctx.set_matrix(B.MAT_X, Xs)
        with pytest.raises(B.SdpsrError) as ei:
            ctx.square(1, 8)
        assert ei.value.code == B.E_UNSUPPORTED
        with pytest.raises(B.SdpsrError):
            ctx.set_square_slices(9)
        with pytest.raises(B.SdpsrError):
            ctx.square(3, 8)                                                   # 8-bit digits: at most 7 slices
        ctx.set_square_slices(6)


def test_int8_square_leaves_badly_scaled_rows_to_dmma():
    """One scale serves the whole matrix, so a row far below the maximum goes to the FP64 path."""
    n = 96
    rng = np.random.default_rng(5)
    L = rng.integers(1, 5, size=(n, n))
    L = np.minimum(L, L.T)
    L[7, :] = L[:, 7] = 5
    P = O.partition_from_values(L)
    r = rng.random(P.nparts)
    r[int(P.matrix[7, 0]) - 1] = 1e-9
    X = O.fill(P, r)
    want = O.refine(P, O.partition_from_values(O.clamp_round(X @ X, ATOL)))
    for tiny_row, launches in ((True, (0, 1)), (False, (1, 0))):
        with B.Context(n, 0, B.F_FORCE_I8 | B.F_TIMING) as ctx:
            ctx.set_labels(L)
            rr = r.copy()
            if not tiny_row:
                rr[int(P.matrix[7, 0]) - 1] = 0.3
            ctx.fill(rr)
            d = ctx.square_round_refine(ATOL)
            t = ctx.timing()
            assert (t["gemm_i8"]["launches"], t["gemm"]["launches"]) == launches
            if tiny_row:
                assert d == want.nparts and np.array_equal(ctx.get_labels(), want.matrix)


def test_nonsymmetric_square_falls_back_to_full_gemm():
    n = 150
    rng = np.random.default_rng(1)
    L = rng.integers(1, 6, size=(n, n))
    P = O.partition_from_values(L)
    r = rng.random(P.nparts)
    X = O.fill(P, r)
    with B.Context(n) as ctx:
        ctx.set_labels(L)
        ctx.fill(r)
        ctx.square_round_refine(ATOL)
        X2 = ctx.get_matrix(B.MAT_X2)
    assert np.max(np.abs(X2 - X @ X)) < 1e-12 * np.abs(X @ X).max()


def test_caller_stream():
    import torch
    prob = pr.lovasz_er(5)
    with B.Context(prob.n) as ctx:
        s = torch.cuda.Stream()
        ctx.set_stream(s.cuda_stream)
        P = S.admissible_subspace(*prob, rand=Coeffs(3), ctx=ctx)
        assert P.nparts == 15
        bd = S.blockDiagonalize(P, False, rand=Coeffs(4))
        assert sorted(bd.blkSizes) == [2, 2, 2, 3]


# ---------------------------------------------------------------------------------
# reduced-SDP assembly (README.md:57-60) and _constraints
# ---------------------------------------------------------------------------------
@pytest.mark.parametrize("prob", [pr.lovasz_er(5), pr.qap_esc16j(os.path.join(GOLDEN, "esc16j.npz")),
                                  pr.synthetic_product_scheme(3, 2, 8)], ids=lambda p: p.name)
def test_reduce_problem_matches_pmat(prob):
    import scipy.sparse as sp
    P = S.admissible_subspace(*prob, rand=Coeffs(4))
    newA, newB, newC = S.reduce_problem(P, prob.C, prob.A, prob.b)
    flat = P.matrix.reshape(-1, order="F").astype(np.int64)
    keep = flat > 0
    PMat = sp.csr_matrix((np.ones(int(keep.sum())), (np.flatnonzero(keep), flat[keep] - 1)),
                         shape=(prob.n ** 2, P.nparts))
    wantA = np.asarray((sp.csr_matrix(prob.A) @ PMat).todense())
    wantC = np.asarray(PMat.T @ np.asarray(prob.C, dtype=np.float64)).reshape(-1)
    assert np.allclose(newA, wantA, rtol=1e-13, atol=1e-13)
    assert np.allclose(newC, wantC, rtol=1e-13, atol=1e-13)
    assert np.array_equal(newB, prob.b)
    from sdpsr_b200.api import _constraints
    cs = _constraints(P)
    assert sum(len(c) for c in cs) == int(keep.sum())
    for i in (0, P.nparts - 1):
        assert np.array_equal(cs[i], np.flatnonzero(flat == i + 1))
    P.release()


# ---------------------------------------------------------------------------------
# complex path (test/runtests.jl:43-57)
# ---------------------------------------------------------------------------------
def test_complex_path_sizes_and_characters(vec):
    c4 = vec["circulant4"]
    P = S.Partition(c4["nparts"], np.array(c4["matrix"], dtype=np.uint32))
    X = S.randomize(P, Coeffs(1))
    assert np.array_equal(X, X.T)                                   # :45
    bd = S.blockDiagonalize(P, True, complex=True, rand=Coeffs(2))
    assert bd.blkSizes == c4["complex_blkSizes"]                    # :47
    P3 = S.Partition(np.array(vec["C3"]["matrix"]))
    bd = S.blockDiagonalize(P3, False, complex=True, rand=Coeffs(4))
    assert bd.blkSizes == vec["C3"]["complex_blkSizes"]            # :57
    # the images of the cyclic shifts are the characters of Z_3: cube roots of unity
    for k in range(3):
        chars = np.array([bd.blks[i][k][0, 0] for i in range(3)])
        assert np.allclose(np.abs(chars), 1.0, atol=1e-10)
        assert np.allclose(chars ** 3, 1.0, atol=1e-9)
    cols = sorted(tuple(np.round(np.angle([bd.blks[i][k][0, 0] for i in range(3)]) / (2 * np.pi / 3)).astype(int) % 3)
                  for k in range(3))
    assert len(set(cols)) == 3                                       # three distinct characters
    so, _ = O.blockDiagonalize(O.partition_from_values(np.array(vec["C3"]["matrix"])), Coeffs(4), complex=True)
    assert list(so) == bd.blkSizes


def s3_regular_labels():
    import itertools
    perms = list(itertools.permutations(range(3)))
    idx = {p: i for i, p in enumerate(perms)}
    comp = lambda p, q: tuple(p[q[i]] for i in range(3))
    inv = lambda p: tuple(sorted(range(3), key=lambda i: p[i]))
    L = np.zeros((6, 6), dtype=np.int64)
    for a in perms:
        for b in perms:
            L[idx[a], idx[b]] = idx[comp(inv(a), b)] + 1       # class of (a,b) = a^-1 b
    return L


def test_complex_path_on_noncommutative_algebra():
    """The group algebra of S3 in its regular representation (6 classes, irreps 1 + 1 + 2x2):
    complex block sizes [1, 1, 2].  The 1x1 blocks must be characters (multiplicative); the 2x2
    block of the reference's construction is built from non-orthogonal eigenvectors of a non-normal
    element and is not multiplicative in the reference either (the oracle shows the same), so only
    its size is pinned."""
    L = s3_regular_labels()
    P = S.Partition(L)
    assert P.nparts == 6
    bd = S.blockDiagonalize(P, False, complex=True, rand=Coeffs(7))
    so, _ = O.blockDiagonalize(O.partition_from_values(L), Coeffs(7), complex=True)
    assert sorted(bd.blkSizes) == sorted(so) == [1, 1, 2]
    Pm = P.matrix.astype(np.int64)
    Bm = [(Pm == i + 1).astype(float) for i in range(6)]
    for k, s in enumerate(bd.blkSizes):
        if s != 1:
            continue
        for i in range(6):
            for j in range(6):
                prod = Bm[i] @ Bm[j]
                coeff = [prod[Bm[t] > 0][0] for t in range(6)]
                want = sum(coeff[t] * bd.blks[t][k] for t in range(6))
                assert np.allclose(bd.blks[i][k] @ bd.blks[j][k], want, atol=1e-8)


# ---------------------------------------------------------------------------------
# BASELINE.json configs[2] at full size (size-independent properties)
# ---------------------------------------------------------------------------------
def test_config3_hamming_4_8_full_size():
    """Theta' of H(4,8), N = 4096: dim 5, labels = distance + 1, blocks 5x[1] equal to the
    Krawtchouk eigenmatrix, multiplicities C(4,j) 7^j (SURVEY.md 8(d) cfg 3)."""
    prob = pr.hamming(4, 8, sparse=True)
    tr = {}
    P = S.admissible_subspace(*prob, rand=Coeffs(20260101), trace=tr)
    assert P.nparts == 5 and tr["init"] == 2 and tr["iters"] == [(2, 4), (4, 5), (5, 5)]
    assert np.array_equal(P.matrix, pr.hamming_distance_matrix(4, 8).astype(np.uint32) + 1)
    K = pr.krawtchouk(4, 8)
    for eig in ("krylov", "syevd"):
        bd = S.blockDiagonalize(P, False, rand=Coeffs(5), eig=eig)
        assert P._eig_mode == eig
        assert bd.blkSizes == [1] * 5
        mult = sorted(int(P._ptrs[r + 1] - P._ptrs[r]) for r in dict.fromkeys(P._kroot.tolist()))
        assert mult == [1, 28, 294, 1372, 2401]
        got = np.array([[bd.blks[i][k][0, 0] for k in range(5)] for i in range(5)])
        for k in range(5):
            assert min(np.abs(K - got[:, [k]]).max(axis=0)) < 1e-8 * np.abs(K).max()
    P.release()


# ---------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 15, 16, 17])
def test_tiny_and_boundary_orders(n):
    rng = np.random.default_rng(n)
    M = rng.integers(0, 3, size=(n, n)).astype(np.float64)
    want = O.partition_from_values(M)
    with B.Context(n) as ctx:
        assert ctx.refine_values(M, ATOL, True) == want.nparts
        assert np.array_equal(ctx.get_labels(), want.matrix)
        if want.nparts:
            r = rng.random(want.nparts)
            ctx.fill(r)
            assert np.array_equal(ctx.get_matrix(B.MAT_X), O.fill(want, r))
            ctx.square_round_refine(ATOL)
            X = O.fill(want, r)
            w2 = O.refine(want, O.partition_from_values(O.clamp_round(X @ X, ATOL)))
            assert np.array_equal(ctx.get_labels(), w2.matrix)


def test_all_zero_matrix_is_the_empty_partition():
    with B.Context(9) as ctx:
        assert ctx.refine_values(np.zeros((9, 9)), ATOL, True) == 0
        assert ctx.zero_count() == 81
        assert not ctx.get_labels().any()
        ctx.fill(np.zeros(0))
        assert not ctx.get_matrix(B.MAT_X).any()


def test_argument_errors():
    with pytest.raises(B.SdpsrError) as e:
        B.Context(70000)
    assert e.value.code == B.E_INVALID
    with B.Context(4) as ctx:
        with pytest.raises(B.SdpsrError) as e:
            ctx.refine_values(np.ones((4, 4)), 0.5, True)          # floor(-log10(atol)) must be >= 1
        assert e.value.code == B.E_INVALID
        with pytest.raises(B.SdpsrError) as e:
            ctx.set_labels(-np.ones((4, 4), dtype=np.int64))       # @assert 0 <= first(M_vals)
        assert e.value.code == B.E_INVALID
        with pytest.raises(B.SdpsrError) as e:
            ctx.project_round_refine(ATOL)                         # no constraints yet
        assert e.value.code == B.E_STATE
        with pytest.raises(B.SdpsrError) as e:
            ctx.square_round_refine(ATOL)                          # no X yet
        assert e.value.code == B.E_STATE
        with pytest.raises(B.SdpsrError) as e:
            ctx.set_constraints(np.zeros((2, 16)))                 # A == 0: nothing to project onto
        assert e.value.code == B.E_SINGULAR
        A = np.zeros((2, 16))
        A[0, :4] = 1.0
        A[1, :4] = 2.0                                             # dependent rows: the second one is dropped
        ctx.set_constraints(A)
        assert ctx.constraint_rank() == 1


def _with_extra_rows(prob, rows, rhs):
    import scipy.sparse as sp
    A = prob.A
    if sp.issparse(A):
        A2 = sp.vstack([A] + [sp.csr_matrix(r) for r in rows]).tocsr()
    else:
        A2 = np.vstack([A] + [np.asarray(r).reshape(1, -1) for r in rows])
    return A2, np.concatenate([prob.b, np.asarray(rhs, dtype=np.float64)])


@pytest.mark.parametrize("prob", [pr.lovasz_er(5), pr.qap_esc16j(os.path.join(GOLDEN, "esc16j.npz"))],
                         ids=lambda p: p.name)
def test_redundant_constraint_rows(prob):
    """A row that is a combination of other rows does not change the row space: same partition as the
    original problem (the reference's sparse `qr(A')` is rank-revealing, src/partitions.jl:124).  The
    engine's pivoted Cholesky of A A' must drop it instead of failing."""
    A = prob.A
    r0 = A[0] + A[1] if not hasattr(A, "tocsr") else (A.tocsr()[0] + A.tocsr()[1])
    A2, b2 = _with_extra_rows(prob, [r0, r0 * 0.5], [prob.b[0] + prob.b[1], 0.5 * (prob.b[0] + prob.b[1])])
    Po = O.admissible_subspace(prob.C, prob.A, prob.b, Coeffs(21))
    Pg = S.admissible_subspace(prob.C, A2, b2, rand=Coeffs(21))
    assert Pg._ctx.constraint_rank() == prob.A.shape[0]
    assert Pg.nparts == Po.nparts == prob.expected_dim
    assert np.array_equal(Pg.matrix, Po.matrix)
    Po2 = O.admissible_subspace(prob.C, A2, b2, Coeffs(21))      # the oracle's least-squares route agrees
    assert np.array_equal(Po2.matrix, Po.matrix)
    Pg.release()


@pytest.mark.parametrize("eps", [1e-3, 1e-5])
def test_nearly_dependent_constraint_rows(eps):
    """Rows (a0, a0 + eps*a1) span the same space as (a0, a1) with cond(A) ~ 1/eps: the projection must not
    lose cond(A)^2 digits (normal equations in double would)."""
    prob = pr.lovasz_er(7)
    A = np.asarray(prob.A, dtype=np.float64)
    A2 = np.vstack([A[0], A[0] + eps * A[1]])
    b2 = np.array([prob.b[0], prob.b[0] + eps * prob.b[1]])
    Po = O.admissible_subspace(prob.C, prob.A, prob.b, Coeffs(22))
    Pg = S.admissible_subspace(prob.C, A2, b2, rand=Coeffs(22), keep_context=False)
    assert Pg.nparts == Po.nparts == prob.expected_dim
    assert np.array_equal(Pg.matrix, Po.matrix)
