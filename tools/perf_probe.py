"""GPU perf probe (run under gpurun): kernel-level timings at the benchmark sizes.
Prints one JSON line per measurement; not a bench (see bench.py)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import sdpsr_b200 as S
from sdpsr_b200 import binding as B
from sdpsr_b200 import problems as pr


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


def out(**kw):
    print(json.dumps(kw), flush=True)


def cublas_dgemm(n):
    A = torch.randn(n, n, dtype=torch.float64, device="cuda")
    Bm = torch.randn(n, n, dtype=torch.float64, device="cuda")
    C = torch.empty_like(A)
    best, med = ev_time(lambda: torch.matmul(A, Bm, out=C), reps=5, warm=2)
    out(what="cublas_dgemm", n=n, ms_best=best, ms_med=med, tflops_best=2 * n ** 3 / best / 1e9)


def our_gemm(n):
    rng = np.random.default_rng(0)
    with B.Context(n, 0, B.F_TIMING) as ctx:
        A = torch.randn(n, n, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        ctx.set_matrix(B.MAT_X, A)
        ctx.set_matrix(B.MAT_Q, A)
        for _ in range(2):
            ctx.gemm(B.MAT_X, B.MAT_Q, B.MAT_X2)
        ctx.timing_reset()
        for _ in range(3):
            ctx.gemm(B.MAT_X, B.MAT_Q, B.MAT_X2)
        t = ctx.timing()["gemm"]
        ms = t["ms"] / t["launches"]
        out(what="sdpsr_gemm", n=n, ms=ms, tflops=2 * n ** 3 / ms / 1e9)
        if n <= 8192:
            Cm = torch.empty(n, n, dtype=torch.float64, device="cuda")
            ctx.lib.sdpsr_get_matrix(ctx._h, B.MAT_X2, Cm.data_ptr())
            ref = (A.T @ A.T).T   # column-major view: device buffer of torch (row-major) A is A^T
            err = ((Cm - ref).abs().max() / ref.abs().max()).item()
            out(what="sdpsr_gemm_check", n=n, relerr=err)


def refine_bw(n, nclasses):
    with B.Context(n, 0, B.F_TIMING) as ctx:
        g = torch.Generator(device="cuda").manual_seed(1)
        M = torch.randint(0, nclasses, (n, n), device="cuda", generator=g).to(torch.float64) * 0.37 + 0.11
        torch.cuda.synchronize()
        for it in range(4):
            if it == 1:
                ctx.timing_reset()
            d = ctx.refine_values(M, 1.4901161193847656e-8, True)
        t = ctx.timing()
        ms = t["refine"]["ms"] / t["refine"]["launches"]
        out(what="refine_pass", n=n, classes=d, ms=ms, gbs=16.0 * n * n / ms / 1e6,
            rank_ms=t["rank"]["ms"] / max(1, t["rank"]["launches"]))
        r = np.random.default_rng(0).random(d)
        ctx.fill(r)
        ctx.get_matrix  # noqa
        ctx.timing_reset()
        X = torch.empty(n, n, dtype=torch.float64, device="cuda")
        for _ in range(3):
            ctx.fill(r)
            ctx.lib.sdpsr_get_matrix(ctx._h, B.MAT_X, X.data_ptr())
        t = ctx.timing()["fill"]
        ms = t["ms"] / t["launches"]
        out(what="fill_pass", n=n, ms=ms, gbs=12.0 * n * n / ms / 1e6)


def full(prob, label):
    rng = np.random.default_rng(20260101)
    rand = lambda k: rng.random(int(k))
    tr = {}
    t0 = time.perf_counter()
    P = S.admissible_subspace(*prob, rand=rand, flags=B.F_TIMING, trace=tr)
    t1 = time.perf_counter()
    tim = P._ctx.timing()
    out(what="admissible_subspace", problem=label, n=prob.n, dim=P.nparts, wall_s=t1 - t0, t_init=tr["t_init"],
        iters=tr["iters"], timing={k: v for k, v in tim.items() if v["launches"]})
    P._ctx.timing_reset()
    t0 = time.perf_counter()
    bd = S.blockDiagonalize(P, False, rand=rand)
    t1 = time.perf_counter()
    tim = P._ctx.timing()
    out(what="blockDiagonalize", problem=label, n=prob.n, blocks=len(bd.blkSizes), wall_s=t1 - t0,
        timing={k: v for k, v in tim.items() if v["launches"]})
    P.release()


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "refine", "h48"]
    if "gemm" in which:
        for n in (4096, 8192, 16384):
            cublas_dgemm(n)
            our_gemm(n)
    if "refine" in which:
        for n, k in ((4096, 5), (16384, 8), (16384, 216), (16384, 5000), (16384, 200000)):
            refine_bw(n, k)
    if "h48" in which:
        t = time.perf_counter()
        prob = pr.hamming(4, 8)
        out(what="build_problem", problem="H(4,8)", s=time.perf_counter() - t)
        full(prob, "H(4,8)")
    if "h74" in which:
        t = time.perf_counter()
        prob = pr.hamming(7, 4)
        out(what="build_problem", problem="H(7,4)", s=time.perf_counter() - t)
        full(prob, "H(7,4)")


def eig_probe(sizes):
    for n in sizes:
        g = torch.Generator(device="cuda").manual_seed(1)
        L = torch.randint(1, 9, (n, n), device="cuda", generator=g)
        L = torch.minimum(L, L.T).to(torch.int64)
        Lh = L.cpu().numpy()
        with B.Context(n, 0, B.F_TIMING) as ctx:
            d = ctx.set_labels(Lh)
            r = np.random.default_rng(0).random(d)
            for rep in range(2):
                ctx.timing_reset()
                t0 = time.perf_counter()
                vals = ctx.eig(r)
                wall = time.perf_counter() - t0
                out(what="sdpsr_eig", n=n, rep=rep, wall_s=wall, eig_ms=ctx.timing()["eig"]["ms"])
        A = torch.randn(n, n, dtype=torch.float64, device="cuda")
        A = A + A.T
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            w, V = torch.linalg.eigh(A)
            torch.cuda.synchronize()
            out(what="torch_eigh", n=n, rep=rep, wall_s=time.perf_counter() - t0)
        del A, w, V


if __name__ == "__main__" and "eig" in sys.argv[1:]:
    eig_probe([int(x) for x in sys.argv[sys.argv.index("eig") + 1:]] or [4096, 8192])
