"""Diagnostic (GPU): how do cuSOLVER's symmetric eigensolvers behave on the highly
degenerate spectra of association-scheme elements?  Not product code."""
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from sdpsr_b200 import problems as pr

cs = C.CDLL("libcusolver.so.11")
h = C.c_void_p()
assert cs.cusolverDnCreate(C.byref(h)) == 0
VEC, LOWER, RANGE_ALL = 1, 0, 1001
vp, ci, cd = C.c_void_p, C.c_int, C.c_double


def out(**kw):
    print(json.dumps(kw), flush=True)


def hamming_element(d, q, seed=0):
    D = torch.from_numpy(pr.hamming_distance_matrix(d, q).astype(np.int64)).cuda()
    r = torch.from_numpy(np.random.default_rng(seed).random(d + 1)).cuda()
    return r[D].contiguous()


def run_syevd(A):
    n = A.shape[0]
    A = A.clone()
    W = torch.empty(n, dtype=torch.float64, device="cuda")
    lwork = ci(0)
    cs.cusolverDnDsyevd_bufferSize(h, VEC, LOWER, n, vp(A.data_ptr()), n, vp(W.data_ptr()), C.byref(lwork))
    work = torch.empty(lwork.value, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    t = time.perf_counter()
    st = cs.cusolverDnDsyevd(h, VEC, LOWER, n, vp(A.data_ptr()), n, vp(W.data_ptr()), vp(work.data_ptr()), lwork,
                             vp(info.data_ptr()))
    torch.cuda.synchronize()
    return time.perf_counter() - t, st, int(info.item()), W, A


def run_xsyevd(A):
    n = A.shape[0]
    A = A.clone()
    W = torch.empty(n, dtype=torch.float64, device="cuda")
    params = C.c_void_p()
    cs.cusolverDnCreateParams(C.byref(params))
    R64 = 1  # CUDA_R_64F
    wd, wh = C.c_size_t(0), C.c_size_t(0)
    i64 = C.c_int64
    cs.cusolverDnXsyevd_bufferSize(h, params, VEC, LOWER, i64(n), R64, vp(A.data_ptr()), i64(n), R64,
                                   vp(W.data_ptr()), R64, C.byref(wd), C.byref(wh))
    work = torch.empty(max(1, wd.value), dtype=torch.uint8, device="cuda")
    hwork = C.create_string_buffer(max(1, wh.value))
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    t = time.perf_counter()
    st = cs.cusolverDnXsyevd(h, params, VEC, LOWER, i64(n), R64, vp(A.data_ptr()), i64(n), R64, vp(W.data_ptr()),
                             R64, vp(work.data_ptr()), wd, hwork, wh, vp(info.data_ptr()))
    torch.cuda.synchronize()
    return time.perf_counter() - t, st, int(info.item()), W, A, wd.value, wh.value


def run_syevdx(A):
    n = A.shape[0]
    A = A.clone()
    W = torch.empty(n, dtype=torch.float64, device="cuda")
    lwork, meig = ci(0), ci(0)
    cs.cusolverDnDsyevdx_bufferSize(h, VEC, RANGE_ALL, LOWER, n, vp(A.data_ptr()), n, cd(0), cd(0), 0, 0,
                                    C.byref(meig), vp(W.data_ptr()), C.byref(lwork))
    work = torch.empty(lwork.value, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    t = time.perf_counter()
    st = cs.cusolverDnDsyevdx(h, VEC, RANGE_ALL, LOWER, n, vp(A.data_ptr()), n, cd(0), cd(0), 0, 0, C.byref(meig),
                              vp(W.data_ptr()), vp(work.data_ptr()), lwork, vp(info.data_ptr()))
    torch.cuda.synchronize()
    return time.perf_counter() - t, st, int(info.item()), W, A


def run_syevj(A, tol=1e-14, sweeps=100):
    n = A.shape[0]
    A = A.clone()
    W = torch.empty(n, dtype=torch.float64, device="cuda")
    params = C.c_void_p()
    cs.cusolverDnCreateSyevjInfo(C.byref(params))
    cs.cusolverDnXsyevjSetTolerance(params, cd(tol))
    cs.cusolverDnXsyevjSetMaxSweeps(params, sweeps)
    lwork = ci(0)
    cs.cusolverDnDsyevj_bufferSize(h, VEC, LOWER, n, vp(A.data_ptr()), n, vp(W.data_ptr()), C.byref(lwork), params)
    work = torch.empty(max(1, lwork.value), dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    t = time.perf_counter()
    st = cs.cusolverDnDsyevj(h, VEC, LOWER, n, vp(A.data_ptr()), n, vp(W.data_ptr()), vp(work.data_ptr()), lwork,
                             vp(info.data_ptr()), params)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    sw = ci(0)
    cs.cusolverDnXsyevjGetSweeps(h, params, C.byref(sw))
    return dt, st, int(info.item()), W, A, sw.value


def check(A0, W, V):
    Q = V.T
    res = (A0 @ Q - Q * W[None, :]).abs().max().item()
    orth = (Q.T @ Q - torch.eye(Q.shape[0], dtype=Q.dtype, device=Q.device)).abs().max().item()
    return res, orth


if __name__ == "__main__":
    d, q = int(sys.argv[1]), int(sys.argv[2])
    A = hamming_element(d, q)
    n = A.shape[0]
    R = torch.randn(n, n, dtype=torch.float64, device="cuda")
    R = R + R.T
    for name, M in (("random", R), ("scheme", A), ("scheme+1e-10noise", A + 1e-10 * R)):
        for rep in range(2):
            dt, st, info, W, V = run_syevd(M)
            res, orth = check(M, W, V)
            out(solver="syevd", matrix=name, n=n, rep=rep, s=dt, status=st, info=info, resid=res, orth=orth)
    for name, M in (("random", R), ("scheme", A), ("scheme", A)):
        dt, st, info, W, V, wd, wh = run_xsyevd(M)
        res, orth = check(M, W, V)
        out(solver="Xsyevd", matrix=name, n=n, s=dt, status=st, info=info, resid=res, orth=orth, wdev=wd, whost=wh)
    if "--only-d" in sys.argv:
        sys.exit(0)
    for name, M in (("random", R), ("scheme", A)):
        dt, st, info, W, V = run_syevdx(M)
        res, orth = check(M, W, V)
        out(solver="syevdx", matrix=name, n=n, s=dt, status=st, info=info, resid=res, orth=orth)
    if n <= 4096:
        for name, M in (("random", R), ("scheme", A)):
            dt, st, info, W, V, sw = run_syevj(M)
            res, orth = check(M, W, V)
            out(solver="syevj", matrix=name, n=n, s=dt, status=st, info=info, sweeps=sw, resid=res, orth=orth)
