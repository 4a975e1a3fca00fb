"""Diagnostic (single process, several GPUs): cusolverMgSyevd vs cusolverDnXsyevd on a scheme-graph
element.  python tools/mg_eig_probe.py NGPU d q [T_A]"""
import ctypes as C
import json
import sys
import time

import numpy as np

# libcusolverMg 12.9 needs libcublas 12.9; torch bundles 12.8 -- load the toolkit's first
for lib in ("libcublasLt.so.12", "libcublas.so.12", "libcusolver.so.11"):
    C.CDLL("/usr/local/cuda/lib64/" + lib, mode=C.RTLD_GLOBAL)
import torch

sys.path.insert(0, ".")
from sdpsr_b200 import problems as pr

ng, d, q = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
TA = int(sys.argv[4]) if len(sys.argv) > 4 else 256
mg = C.CDLL("/usr/local/cuda/lib64/libcusolverMg.so.11")
vp, ci, i64 = C.c_void_p, C.c_int, C.c_int64

D = torch.from_numpy(pr.hamming_distance_matrix(d, q).astype(np.int64))
r = torch.from_numpy(np.random.default_rng(0).random(d + 1))
A = r[D].contiguous()          # symmetric, so row-major == column-major
N = A.shape[0]

# reference: single GPU torch eigh (cusolver syevd)
A0 = A.cuda(0)
torch.linalg.eigh(A0)
torch.cuda.synchronize()
t = time.perf_counter()
w0, V0 = torch.linalg.eigh(A0)
torch.cuda.synchronize()
t_single = time.perf_counter() - t
print(json.dumps({"solver": "syevd 1 GPU", "n": N, "s": t_single}), flush=True)
del V0

devs = (ci * ng)(*range(ng))
for a in range(ng):
    torch.cuda.set_device(a)
    for b in range(ng):
        if a != b:
            try:
                torch.cuda.cudart().cudaDeviceEnablePeerAccess(b, 0)
            except Exception:
                pass
h = vp()
assert mg.cusolverMgCreate(C.byref(h)) == 0
assert mg.cusolverMgDeviceSelect(h, ng, devs) == 0
grid = vp()
assert mg.cusolverMgCreateDeviceGrid(C.byref(grid), 1, ng, devs, 0) == 0
desc = vp()
assert mg.cusolverMgCreateMatrixDesc(C.byref(desc), i64(N), i64(N), i64(N), i64(TA), 1, grid) == 0   # CUDA_R_64F = 1

nblk = (N + TA - 1) // TA
local = []
for p in range(ng):
    cols = sum(min(TA, N - j * TA) for j in range(p, nblk, ng))
    local.append(torch.zeros(max(cols, 1), N, dtype=torch.float64, device=f"cuda:{p}"))   # (cols, N) row-major == N x cols col-major


def scatter():
    for j in range(nblk):
        p, lj = j % ng, j // ng
        w = min(TA, N - j * TA)
        local[p][lj * TA: lj * TA + w, :].copy_(A0[j * TA: j * TA + w, :])     # column block j (A symmetric)
    for p in range(ng):
        torch.cuda.synchronize(p)


arrA = (vp * ng)(*[t_.data_ptr() for t_ in local])
W = np.zeros(N)
lwork = i64(0)
st = mg.cusolverMgSyevd_bufferSize(h, 1, 0, N, arrA, 1, 1, desc, W.ctypes.data_as(vp), 1, 1, C.byref(lwork))
assert st == 0, st
works = [torch.empty(lwork.value, dtype=torch.float64, device=f"cuda:{p}") for p in range(ng)]
arrW = (vp * ng)(*[t_.data_ptr() for t_ in works])
info = ci(0)
for rep in range(2):
    scatter()
    t = time.perf_counter()
    st = mg.cusolverMgSyevd(h, 1, 0, N, arrA, 1, 1, desc, W.ctypes.data_as(vp), 1, 1, arrW, lwork, C.byref(info))
    for p in range(ng):
        torch.cuda.synchronize(p)
    dt = time.perf_counter() - t
    err = float(np.abs(np.sort(W) - w0.cpu().numpy()).max())
    print(json.dumps({"solver": f"cusolverMgSyevd {ng} GPU", "n": N, "TA": TA, "rep": rep, "s": dt, "status": st,
                      "info": info.value, "eig_err": err, "lwork": lwork.value}), flush=True)
