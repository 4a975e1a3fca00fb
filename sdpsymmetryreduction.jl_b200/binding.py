"""ctypes binding of ``libsdpsr_cuda.so`` (C ABI: ``include/sdpsr.h``).

This is the same call sequence a Julia ``ccall`` wrapper makes;
Python is the host language here only because the image has no Julia.  There is
no CPU fallback: if the shared library has not been built, or there is no CUDA
device, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# status codes (include/sdpsr.h)
OK = 0
E_INVALID, E_CUDA, E_NO_DEVICE, E_ALLOC, E_LABEL_OVERFLOW, E_NOT_SYMMETRIC = -1, -2, -3, -4, -5, -6
E_CUSOLVER, E_STATE, E_SINGULAR, E_NCCL, E_UNSUPPORTED, E_KRYLOV = -7, -8, -9, -10, -11, -12

F_FORCE_BITMAP_RANK, F_TINY_TABLE, F_NO_SMEM_CACHE, F_TIMING, F_NO_SYRK, F_NCCL_EXCHANGE = 1, 2, 4, 8, 16, 32
F_NO_I8, F_FORCE_I8 = 64, 128
MAT_X, MAT_X2, MAT_Q, MAT_W = 0, 1, 2, 3
K_REFINE, K_GEMM, K_FILL, K_PROJECT, K_RANK, K_EIG, K_BASIS, K_MISC, K_KRYLOV, K_GEMM_I8 = range(10)
K_NAMES = ["refine", "gemm", "fill", "project", "rank", "eig", "basis", "misc", "krylov", "gemm_i8"]


class DeviceCSR:
    """A constraint matrix A (m x N^2, CSR) whose column indices and values already live in device memory (or in
    pinned host memory): `rowptr` is a host int64 array, `indices_ptr` / `data_ptr` are raw pointers to `nnz`
    int32 or int64 indices and `nnz` doubles.  `keep` holds whatever owns that memory (e.g. torch tensors)."""

    def __init__(self, shape, rowptr, indices_ptr: int, data_ptr: int, index_bytes: int, keep=None):
        self.shape = tuple(shape)
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        self.indices_ptr, self.data_ptr, self.index_bytes, self.keep = int(indices_ptr), int(data_ptr), int(index_bytes), keep
        assert index_bytes in (4, 8)


def tile_deal(kind: int, n: int, nranks: int, rank: int):
    """The (tm, tn) tiles `rank` of `nranks` computes, straight from the library (host-only test hook)."""
    lib = load_library()
    cnt = _i64(0)
    assert lib.sdpsr_debug_tile_deal(kind, n, nranks, rank, None, 0, C.byref(cnt)) == OK
    out = np.zeros((max(cnt.value, 1), 2), dtype=np.int32)
    assert lib.sdpsr_debug_tile_deal(kind, n, nranks, rank, out.ctypes.data, cnt.value, C.byref(cnt)) == OK
    return [tuple(int(x) for x in row) for row in out[:cnt.value]]


def i8_schedule(n: int, nranks: int, rank: int, grid: int):
    """The work list of the single-CTA INT8 square for `rank` of `nranks` on `grid` CTAs, straight from the library
    (host-only test hook): (items, info) with items an int32 array of rows (tm, tn, kb0, kb1, slot, part, nparts, sem)
    and info = dict(nmain, nslots, nsems, kblocks)."""
    lib = load_library()
    cnt = _i64(0)
    info = np.zeros(4, dtype=np.int32)
    assert lib.sdpsr_debug_i8_schedule(n, nranks, rank, grid, None, 0, C.byref(cnt), info.ctypes.data) == OK
    out = np.zeros((max(cnt.value, 1), 8), dtype=np.int32)
    assert lib.sdpsr_debug_i8_schedule(n, nranks, rank, grid, out.ctypes.data, cnt.value, C.byref(cnt), info.ctypes.data) == OK
    return out[:cnt.value], dict(nmain=int(info[0]), nslots=int(info[1]), nsems=int(info[2]), kblocks=int(info[3]))


class LibraryNotBuilt(RuntimeError):
    pass


class SdpsrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsdpsr_cuda error {code}: {msg}")
        self.code = code
        self.msg = msg


def lib_path() -> str:
    return os.environ.get("SDPSR_LIB", os.path.join(_HERE, "libsdpsr_cuda.so"))


_i64 = C.c_int64
_p = C.c_void_p
_SIGNATURES = {
    "sdpsr_version": ([], C.c_int),
    "sdpsr_last_error": ([_p], C.c_char_p),
    "sdpsr_create": ([C.POINTER(_p), _i64, C.c_int, C.c_uint32], C.c_int),
    "sdpsr_destroy": ([_p], C.c_int),
    "sdpsr_device_count": ([C.POINTER(C.c_int)], C.c_int),
    "sdpsr_set_constraints_dense": ([_p, _i64, _p], C.c_int),
    "sdpsr_set_constraints_csr": ([_p, _i64, _p, _p, _p, C.c_int], C.c_int),
    "sdpsr_set_constraints_csr_i32": ([_p, _i64, _p, _p, _p, C.c_int], C.c_int),
    "sdpsr_set_constraints_csc": ([_p, _i64, _p, _p, _p, C.c_int], C.c_int),
    "sdpsr_constraint_patterns": ([_p, C.POINTER(_i64)], C.c_int),
    "sdpsr_constraint_rank": ([_p, C.POINTER(_i64)], C.c_int),
    "sdpsr_partition_reset": ([_p], C.c_int),
    "sdpsr_partition_set_labels": ([_p, _p, C.c_int, C.POINTER(_i64)], C.c_int),
    "sdpsr_partition_get_labels": ([_p, _p, C.c_int], C.c_int),
    "sdpsr_partition_dim": ([_p, C.POINTER(_i64)], C.c_int),
    "sdpsr_partition_zero_count": ([_p, C.POINTER(_i64)], C.c_int),
    "sdpsr_partition_is_symmetric": ([_p, C.POINTER(C.c_int)], C.c_int),
    "sdpsr_refine_values": ([_p, _p, C.c_double, C.c_int, C.POINTER(_i64)], C.c_int),
    "sdpsr_refine_labels": ([_p, _p, C.c_int, C.POINTER(_i64)], C.c_int),
    "sdpsr_init_partition": ([_p, _p, _p, C.c_double, C.c_int, C.POINTER(_i64)], C.c_int),
    "sdpsr_fill": ([_p, _p, _i64], C.c_int),
    "sdpsr_project_round_refine": ([_p, C.c_double, C.POINTER(_i64)], C.c_int),
    "sdpsr_square_round_refine": ([_p, C.c_double, C.POINTER(_i64)], C.c_int),
    "sdpsr_product_round_refine": ([_p, _p, _p, _i64, C.c_double, C.POINTER(_i64)], C.c_int),
    "sdpsr_eig": ([_p, _p, _i64, _p], C.c_int),
    "sdpsr_debug_tile_deal": ([C.c_int, _i64, C.c_int, C.c_int, _p, _i64, C.POINTER(_i64)], C.c_int),
    "sdpsr_debug_i8_schedule": ([_i64, C.c_int, C.c_int, C.c_int, _p, _i64, C.POINTER(_i64), _p], C.c_int),
    "sdpsr_stage_objective": ([_p, _p], C.c_int),
    "sdpsr_partition_get_labels_async": ([_p, _p, C.c_int], C.c_int),
    "sdpsr_partition_labels_wait": ([_p], C.c_int),
    "sdpsr_partition_constraints": ([_p, _p, _p, _i64, C.c_int], C.c_int),
    "sdpsr_block_norms": ([_p, _p, _i64, _p, _i64, _p], C.c_int),
    "sdpsr_irreducible": ([_p, _p, _i64, _p, _i64, _p, C.c_double, _p, C.POINTER(_i64)], C.c_int),
    "sdpsr_eig_krylov": ([_p, _p, _i64, _i64, C.c_double, _p, _p, C.POINTER(_i64)], C.c_int),
    "sdpsr_block_norms_krylov": ([_p, _p, _i64, _p], C.c_int),
    "sdpsr_irreducible_krylov": ([_p, _p, _i64, _p, C.c_double, _p, C.POINTER(_i64)], C.c_int),
    "sdpsr_get_qhat": ([_p, _p, _i64], C.c_int),
    "sdpsr_set_qhat": ([_p, _p, _p, _i64], C.c_int),
    "sdpsr_basis_image": ([_p, C.c_double, _p, _i64], C.c_int),
    "sdpsr_reduce_problem": ([_p, _p, _p, _p], C.c_int),
    "sdpsr_eig_complex": ([_p, _p, _i64, _p], C.c_int),
    "sdpsr_block_norms_complex": ([_p, _p, _i64, _p, _i64, _p], C.c_int),
    "sdpsr_irreducible_complex": ([_p, _p, _i64, _p, _i64, _p, C.c_double, _p, C.POINTER(_i64)], C.c_int),
    "sdpsr_get_qhat_complex": ([_p, _p, _i64], C.c_int),
    "sdpsr_basis_image_complex": ([_p, C.c_double, _p, _i64], C.c_int),
    "sdpsr_get_matrix": ([_p, C.c_int, _p], C.c_int),
    "sdpsr_set_matrix": ([_p, C.c_int, _p], C.c_int),
    "sdpsr_gemm": ([_p, C.c_int, C.c_int, C.c_int], C.c_int),
    "sdpsr_square": ([_p, C.c_int, C.c_int], C.c_int),
    "sdpsr_set_square_slices": ([_p, C.c_int], C.c_int),
    "sdpsr_timing_reset": ([_p], C.c_int),
    "sdpsr_timing_get": ([_p, C.c_int, C.POINTER(C.c_double), C.POINTER(_i64), C.POINTER(C.c_double)], C.c_int),
    "sdpsr_launch_count": ([_p, C.POINTER(_i64)], C.c_int),
    "sdpsr_set_stream": ([_p, _p], C.c_int),
    "sdpsr_comm_unique_id": ([_p], C.c_int),
    "sdpsr_comm_init": ([_p, C.c_int, C.c_int, _p], C.c_int),
    "sdpsr_comm_info": ([_p, C.POINTER(C.c_int), C.POINTER(C.c_int)], C.c_int),
    "sdpsr_comm_local_group": ([C.POINTER(_p), C.c_int], C.c_int),
    "sdpsr_comm_init_local": ([_p, _p, C.c_int], C.c_int),
}


def declared_symbols():
    return sorted(_SIGNATURES)


def load_library():
    """dlopen libsdpsr_cuda.so and declare every prototype; raises LibraryNotBuilt."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise LibraryNotBuilt(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = restype
    _LIB = lib
    return lib


def device_count() -> int:
    lib = load_library()
    c = C.c_int(0)
    lib.sdpsr_device_count(C.byref(c))
    return c.value


def _ptr(x):
    """Host ndarray -> address; torch CUDA tensor -> device address (kept resident)."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(f"unsupported buffer type {type(x)}")


def _f64(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.float64)


class Context:
    """One ``sdpsr_ctx``: the device state of an N x N problem on one GPU."""

    def __init__(self, n: int, device: int = 0, flags: int = 0):
        self.lib = load_library()
        self.n = int(n)
        self.device = device
        self.flags = int(flags)
        self._h = _p()
        st = self.lib.sdpsr_create(C.byref(self._h), self.n, device, flags)
        if st != OK:
            msg = self.lib.sdpsr_last_error(None).decode()
            self._h = None
            raise SdpsrError(st, msg)

    # -- plumbing ---------------------------------------------------------------
    def _check(self, st: int):
        if st != OK:
            raise SdpsrError(st, self.lib.sdpsr_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.sdpsr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- constraints -------------------------------------------------------------
    def set_constraints(self, A):
        """A: (m, N^2) ndarray or scipy.sparse matrix (src/partitions.jl:112)."""
        import scipy.sparse as sp
        if isinstance(A, DeviceCSR):
            m = A.shape[0]
            assert A.shape[1] == self.n * self.n
            fn = self.lib.sdpsr_set_constraints_csr_i32 if A.index_bytes == 4 else self.lib.sdpsr_set_constraints_csr
            self._check(fn(self._h, m, A.rowptr.ctypes.data, A.indices_ptr, A.data_ptr, 0))
        elif sp.issparse(A):
            A = A.tocsr()
            A.sum_duplicates()
            A.sort_indices()
            m = A.shape[0]
            assert A.shape[1] == self.n * self.n
            rp = np.ascontiguousarray(A.indptr, dtype=np.int64)
            vv = _f64(A.data)
            if A.indices.dtype == np.int32:          # scipy's default: passed as it is (no 8-byte copy of 46 M indices)
                ci = np.ascontiguousarray(A.indices)
                self._check(self.lib.sdpsr_set_constraints_csr_i32(self._h, m, rp.ctypes.data, ci.ctypes.data,
                                                                   vv.ctypes.data, 0))
            else:
                ci = np.ascontiguousarray(A.indices, dtype=np.int64)
                self._check(self.lib.sdpsr_set_constraints_csr(self._h, m, rp.ctypes.data, ci.ctypes.data,
                                                               vv.ctypes.data, 0))
        else:
            A = np.asarray(A, dtype=np.float64)
            m = A.shape[0]
            assert A.shape[1] == self.n * self.n
            Af = np.asfortranarray(A)       # Julia Matrix layout: m x N^2 column-major
            self._check(self.lib.sdpsr_set_constraints_dense(self._h, m, Af.ctypes.data))
        self.m = m
        self._constraints_set = True

    def set_constraints_csc(self, m, colptr, rowval, nzval, index_base=0):
        colptr = np.ascontiguousarray(colptr, dtype=np.int64)
        rowval = np.ascontiguousarray(rowval, dtype=np.int64)
        nzval = _f64(nzval)
        self._check(self.lib.sdpsr_set_constraints_csc(self._h, m, colptr.ctypes.data, rowval.ctypes.data,
                                                       nzval.ctypes.data, index_base))
        self.m = m

    def constraint_patterns(self) -> int:
        v = _i64(0)
        self._check(self.lib.sdpsr_constraint_patterns(self._h, C.byref(v)))
        return v.value

    def constraint_rank(self) -> int:
        """Numerical rank of A (dependent rows are dropped by the projector)."""
        v = _i64(0)
        self._check(self.lib.sdpsr_constraint_rank(self._h, C.byref(v)))
        return v.value

    # -- partition ---------------------------------------------------------------
    def reset(self):
        self._check(self.lib.sdpsr_partition_reset(self._h))

    def set_labels(self, M) -> int:
        M = np.asarray(M)
        assert M.shape == (self.n, self.n)
        if M.dtype not in (np.uint8, np.uint16, np.uint32, np.int64):
            M = M.astype(np.int64)
        Mf = np.asfortranarray(M)
        d = _i64(0)
        self._check(self.lib.sdpsr_partition_set_labels(self._h, Mf.ctypes.data, Mf.dtype.itemsize, C.byref(d)))
        return d.value

    def refine_labels(self, M) -> int:
        M = np.asarray(M)
        assert M.shape == (self.n, self.n)
        if M.dtype not in (np.uint8, np.uint16, np.uint32, np.int64):
            M = M.astype(np.int64)
        Mf = np.asfortranarray(M)
        d = _i64(0)
        self._check(self.lib.sdpsr_refine_labels(self._h, Mf.ctypes.data, Mf.dtype.itemsize, C.byref(d)))
        return d.value

    def get_labels(self, dtype=np.uint32, out=None):
        dtype = np.dtype(dtype)
        if out is None:
            out = np.empty((self.n, self.n), dtype=dtype, order="F")
            self._check(self.lib.sdpsr_partition_get_labels(self._h, out.ctypes.data, dtype.itemsize))
            return out
        self._check(self.lib.sdpsr_partition_get_labels(self._h, _ptr(out), dtype.itemsize))
        return out

    def get_labels_async(self, dtype, out):
        """Start the label export into `out` (pinned host memory for a real overlap) on the copy stream; `labels_wait`
        returns once it has arrived."""
        dtype = np.dtype(dtype)
        self._check(self.lib.sdpsr_partition_get_labels_async(self._h, _ptr(out), dtype.itemsize))
        return out

    def labels_wait(self):
        self._check(self.lib.sdpsr_partition_labels_wait(self._h))

    def stage_objective(self, Cvec):
        """Start the upload of a host objective on the copy stream (overlaps set_constraints); pass the SAME buffer to
        init_partition next."""
        self._check(self.lib.sdpsr_stage_objective(self._h, _ptr(Cvec)))

    def dim(self) -> int:
        d = _i64(0)
        self._check(self.lib.sdpsr_partition_dim(self._h, C.byref(d)))
        return d.value

    def zero_count(self) -> int:
        d = _i64(0)
        self._check(self.lib.sdpsr_partition_zero_count(self._h, C.byref(d)))
        return d.value

    def is_symmetric(self) -> bool:
        v = C.c_int(0)
        self._check(self.lib.sdpsr_partition_is_symmetric(self._h, C.byref(v)))
        return bool(v.value)

    def refine_values(self, M, atol: float, do_round: bool = True) -> int:
        """S = refine!(S, Partition(round?(M))); M ndarray (N,N) or a device tensor
        holding the column-major N x N doubles."""
        if isinstance(M, np.ndarray):
            assert M.shape == (self.n, self.n)
            M = np.asfortranarray(M, dtype=np.float64)
        d = _i64(0)
        self._check(self.lib.sdpsr_refine_values(self._h, _ptr(M), float(atol), int(do_round), C.byref(d)))
        return d.value

    def init_partition(self, Cvec, b, atol: float, snap_decimals: Optional[int] = 12) -> int:
        if isinstance(Cvec, np.ndarray):
            Cvec = _f64(Cvec).reshape(-1)
            assert Cvec.size == self.n * self.n
        b = _f64(b)
        d = _i64(0)
        snap = -1 if snap_decimals is None else int(snap_decimals)
        self._check(self.lib.sdpsr_init_partition(self._h, _ptr(Cvec), b.ctypes.data, float(atol), snap, C.byref(d)))
        return d.value

    def fill(self, values):
        values = _f64(values)
        self._check(self.lib.sdpsr_fill(self._h, values.ctypes.data, values.size))

    def project_round_refine(self, atol: float) -> int:
        d = _i64(0)
        self._check(self.lib.sdpsr_project_round_refine(self._h, float(atol), C.byref(d)))
        return d.value

    def square_round_refine(self, atol: float) -> int:
        d = _i64(0)
        self._check(self.lib.sdpsr_square_round_refine(self._h, float(atol), C.byref(d)))
        return d.value

    def product_round_refine(self, rx, ry, atol: float) -> int:
        rx, ry = _f64(rx), _f64(ry)
        assert rx.size == ry.size
        d = _i64(0)
        self._check(self.lib.sdpsr_product_round_refine(self._h, rx.ctypes.data, ry.ctypes.data, rx.size,
                                                        float(atol), C.byref(d)))
        return d.value

    # -- block diagonalisation ------------------------------------------------------
    def eig(self, r1) -> np.ndarray:
        r1 = _f64(r1)
        vals = np.empty(self.n, dtype=np.float64)
        self._check(self.lib.sdpsr_eig(self._h, r1.ctypes.data, r1.size, vals.ctypes.data))
        return vals

    def block_norms(self, r2, ptrs) -> np.ndarray:
        r2 = _f64(r2)
        ptrs = np.ascontiguousarray(ptrs, dtype=np.int64)
        ne = ptrs.size - 1
        norms = np.zeros((ne, ne), dtype=np.float64, order="F")
        self._check(self.lib.sdpsr_block_norms(self._h, r2.ctypes.data, r2.size, ptrs.ctypes.data, ptrs.size,
                                               norms.ctypes.data))
        return norms

    def irreducible(self, r3, ptrs, kroot, atol: float):
        r3 = _f64(r3)
        ptrs = np.ascontiguousarray(ptrs, dtype=np.int64)
        kroot = np.ascontiguousarray(kroot, dtype=np.int64)
        sizes = np.zeros(ptrs.size - 1, dtype=np.int64)
        nblk = _i64(0)
        self._check(self.lib.sdpsr_irreducible(self._h, r3.ctypes.data, r3.size, ptrs.ctypes.data, ptrs.size,
                                               kroot.ctypes.data, float(atol), sizes.ctypes.data, C.byref(nblk)))
        return sizes[:nblk.value].copy()

    def partition_constraints(self, index_base: int = 0):
        """``_constraints(P)`` (src/diagonalize.jl:42-50): list of ascending linear-index arrays, one per class."""
        d = self.dim()
        total = self.n * self.n - self.zero_count()
        ptr = np.zeros(d + 1, dtype=np.int64)
        idx = np.zeros(max(total, 1), dtype=np.uint32)
        self._check(self.lib.sdpsr_partition_constraints(self._h, ptr.ctypes.data, idx.ctypes.data, total, index_base))
        return [idx[ptr[i]:ptr[i + 1]] for i in range(d)]

    # -- Krylov variant (few distinct eigenvalues): raises SdpsrError(E_KRYLOV) when not applicable
    def eig_krylov(self, r1, atol: float, max_dim: int = 1024):
        """Distinct eigenvalues (ascending, clustered with `atol`) of fill(S, r1) and the dimensions of their
        eigenspaces, computed inside the module generated by one unit vector per diagonal class."""
        r1 = _f64(r1)
        max_dim = int(max(1, min(max_dim, 4096, self.n)))
        vals = np.zeros(max_dim, dtype=np.float64)
        mult = np.zeros(max_dim, dtype=np.int64)
        ne = _i64(0)
        self._check(self.lib.sdpsr_eig_krylov(self._h, r1.ctypes.data, r1.size, max_dim, float(atol),
                                              vals.ctypes.data, mult.ctypes.data, C.byref(ne)))
        return vals[:ne.value].copy(), mult[:ne.value].copy()

    def block_norms_krylov(self, r2, ne: int) -> np.ndarray:
        r2 = _f64(r2)
        norms = np.zeros((ne, ne), dtype=np.float64, order="F")
        self._check(self.lib.sdpsr_block_norms_krylov(self._h, r2.ctypes.data, r2.size, norms.ctypes.data))
        return norms

    def irreducible_krylov(self, r3, kroot, atol: float):
        r3 = _f64(r3)
        kroot = np.ascontiguousarray(kroot, dtype=np.int64)
        sizes = np.zeros(kroot.size, dtype=np.int64)
        nblk = _i64(0)
        self._check(self.lib.sdpsr_irreducible_krylov(self._h, r3.ctypes.data, r3.size, kroot.ctypes.data,
                                                      float(atol), sizes.ctypes.data, C.byref(nblk)))
        return sizes[:nblk.value].copy()

    def get_qhat(self, sizes) -> list:
        S = int(np.sum(sizes))
        buf = np.empty((self.n, S), dtype=np.float64, order="F")
        self._check(self.lib.sdpsr_get_qhat(self._h, buf.ctypes.data, buf.size))
        out, c = [], 0
        for s in sizes:
            out.append(np.ascontiguousarray(buf[:, c:c + int(s)]))
            c += int(s)
        return out

    def set_qhat(self, qhat: Sequence[np.ndarray]):
        sizes = np.array([q.shape[1] for q in qhat], dtype=np.int64)
        buf = np.asfortranarray(np.concatenate([np.asarray(q, dtype=np.float64) for q in qhat], axis=1))
        self._check(self.lib.sdpsr_set_qhat(self._h, buf.ctypes.data, sizes.ctypes.data, sizes.size))

    def basis_image(self, sizes, atol: float, dim: Optional[int] = None) -> list:
        d = self.dim() if dim is None else dim
        sq = int(sum(int(s) * int(s) for s in sizes))
        out = np.zeros(d * sq, dtype=np.float64)
        self._check(self.lib.sdpsr_basis_image(self._h, float(atol), out.ctypes.data, out.size))
        blks = []
        for i in range(d):
            row, off = [], i * sq
            for s in sizes:
                s = int(s)
                row.append(out[off:off + s * s].reshape(s, s, order="F").copy())
                off += s * s
            blks.append(row)
        return blks

    # -- complex path (interleaved re/im == numpy complex128 memory layout) ----------------------
    def eig_complex(self, r1) -> np.ndarray:
        r1 = np.ascontiguousarray(r1, dtype=np.complex128)
        vals = np.empty(self.n, dtype=np.complex128)
        self._check(self.lib.sdpsr_eig_complex(self._h, r1.ctypes.data, r1.size, vals.ctypes.data))
        return vals

    def block_norms_complex(self, r2, ptrs) -> np.ndarray:
        r2 = np.ascontiguousarray(r2, dtype=np.complex128)
        ptrs = np.ascontiguousarray(ptrs, dtype=np.int64)
        ne = ptrs.size - 1
        norms = np.zeros((ne, ne), dtype=np.float64, order="F")
        self._check(self.lib.sdpsr_block_norms_complex(self._h, r2.ctypes.data, r2.size, ptrs.ctypes.data, ptrs.size,
                                                       norms.ctypes.data))
        return norms

    def irreducible_complex(self, r3, ptrs, kroot, atol: float):
        r3 = np.ascontiguousarray(r3, dtype=np.complex128)
        ptrs = np.ascontiguousarray(ptrs, dtype=np.int64)
        kroot = np.ascontiguousarray(kroot, dtype=np.int64)
        sizes = np.zeros(ptrs.size - 1, dtype=np.int64)
        nblk = _i64(0)
        self._check(self.lib.sdpsr_irreducible_complex(self._h, r3.ctypes.data, r3.size, ptrs.ctypes.data, ptrs.size,
                                                       kroot.ctypes.data, float(atol), sizes.ctypes.data,
                                                       C.byref(nblk)))
        return sizes[:nblk.value].copy()

    def get_qhat_complex(self, sizes) -> list:
        S = int(np.sum(sizes))
        buf = np.empty((self.n, S), dtype=np.complex128, order="F")
        self._check(self.lib.sdpsr_get_qhat_complex(self._h, buf.ctypes.data, buf.size))
        out, c = [], 0
        for s in sizes:
            out.append(np.ascontiguousarray(buf[:, c:c + int(s)]))
            c += int(s)
        return out

    def basis_image_complex(self, sizes, atol: float, dim: Optional[int] = None) -> list:
        d = self.dim() if dim is None else dim
        sq = int(sum(int(s) * int(s) for s in sizes))
        out = np.zeros(d * sq, dtype=np.complex128)
        self._check(self.lib.sdpsr_basis_image_complex(self._h, float(atol), out.ctypes.data, out.size))
        blks = []
        for i in range(d):
            row, off = [], i * sq
            for s in sizes:
                s = int(s)
                row.append(out[off:off + s * s].reshape(s, s, order="F").copy())
                off += s * s
            blks.append(row)
        return blks

    def reduce_problem(self, Cvec, m: int, want_A: bool = True, want_C: bool = True):
        """(A * PMat, C' * PMat) for the context's partition (README.md:57-60)."""
        d = self.dim()
        newA = np.zeros((m, d), dtype=np.float64, order="F") if want_A else None
        newC = np.zeros(d, dtype=np.float64) if want_C else None
        if isinstance(Cvec, np.ndarray):
            Cvec = _f64(Cvec).reshape(-1)
        self._check(self.lib.sdpsr_reduce_problem(self._h, _ptr(Cvec) if want_C else None,
                                                  newA.ctypes.data if want_A else None,
                                                  newC.ctypes.data if want_C else None))
        return newA, newC

    # -- matrices / tools --------------------------------------------------------------
    def get_matrix(self, which: int) -> np.ndarray:
        out = np.empty((self.n, self.n), dtype=np.float64, order="F")
        self._check(self.lib.sdpsr_get_matrix(self._h, which, out.ctypes.data))
        return out

    def set_matrix(self, which: int, M):
        if isinstance(M, np.ndarray):
            M = np.asfortranarray(M, dtype=np.float64)
        self._check(self.lib.sdpsr_set_matrix(self._h, which, _ptr(M)))

    def gemm(self, a: int, b: int, c: int):
        self._check(self.lib.sdpsr_gemm(self._h, a, b, c))

    def square(self, method: int = 0, slices: int = 0):
        """X2 = X * X alone: method 0 = FP64 DMMA GEMM, 1 = INT8 tensor path (symmetric X only; digit
        width chosen by N), 2 / 3 = INT8 with 7-bit / 8-bit digits; slices 0 = default."""
        self._check(self.lib.sdpsr_square(self._h, method, slices))

    def set_square_slices(self, slices: int):
        self._check(self.lib.sdpsr_set_square_slices(self._h, slices))

    def timing_reset(self):
        self._check(self.lib.sdpsr_timing_reset(self._h))

    def timing(self) -> dict:
        out = {}
        for fam, name in enumerate(K_NAMES):
            ms, n, w = C.c_double(0), _i64(0), C.c_double(0)
            self._check(self.lib.sdpsr_timing_get(self._h, fam, C.byref(ms), C.byref(n), C.byref(w)))
            out[name] = {"ms": ms.value, "launches": n.value, "work": w.value}
        return out

    def set_stream(self, cuda_stream: int):
        """Run on the caller's stream (e.g. ``torch.cuda.current_stream().cuda_stream``)."""
        self._check(self.lib.sdpsr_set_stream(self._h, _p(cuda_stream)))

    def launch_count(self) -> int:
        v = _i64(0)
        self._check(self.lib.sdpsr_launch_count(self._h, C.byref(v)))
        return v.value

    # -- multi-GPU ----------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        lib = load_library()
        buf = C.create_string_buffer(128)
        st = lib.sdpsr_comm_unique_id(buf)
        if st != OK:
            raise SdpsrError(st, "ncclGetUniqueId failed / NCCL not loadable")
        return buf.raw

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self._check(self.lib.sdpsr_comm_init(self._h, nranks, rank, buf))

    @staticmethod
    def local_group(nranks: int):
        """Handle of an in-process communicator for ``nranks`` contexts (one host thread per rank)."""
        lib = load_library()
        g = _p()
        st = lib.sdpsr_comm_local_group(C.byref(g), int(nranks))
        if st != OK:
            raise SdpsrError(st, "sdpsr_comm_local_group failed")
        return g

    def comm_init_local(self, group, rank: int):
        self._check(self.lib.sdpsr_comm_init_local(self._h, group, int(rank)))

    def comm_info(self):
        a, b = C.c_int(0), C.c_int(0)
        self._check(self.lib.sdpsr_comm_info(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value
