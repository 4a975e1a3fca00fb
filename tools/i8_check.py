"""GPU check of the INT8 tensor-path square (csrc/gemm_i8.cu) against an exact host model.

The device kernel is integer-exact, so for small N the host can reproduce it BIT FOR BIT:
slice X into balanced base-128 digits, form P_c = sum_{s+t=c} D_s D_t with dgemm on
integer-valued doubles (exact below 2^53), and fold sum_c 2^(2e-12-7c) P_c from the smallest
weight up.  For large N the INT8 result is compared with the FP64 DMMA GEMM and timed.

    python tools/i8_check.py [--small 64,200,384] [--big 2048,4096] [--time 16384]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdpsr_b200  # noqa: E402,F401  (import shim at the repo root)
from sdpsr_b200 import binding  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from i8_model import exact_square  # noqa: E402


def sym_random(n: int, rng, kind: str) -> np.ndarray:
    if kind == "lut":          # what the closure loop squares: d distinct uniform values on a symmetric pattern
        d = 7
        lab = rng.integers(0, d + 1, size=(n, n))
        lab = np.triu(lab) + np.triu(lab, 1).T
        lut = np.concatenate([[0.0], rng.random(d)])
        return np.asfortranarray(lut[lab])
    A = rng.standard_normal((n, n)) * np.exp(rng.uniform(-6, 2, size=(n, n)))
    return np.asfortranarray(np.triu(A) + np.triu(A, 1).T)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", default="1,15,64,130,200,384")
    ap.add_argument("--big", default="2048")
    ap.add_argument("--time", default="")
    ap.add_argument("--slices", default="8,7,5,2")
    args = ap.parse_args()
    rng = np.random.default_rng(7)
    ok = True
    res = []
    for n in [int(v) for v in args.small.split(",") if v]:
        for kind in ("lut", "wide"):
            X = sym_random(n, rng, kind)
            with binding.Context(n, flags=binding.F_TIMING) as ctx:
                ctx.set_matrix(binding.MAT_X, X)
                for S in [int(v) for v in args.slices.split(",")]:
                    bits = 7 if S == 8 else 8
                    ctx.square(2 if bits == 7 else 3, S)
                    got = ctx.get_matrix(binding.MAT_X2)
                    want = exact_square(X, S, bits)
                    exact = bool(np.array_equal(got, want))
                    ref = X @ X
                    rel = float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300))
                    symm = bool(np.array_equal(got, got.T))
                    ok &= exact and symm
                    res.append({"n": n, "kind": kind, "S": S, "bit_exact": exact, "symmetric": symm, "rel_vs_f64": rel})
                    print(json.dumps(res[-1]), flush=True)
    for n in [int(v) for v in args.big.split(",") if v]:
        X = sym_random(n, rng, "lut")
        with binding.Context(n, flags=binding.F_TIMING) as ctx:
            ctx.set_matrix(binding.MAT_X, X)
            ctx.square(0, 0)
            ref = ctx.get_matrix(binding.MAT_X2)
            for S in (8, 6):
                ctx.square(2, S)
                got = ctx.get_matrix(binding.MAT_X2)
                rel = float(np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
                symm = bool(np.array_equal(got, got.T))
                good = rel < (1e-13 if S == 8 else 1e-9) and symm
                ok &= good
                res.append({"n": n, "S": S, "rel_vs_dmma": rel, "symmetric": symm, "ok": good})
                print(json.dumps(res[-1]), flush=True)
    for n in [int(v) for v in args.time.split(",") if v]:
        X = sym_random(n, rng, "lut")
        with binding.Context(n, flags=binding.F_TIMING) as ctx:
            ctx.set_matrix(binding.MAT_X, X)
            del X
            combos = ((0, 0), (2, 8), (3, 7), (3, 6), (3, 5), (3, 4)) if n <= 16384 else ((2, 8), (3, 7))
            for method, S in combos:
                for _ in range(2):
                    ctx.square(method, S)
                ctx.timing_reset()
                reps = 3
                t0 = time.perf_counter()
                for _ in range(reps):
                    ctx.square(method, S)
                wall = (time.perf_counter() - t0) / reps
                t = ctx.timing()
                fam = "gemm_i8" if method else "gemm"
                ms = t[fam]["ms"] / reps
                work = t[fam]["work"] / reps
                res.append({"n": n, "method": {0: "dmma", 2: "i8 7-bit digits", 3: "i8 8-bit digits"}[method], "S": S, "kernel_ms": ms, "wall_ms": wall * 1e3,
                            "tera_ops_per_s": work / ms / 1e9, "misc_ms": t["misc"]["ms"] / reps})
                print(json.dumps(res[-1]), flush=True)
    print("I8_CHECK", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
