"""Diagnostic (GPU): per C-ABI call of one RESIDENT job, wall time against the CUDA-event kernel
time the library recorded inside it -> host-side gap of each call.
   python tools/host_gaps.py [h74|h48]"""
import json
import sys
import time
from collections import defaultdict

import numpy as np
import torch

sys.path.insert(0, ".")
import sdpsr_b200 as S
from sdpsr_b200 import binding as B
from sdpsr_b200 import problems as pr


class Coeffs:
    def __init__(self, seed=1):
        self.rng = np.random.default_rng(seed)

    def __call__(self, n):
        return self.rng.random(int(n))


which = sys.argv[1] if len(sys.argv) > 1 else "h74"
d, q = {"h74": (7, 4), "h48": (4, 8)}[which]
prob = pr.hamming(d, q, sparse=True)
N = prob.n
C_dev = torch.ones(N * N, dtype=torch.float64, device="cuda")
ctx = B.Context(N, 0, B.F_TIMING)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)

acc = defaultdict(lambda: [0, 0.0, 0.0])
skip = {"timing", "timing_reset", "_check", "close", "launch_count"}
for name in dir(B.Context):
    fn = getattr(B.Context, name)
    if callable(fn) and not name.startswith("__") and name not in skip:
        def wrap(f, nm):
            def g(self, *a, **k):
                B.Context.timing_reset(self)
                t = time.perf_counter()
                try:
                    return f(self, *a, **k)
                finally:
                    w = time.perf_counter() - t
                    kms = sum(v["ms"] for v in B.Context.timing(self).values())
                    acc[nm][0] += 1
                    acc[nm][1] += w * 1e3
                    acc[nm][2] += kms
            return g
        setattr(B.Context, name, wrap(fn, name))

for rep in range(3):
    acc.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rand = Coeffs(20260101)
    P = S.admissible_subspace(C_dev, prob.A, prob.b, rand=rand, ctx=ctx, fetch_labels=False)
    t1 = time.perf_counter()
    bd = S.blockDiagonalize(P, False, rand=rand)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    calls = {k: {"n": v[0], "wall_ms": round(v[1], 3), "kernel_ms": round(v[2], 3), "gap_ms": round(v[1] - v[2], 3)}
             for k, v in sorted(acc.items(), key=lambda kv: -(kv[1][1] - kv[1][2]))}
    in_calls = sum(v[1] for v in acc.values())
    print(json.dumps({"rep": rep, "admissible_ms": (t1 - t0) * 1e3, "blockdiag_ms": (t2 - t1) * 1e3,
                      "in_calls_ms": in_calls, "python_between_calls_ms": (t2 - t0) * 1e3 - in_calls,
                      "calls": calls}), flush=True)
