"""GPU micro-benchmark of the refine pass: python tools/refine_bench.py N d [flags]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from sdpsr_b200 import binding as B

n, d = int(sys.argv[1]), int(sys.argv[2])
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
with B.Context(n, 0, B.F_TIMING | flags) as ctx:
    g = torch.Generator(device="cuda").manual_seed(1)
    M = torch.randint(0, d, (n, n), device="cuda", generator=g).to(torch.float64) * 0.37 + 0.11
    torch.cuda.synchronize()
    for it in range(4):
        if it == 1:
            ctx.timing_reset()
        dim = ctx.refine_values(M, 1.4901161193847656e-8, True)
    t = ctx.timing()["refine"]
    ms = t["ms"] / t["launches"]
    print(json.dumps({"n": n, "classes": dim, "flags": flags, "ms": ms, "gbs": 16.0 * n * n / ms / 1e6}))
