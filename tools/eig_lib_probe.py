"""Diagnostic (GPU): sdpsr_eig on scheme-graph partitions, repeated, to separate one-time
cuSOLVER initialisation from steady-state time."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from sdpsr_b200 import binding as B
from sdpsr_b200 import problems as pr

d, q = int(sys.argv[1]), int(sys.argv[2])
D = pr.hamming_distance_matrix(d, q).astype(np.int64) + 1
n = D.shape[0]
for trial in range(2):
    with B.Context(n, 0, B.F_TIMING) as ctx:
        dim = ctx.set_labels(D)
        for rep in range(3):
            r = np.random.default_rng(rep).random(dim)
            ctx.timing_reset()
            t0 = time.perf_counter()
            vals = ctx.eig(r)
            wall = time.perf_counter() - t0
            print(json.dumps({"n": n, "trial": trial, "rep": rep, "wall_s": wall, "eig_ms": ctx.timing()["eig"]["ms"],
                              "distinct": int(len(np.unique(np.round(vals, 6))))}), flush=True)
