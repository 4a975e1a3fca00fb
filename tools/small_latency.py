"""GPU wall time of whole jobs on the reference's own small test problems (latency regime)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import sdpsr_b200 as S
from sdpsr_b200 import problems as pr


class Coeffs:
    def __init__(self, seed=1):
        self.rng = np.random.default_rng(seed)

    def __call__(self, n):
        return self.rng.random(int(n))


probs = [pr.petersen(), pr.lovasz_er(3), pr.lovasz_er(7), pr.qap_esc16j("tests/golden/esc16j.npz"),
         pr.hamming(3, 8), pr.kneser(12, 5)]
for prob in probs:
    ts = []
    for rep in range(6):
        t0 = time.perf_counter()
        P = S.admissible_subspace(*prob, rand=Coeffs())
        t1 = time.perf_counter()
        bd = S.blockDiagonalize(P, False, rand=Coeffs(2))
        t2 = time.perf_counter()
        P.release()
        ts.append((t1 - t0, t2 - t1))
    t0 = time.perf_counter()
    Po = O.admissible_subspace(*prob, Coeffs())
    t1 = time.perf_counter()
    O.blockDiagonalize(Po, Coeffs(2))
    t2 = time.perf_counter()
    a = np.array(ts[2:])
    print(json.dumps({"problem": prob.name, "N": prob.n, "dim": P.nparts,
                      "gpu_admissible_ms": round(1e3 * float(np.median(a[:, 0])), 2),
                      "gpu_blockdiag_ms": round(1e3 * float(np.median(a[:, 1])), 2),
                      "cpu_oracle_admissible_ms": round(1e3 * (t1 - t0), 2),
                      "cpu_oracle_blockdiag_ms": round(1e3 * (t2 - t1), 2)}), flush=True)
