"""Diagnostic (GPU): wall time of every C-ABI call of one job, to find host-side overhead."""
import json
import sys
import time
from collections import defaultdict

import numpy as np

sys.path.insert(0, ".")
import sdpsr_b200 as S
from sdpsr_b200 import binding as B
from sdpsr_b200 import problems as pr

acc = defaultdict(lambda: [0, 0.0])
for name in dir(B.Context):
    fn = getattr(B.Context, name)
    if callable(fn) and not name.startswith("__") and name not in ("_check",):
        def wrap(f, nm):
            def g(*a, **k):
                t = time.perf_counter()
                try:
                    return f(*a, **k)
                finally:
                    acc[nm][0] += 1
                    acc[nm][1] += time.perf_counter() - t
            return g
        setattr(B.Context, name, wrap(fn, name))
_init = B.Context.__init__


def init(self, *a, **k):
    t = time.perf_counter()
    _init(self, *a, **k)
    acc["__init__"][0] += 1
    acc["__init__"][1] += time.perf_counter() - t


B.Context.__init__ = init


class Coeffs:
    def __init__(self, seed=1):
        self.rng = np.random.default_rng(seed)

    def __call__(self, n):
        return self.rng.random(int(n))


which = sys.argv[1] if len(sys.argv) > 1 else "h74"
t = time.perf_counter()
prob = {"h74": lambda: pr.hamming(7, 4, sparse=True), "k205": lambda: pr.kneser(20, 5, sparse=True), "h48": lambda: pr.hamming(4, 8, sparse=True),
        "syn4096": lambda: pr.synthetic_product_scheme(3, 4, 32, keep_orbitals=False),
        "syn32768": lambda: pr.synthetic_product_scheme(3, 5, 64, keep_orbitals=False)}[which]()
print(json.dumps({"build_s": time.perf_counter() - t}))
for rep in range(3):
    acc.clear()
    t0 = time.perf_counter()
    P = S.admissible_subspace(*prob, rand=Coeffs(), fetch_labels=False)
    t1 = time.perf_counter()
    bd = S.blockDiagonalize(P, False, rand=Coeffs(2))
    t2 = time.perf_counter()
    P.release()
    t3 = time.perf_counter()
    print(json.dumps({"rep": rep, "admissible_s": t1 - t0, "blockdiag_s": t2 - t1, "release_s": t3 - t2,
                      "calls": {k: [v[0], round(v[1], 4)] for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1])}}))
