"""Diagnostic (GPU): over how many coefficient draws does the module path of blockDiagonalize apply, and why not?
   python tools/krylov_seeds.py [h54|h74|h48|k104|k205|syn32768] [nseeds]"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import sdpsr_b200 as S
from sdpsr_b200 import problems as pr


class Coeffs:
    def __init__(self, seed=1):
        self.rng = np.random.default_rng(seed)

    def __call__(self, n):
        return self.rng.random(int(n))


which = sys.argv[1] if len(sys.argv) > 1 else "h54"
nseeds = int(sys.argv[2]) if len(sys.argv) > 2 else 12
prob = {"h54": lambda: pr.hamming(5, 4, sparse=True), "h74": lambda: pr.hamming(7, 4, sparse=True),
        "h48": lambda: pr.hamming(4, 8, sparse=True), "k104": lambda: pr.kneser(10, 4),
        "k205": lambda: pr.kneser(20, 5, sparse=True),
        "syn32768": lambda: pr.synthetic_product_scheme(3, 5, 64, keep_orbitals=False)}[which]()
P = S.admissible_subspace(*prob, rand=Coeffs(1), fetch_labels=False)
out = []
for seed in range(nseeds):
    try:
        bd = S.blockDiagonalize(P, False, rand=Coeffs(seed), eig="krylov")
        out.append({"seed": seed, "ok": True, "sizes": list(bd.blkSizes)})
    except Exception as e:  # noqa: BLE001
        out.append({"seed": seed, "ok": False, "why": f"{type(e).__name__}: {e}"[:300]})
    print(json.dumps(out[-1]), flush=True)
print(json.dumps({"problem": which, "applicable": sum(o["ok"] for o in out), "of": nseeds}))
