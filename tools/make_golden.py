"""Generate the committed fixtures under tests/golden/ from the reference's own
test data.  Runs only in the build container (needs /root/reference); the
outputs are small and are committed, so nothing reads /root/reference at test
or bench time.

  tests/golden/esc16j.npz              <- test/qapdata/esc16j.dat (QAPLIB instance)
  tests/golden/numerical_issues_P.npy  <- literal 64x64 label matrix in
                                          test/numerical_issues.jl:1-66
  tests/golden/runtests_vectors.json   <- literal matrices of test/runtests.jl:22-25,40,43-53
"""
import json
import os
import re
import sys

import numpy as np

REF = os.environ.get("SDPSR_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    # --- esc16j -------------------------------------------------------------
    tok = open(os.path.join(REF, "test", "qapdata", "esc16j.dat")).read().split()
    n = int(tok[0])
    vals = np.array(tok[1:1 + 2 * n * n], dtype=np.float64).reshape(2 * n, n)
    np.savez_compressed(os.path.join(OUT, "esc16j.npz"), flowA=vals[:n], flowB=vals[n:])
    # --- numerical_issues fixture ---------------------------------------------
    src = open(os.path.join(REF, "test", "numerical_issues.jl")).read()
    body = src[src.index("P = [") + 5: src.index("]")]
    rows = [r.split() for r in body.replace("\n", " ").split(";")]
    P = np.array([[int(x) for x in r] for r in rows if r], dtype=np.int32)
    assert P.shape == (64, 64), P.shape
    assert len(np.unique(P)) == 1312
    np.save(os.path.join(OUT, "numerical_issues_P.npy"), P)
    # --- literal vectors of test/runtests.jl ------------------------------------
    vec = {
        "P1": [[1, 2, 2], [2, 3, 3], [2, 3, 3]],                  # runtests.jl:22
        "P2": [[1, 1, 2], [1, 1, 2], [1, 1, 3]],                  # :23
        "P3_coarsest_P1_P2": [[1, 2, 4], [2, 3, 5], [2, 3, 6]],   # :24-25
        "unsymmetrize_P1": {"nparts": 4, "matrix": [[1, 3, 3], [2, 4, 4], [2, 4, 4]]},  # :40
        "circulant4": {"nparts": 3,
                       "matrix": [[1, 2, 3, 2], [2, 1, 2, 3], [3, 2, 1, 2], [2, 3, 2, 1]],
                       "complex_blkSizes": [1, 1, 1]},              # :43-47
        "C3": {"matrix": [[1, 3, 2], [2, 1, 3], [3, 2, 1]],
               "real_throws": "InvalidDecompositionField",
               "complex_blkSizes": [1, 1, 1]},                      # :50-57
    }
    with open(os.path.join(OUT, "runtests_vectors.json"), "w") as fh:
        json.dump(vec, fh, indent=1)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    sys.exit(main())
