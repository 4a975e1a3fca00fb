"""Run BASELINE.json configs 3-5 at full size on one GPU and check them against known answers
(size-independent properties; the CPU oracle cannot finish these in seconds).
    python tools/big_configs.py h48 k205 h74 syn32768
Writes one JSON line per config (also to gpurun_out/configs_r1s3.jsonl)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdpsr_b200 as S
from sdpsr_b200 import binding as B
from sdpsr_b200 import problems as pr


class Coeffs:
    def __init__(self, seed=20260101):
        self.rng = np.random.default_rng(seed)

    def __call__(self, n):
        return self.rng.random(int(n))


def check_partition_equals(labels, truth, nsample=20_000_000, seed=0):
    """labels and truth induce the same partition: equal class-size multisets and a consistent
    label<->truth map on a large random sample (plus the full first column and diagonal)."""
    n = labels.shape[0]
    sl = np.sort(np.bincount(labels.reshape(-1)))
    st = np.sort(np.bincount(truth.reshape(-1).astype(np.int64)))
    sl, st = sl[sl > 0], st[st > 0]
    if sl.size != st.size or not np.array_equal(sl, st):
        return False
    rng = np.random.default_rng(seed)
    ii = np.concatenate([rng.integers(0, n, nsample), np.arange(n), np.arange(n)])
    jj = np.concatenate([rng.integers(0, n, nsample), np.zeros(n, dtype=np.int64), np.arange(n)])
    a = labels[ii, jj].astype(np.int64)
    t = truth[ii, jj].astype(np.int64)
    pairs = np.unique(a * (int(t.max()) + 1) + t)
    return pairs.size == len(np.unique(a)) == len(np.unique(t))


def run(name, prob, truth, value_check=None, fetch=True):
    rec = {"config": name, "N": prob.n, "m": int(prob.A.shape[0])}
    rand = Coeffs()
    tr = {}
    t0 = time.perf_counter()
    P = S.admissible_subspace(*prob, rand=rand, flags=B.F_TIMING, trace=tr, fetch_labels=fetch)
    rec["admissible_subspace_s"] = time.perf_counter() - t0
    rec.update({"dim": P.nparts, "expected_dim": prob.expected_dim, "init_dim": tr["init"], "iters": tr["iters"]})
    tim = P._ctx.timing()
    rec["adm_kernel_ms"] = {k: round(v["ms"], 3) for k, v in tim.items() if v["launches"]}
    g, r = tim["gemm"], tim["refine"]
    rec["gemm_tflops"] = g["work"] / g["ms"] / 1e9 if g["ms"] else None
    rec["refine_gbs"] = r["work"] / r["ms"] / 1e6 if r["ms"] else None
    gi = tim.get("gemm_i8", {"ms": 0, "work": 0, "launches": 0})
    rec["i8_square_tops"] = gi["work"] / gi["ms"] / 1e9 if gi["ms"] else None
    rec["i8_square_ms_per_launch"] = gi["ms"] / gi["launches"] if gi["launches"] else None
    P._ctx.timing_reset()
    t0 = time.perf_counter()
    bd = S.blockDiagonalize(P, False, rand=rand)
    rec["blockDiagonalize_s"] = time.perf_counter() - t0
    tim = P._ctx.timing()
    rec["blk_kernel_ms"] = {k: round(v["ms"], 3) for k, v in tim.items() if v["launches"]}
    rec["blocks_ok"] = sorted(bd.blkSizes) == prob.expected_blocks
    mult = sorted(int(P._ptrs[r_ + 1] - P._ptrs[r_]) for r_ in dict.fromkeys(P._kroot.tolist()))
    rec["mult_ok"] = mult == prob.expected_mult
    # size-independent identities for 1x1 blocks: sum_k m_k b_ik = tr(B_i), sum_k m_k b_ik^2 = tr(B_i^2)
    if all(s == 1 for s in bd.blkSizes) and fetch:
        m_k = np.array([int(P._ptrs[r_ + 1] - P._ptrs[r_]) for r_ in dict.fromkeys(P._kroot.tolist())], dtype=np.float64)
        bvals = np.array([[bd.blks[i][k][0, 0] for k in range(len(bd.blkSizes))] for i in range(P.nparts)])
        sizes = np.bincount(P.matrix.reshape(-1), minlength=P.nparts + 1)[1:].astype(np.float64)
        diag_cls = int(P.matrix[0, 0]) - 1
        tr1 = bvals @ m_k
        want1 = np.zeros(P.nparts)
        want1[diag_cls] = prob.n
        tr2 = (bvals ** 2) @ m_k                      # tr(B_i^2) = number of entries of class i
        rec["trace_identity_err"] = float(np.abs(tr1 - want1).max() / prob.n)
        rec["trace2_identity_err"] = float((np.abs(tr2 - sizes) / sizes).max())
        if value_check is not None:
            rec["closed_form_err"] = value_check(bvals)
    if fetch and truth is not None:
        rec["partition_ok"] = bool(check_partition_equals(P.matrix, truth))
    rec["ok"] = bool(rec["dim"] == prob.expected_dim and rec["blocks_ok"] and rec["mult_ok"]
                     and rec.get("partition_ok", True) and rec.get("trace_identity_err", 0) < 1e-8
                     and rec.get("trace2_identity_err", 0) < 1e-8 and rec.get("closed_form_err", 0) < 1e-8)
    P.release()
    print(json.dumps(rec), flush=True)
    with open(os.path.join("gpurun_out", "configs_r1s3.jsonl"), "a") as fh:
        fh.write(json.dumps(rec) + "\n")
    return rec


def hamming_case(d, q):
    prob = pr.hamming(d, q, sparse=True)
    D = pr.hamming_distance_matrix(d, q)
    K = pr.krawtchouk(d, q)

    def vc(bvals):   # every column of bvals must be a column of the eigenmatrix (labels 1..d+1 <-> distance 0..d)
        return float(max(np.abs(K - bvals[:, [k]]).max(axis=0).min() for k in range(d + 1)) / np.abs(K).max())
    return prob, D.astype(np.int32), vc


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    which = sys.argv[1:] or ["h48"]
    ok = True
    if "h48" in which:
        prob, D, vc = hamming_case(4, 8)
        ok &= run("cfg3 theta' H(4,8)", prob, D, vc)["ok"]
    if "k205" in which:
        prob = pr.kneser(20, 5, sparse=True)
        truth = pr.kneser_intersection_sizes(20, 5).astype(np.int32)
        ok &= run("cfg4 theta' K(20,5)", prob, truth)["ok"]
    if "h74" in which:
        prob, D, vc = hamming_case(7, 4)
        ok &= run("cfg4' theta' H(7,4)", prob, D, vc)["ok"]
    if "syn4096" in which:
        prob = pr.synthetic_product_scheme(3, 4, 32)
        ok &= run("cfg5-small synthetic 3xH(4,2) m=32", prob, prob.meta["orbitals"])["ok"]
    if "syn32768" in which:
        t = time.perf_counter()
        prob = pr.synthetic_product_scheme(3, 5, 64)
        print(json.dumps({"build_s": time.perf_counter() - t}), flush=True)
        ok &= run("cfg5 synthetic 3xH(5,2) m=64", prob, prob.meta["orbitals"])["ok"]
    sys.exit(0 if ok else 1)
