"""Diagnostic (multi-GPU, torchrun): wall time per C-ABI call of one sharded job.
   torchrun ... tools/mgpu_breakdown.py d q [e2e]     (e2e: host-resident C from pinned memory, labels fetched by rank 0)"""
import json
import os
import sys
import time
from collections import defaultdict

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdpsr_b200 as S
from sdpsr_b200 import binding as B
from sdpsr_b200 import problems as pr

acc = defaultdict(lambda: [0, 0.0])
for name in ("set_constraints", "init_partition", "fill", "project_round_refine", "square_round_refine", "eig",
             "block_norms", "irreducible", "basis_image", "get_labels"):
    def wrap(f, nm):
        def g(*a, **k):
            t = time.perf_counter()
            try:
                return f(*a, **k)
            finally:
                acc[nm][0] += 1
                acc[nm][1] += time.perf_counter() - t
        return g
    setattr(B.Context, name, wrap(getattr(B.Context, name), name))


class Coeffs:
    def __init__(self, seed=1):
        self.rng = np.random.default_rng(seed)

    def __call__(self, n):
        return self.rng.random(int(n))


rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d, q = (int(x) for x in (sys.argv[1:3] if len(sys.argv) > 2 else (6, 4)))
e2e = len(sys.argv) > 3 and sys.argv[3] == "e2e"
prob = pr.hamming(d, q, sparse=True)
labels_pinned = None
if e2e:
    Cp = torch.from_numpy(np.ascontiguousarray(prob.C)).pin_memory()
    prob.C = Cp.numpy()
    lp = torch.empty(prob.n * prob.n, dtype=torch.int16).pin_memory()
    labels_pinned = lp.numpy().view(np.uint16).reshape(prob.n, prob.n, order="F")
for flags in ((0,) if e2e else (0, B.F_NCCL_EXCHANGE)):
    ctx = B.Context(prob.n, local, B.F_TIMING | flags)
    box = [B.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(world, rank, box[0])
    for rep in range(3):
        acc.clear()
        ctx.timing_reset()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if e2e:
            P = S.admissible_subspace(prob.C, prob.A, prob.b, rand=Coeffs(), ctx=ctx, labels_out=labels_pinned,
                                      label_dtype=np.uint16, fetch_labels=rank == 0)
        else:
            P = S.admissible_subspace(*prob, rand=Coeffs(), ctx=ctx, fetch_labels=False)
        t1 = time.perf_counter()
        bd = S.blockDiagonalize(P, False, rand=Coeffs(2))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if rep and (rank == 0 or (e2e and rank == world - 1)):
            print(json.dumps({"rank": rank, "e2e": e2e, "flags": flags, "rep": rep, "wall_s": round(wall, 4),
                              "admissible_s": round(t1 - t0, 4),
                              "calls": {k: [v[0], round(v[1], 4)] for k, v in acc.items()},
                              "kernels_ms": {k: round(v["ms"], 2) for k, v in ctx.timing().items() if v["launches"]}}),
                  flush=True)
    ctx.close()
dist.destroy_process_group()
