"""Multi-GPU parity check, run under torchrun (one rank per GPU):
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py
Every rank runs the sharded path and compares with the CPU oracle; rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import sdpsr_b200 as S
from sdpsr_b200 import binding as B
from sdpsr_b200 import problems as pr


class Coeffs:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def __call__(self, n):
        return self.rng.random(int(n))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    results = []
    probs = [pr.lovasz_er(7), pr.qap_esc16j(os.path.join("tests", "golden", "esc16j.npz")), pr.kneser(10, 4),
             pr.hamming(3, 8), pr.synthetic_product_scheme(3, 3, 16), pr.hamming(5, 4)]
    ok_all = True
    flags = int(os.environ.get("SDPSR_MGPU_FLAGS", "0"))      # 32 = exchange with NCCL broadcasts instead of peer stores
    for prob in probs:
        ctx = B.Context(prob.n, local, flags)
        box = [B.Context.comm_unique_id() if rank == 0 else None]     # one NCCL id per communicator
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(world, rank, box[0])
        assert ctx.comm_info() == (world, rank)
        Pg = S.admissible_subspace(*prob, rand=Coeffs(11), ctx=ctx)
        bd = S.blockDiagonalize(Pg, False, rand=Coeffs(12))
        Po = O.admissible_subspace(*prob, Coeffs(11))
        so, bo = O.blockDiagonalize(Po, Coeffs(12))
        same_labels = bool(np.array_equal(Pg.matrix, Po.matrix))
        same_sizes = list(bd.blkSizes) == list(so)
        err = 0.0
        if same_sizes:
            err = max(float(np.abs(bd.blks[i][k] - bo[i][k]).max() / max(1.0, np.abs(bo[i][k]).max()))
                      for i in range(Po.nparts) for k in range(len(so)))
        ok = same_labels and same_sizes and err < 1e-8 and Pg.nparts == prob.expected_dim
        # every rank must hold the same labels
        t = torch.from_numpy(np.ascontiguousarray(Pg.matrix.astype(np.int64))).cuda()
        t0 = t.clone()
        dist.broadcast(t0, src=0)
        ok = ok and bool(torch.equal(t, t0))
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok_all = ok_all and bool(flag.item())
        results.append({"problem": prob.name, "n": prob.n, "dim": Pg.nparts, "labels": same_labels,
                        "sizes": same_sizes, "blk_err": err, "all_ranks_ok": bool(flag.item())})
        ctx.close()
    if rank == 0:
        print(json.dumps({"world": world, "flags": flags, "ok": ok_all, "results": results}), flush=True)
    dist.destroy_process_group()
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
