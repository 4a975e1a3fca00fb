#!/bin/bash
# bench on G GPUs of one box (torchrun, NCCL): tools/bench_multi.sh G [extra bench flags]
set -u
G=${1:-8}; shift || true
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29591 bench.py --gpus $G --steps 4 --warmup 3 --no-extras "$@" > gpurun_out/bench_multi_g$G.json 2> gpurun_out/bench_multi_g$G.err; echo "rc=$?"
tail -c 400 gpurun_out/bench_multi_g$G.err
python -c "
import json
for l in open('gpurun_out/bench_multi_g$G.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['e2e']['value'], d['kernel_ms_per_step'], d['parity']['checked'])
"
