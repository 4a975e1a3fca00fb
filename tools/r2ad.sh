#!/bin/bash
# 4 GPUs: the multi-rank test file on its own (cold cuSOLVER, loader race fixed), then the bench with and without
# the split tail (snake deal in both)
set -u
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "dense_eigen or split_tail or torchrun" > gpurun_out/r2ad_tests_multi.log 2>&1; tail -3 gpurun_out/r2ad_tests_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
for t in 1 0; do
  SDPSR_I8_TAIL=$t timeout 400 $TR --master-port 2958$t bench.py --gpus 4 --steps 4 --warmup 3 --no-extras > gpurun_out/r2ad_bench_g4_tail$t.json 2> gpurun_out/r2ad_bench_g4_tail$t.err; echo "TAIL=$t rc=$?"
  python -c "
import json
for l in open('gpurun_out/r2ad_bench_g4_tail$t.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['kernel_ms_per_step'], d['parity']['checked'])
"
done
