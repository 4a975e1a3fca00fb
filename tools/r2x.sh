#!/bin/bash
# 4 GPUs: where does the e2e job spend its time?  (per C-ABI call wall clock on rank 0 and on the last rank)
set -u
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29561 tools/mgpu_breakdown.py 7 4 e2e > gpurun_out/r2x_mgpu4_e2e.log 2>&1; grep '^{' gpurun_out/r2x_mgpu4_e2e.log | cut -c1-900
timeout 300 $TR --master-port 29562 tools/mgpu_breakdown.py 7 4 > gpurun_out/r2x_mgpu4_res.log 2>&1; grep '^{' gpurun_out/r2x_mgpu4_res.log | head -2 | cut -c1-900
timeout 400 $TR --master-port 29563 bench.py --gpus 4 --steps 4 --warmup 3 --no-extras > gpurun_out/r2x_bench_g4.json 2> gpurun_out/r2x_bench_g4.err; tail -c 300 gpurun_out/r2x_bench_g4.err; cut -c1-400 gpurun_out/r2x_bench_g4.json
