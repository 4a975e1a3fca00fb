#!/bin/bash
# 2 GPUs over NCCL after the KeySlot / pacing changes: the torchrun parity test and the bench line
set -u
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k torchrun > gpurun_out/r2aa_tests.log 2>&1; tail -2 gpurun_out/r2aa_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29571 bench.py --gpus 2 --steps 4 --warmup 3 --no-extras > gpurun_out/r2aa_bench_g2.json 2> gpurun_out/r2aa_bench_g2.err; tail -c 300 gpurun_out/r2aa_bench_g2.err; cut -c1-200 gpurun_out/r2aa_bench_g2.json
timeout 300 $TR --master-port 29572 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2aa_ref_g2.json 2> gpurun_out/r2aa_ref_g2.err; tail -c 200 gpurun_out/r2aa_ref_g2.err; cut -c1-300 gpurun_out/r2aa_ref_g2.json
