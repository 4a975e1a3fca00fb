#!/bin/bash
# 16-byte key slots (one L2 sector per global lookup): parity first, then the sweep with and without the CTA cache
set -u
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_reference_suite_gpu.py tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2z_tests.log 2>&1; tail -2 gpurun_out/r2z_tests.log
for d in 8 216 1000 2000 3000 4000 20000 200000; do python tools/refine_bench.py 16384 $d; done > gpurun_out/r2z_refine_sweep.jsonl 2>&1
cat gpurun_out/r2z_refine_sweep.jsonl
echo "-- no CTA cache, joint global path from 1 class"
for d in 216 1000 2000 3000 4000; do SDPSR_REFINE_CACHE_LIMIT=0 SDPSR_REFINE_JOINT_MIN=0 python tools/refine_bench.py 16384 $d; done > gpurun_out/r2z_refine_sweep_nocache.jsonl 2>&1
cat gpurun_out/r2z_refine_sweep_nocache.jsonl
