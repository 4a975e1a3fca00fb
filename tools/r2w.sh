#!/bin/bash
# refine pass after the joint-probe rewrite: parity tests that exercise it, then the class-count sweep
set -u
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_suite_gpu.py -m gpu -q -x > gpurun_out/r2w_tests.log 2>&1; tail -2 gpurun_out/r2w_tests.log
for d in 8 216 1000 3000 20000 200000; do python tools/refine_bench.py 16384 $d; done > gpurun_out/r2w_refine_sweep.jsonl 2>&1
cat gpurun_out/r2w_refine_sweep.jsonl
