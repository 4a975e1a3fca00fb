#!/bin/bash
# K-split tail of the INT8 square + snake deal: bit-exactness first, then A/B timing, then the job
set -u
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "int8 or i8 or square" > gpurun_out/r2ab_tests_i8.log 2>&1; tail -3 gpurun_out/r2ab_tests_i8.log
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2ab_tests_multi.log 2>&1; tail -3 gpurun_out/r2ab_tests_multi.log
for t in 0 1; do
  echo "TAIL=$t"
  SDPSR_I8_TAIL=$t timeout 300 python tools/i8_check.py --small "" --big "" --time 16384 2>&1 | grep "8-bit digits\", \"S\": 7"
done
for t in 0 1; do
  SDPSR_I8_TAIL=$t timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2ab_bench_tail$t.json 2> gpurun_out/r2ab_bench_tail$t.err; echo "bench TAIL=$t rc=$?"; python -c "
import json,sys
for l in open('gpurun_out/r2ab_bench_tail$t.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['kernel_ms_per_step'], d['parity']['checked'])
"
done
