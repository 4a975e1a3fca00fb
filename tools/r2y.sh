#!/bin/bash
# refine variants after the JOINT switch (parity + sweep), then the paced INT8 square at N = 32768 under ncu
set -u
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "refine or partition or round or fill or many" > gpurun_out/r2y_tests.log 2>&1; tail -2 gpurun_out/r2y_tests.log
for d in 216 1000 2000 3000 4000 20000; do python tools/refine_bench.py 16384 $d; done > gpurun_out/r2y_refine_sweep.jsonl 2>&1
cat gpurun_out/r2y_refine_sweep.jsonl
ncu --set full --clock-control none --import-source on -k "regex:square_i8" -s 5 -c 1 -f -o gpurun_out/prof_i8_n32768_paced \
    python tools/i8_check.py --small "" --big "" --time 32768 > gpurun_out/ncu_i8_n32768_paced.log 2>&1
tail -2 gpurun_out/ncu_i8_n32768_paced.log
