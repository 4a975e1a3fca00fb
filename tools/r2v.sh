#!/bin/bash
# round-2 checkpoint after the producer pacing: full GPU suite, default bench line, one full capture of the paced square
set -u
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2v_tests.log 2>&1; tail -3 gpurun_out/r2v_tests.log
python bench.py --steps 4 --warmup 3 > gpurun_out/r2v_bench_h74.json 2> gpurun_out/r2v_bench_h74.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2v_bench_h74.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
ncu --set full --clock-control none --import-source on -k regex:square_i8_kernel -s 5 -c 1 \
    -f -o gpurun_out/prof_i8_r2v $CMD > gpurun_out/ncu_i8_r2v.log 2>&1
echo "i8 capture rc=$?"
