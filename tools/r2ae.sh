#!/bin/bash
# round-2 final check on one GPU: smoke, the whole GPU suite, the default bench line, the ncu passes
set -u
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2ae_smoke.log 2>&1; tail -1 gpurun_out/r2ae_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2ae_tests.log 2>&1; tail -3 gpurun_out/r2ae_tests.log
python bench.py > gpurun_out/r2ae_bench_h74.json 2> gpurun_out/r2ae_bench_h74.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2ae_bench_h74.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2ae_bench_ref.json 2> gpurun_out/r2ae_bench_ref.err; echo "ref rc=$?"
bash tools/profile.sh r2b > gpurun_out/r2ae_profile.log 2>&1; tail -3 gpurun_out/r2ae_profile.log
