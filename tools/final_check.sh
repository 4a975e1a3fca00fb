#!/bin/bash
set -u
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; tail -3 gpurun_out/final_tests.log
python bench.py > gpurun_out/final_bench_h74.json 2> gpurun_out/final_bench_h74.err; echo "bench rc=$?"; tail -c 200 gpurun_out/final_bench_h74.err
python -c "
import json
for l in open('gpurun_out/final_bench_h74.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['workload'], d['value'], d['e2e']['value'], d['kernel_ms_per_step'], d['parity']['checked'], d['roofline']['frac'])
"
