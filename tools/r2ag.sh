#!/bin/bash
# copy-stream overlaps (staged upload of C, asynchronous label export) + small-tail split: suite, then the bench lines
set -u
SDPSR_SKIP_SLOW=1 timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2ag_tests.log 2>&1; tail -3 gpurun_out/r2ag_tests.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2ag_bench_h74.json 2> gpurun_out/r2ag_bench_h74.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2ag_bench_h74.err
python bench.py --workload "theta-K(20,5)-N15504" --steps 4 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2ag_bench_k205.json 2> gpurun_out/r2ag_bench_k205.err; echo "k205 rc=$?"; tail -c 300 gpurun_out/r2ag_bench_k205.err
python -c "
import json
for f in ('gpurun_out/r2ag_bench_h74.json','gpurun_out/r2ag_bench_k205.json'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(d['config']['workload'], d['value'], d['e2e']['value'], d['kernel_ms_per_step'], d['parity']['checked'])
"
