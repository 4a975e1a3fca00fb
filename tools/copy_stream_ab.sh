#!/bin/bash
# A/B of the copy stream on one box: e2e with SDPSR_COPY_STREAM=1 / 0 (H(7,4) and K(20,5)), then the two new tests
set -u
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "staged or async" > gpurun_out/copy_stream_ab_tests.log 2>&1; tail -2 gpurun_out/copy_stream_ab_tests.log
for w in "theta-H(7,4)-N16384" "theta-K(20,5)-N15504"; do
for cs in 1 0 1; do
  SDPSR_COPY_STREAM=$cs python bench.py --workload "$w" --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/copy_stream_ab_bench.json 2> gpurun_out/copy_stream_ab_bench.err || tail -c 300 gpurun_out/copy_stream_ab_bench.err
  python -c "
import json
for l in open('gpurun_out/copy_stream_ab_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print('COPY_STREAM=$cs', d['config']['workload'], 'value', round(d['value'],4), 'e2e', round(d['e2e']['value'],4), 'diff_ms', round(1e3*(d['e2e']['value']-d['value']),1))
"
done
done
