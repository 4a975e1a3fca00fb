#!/bin/bash
# A/B of the producer pacing of the INT8 square (SDPSR_I8_PACE, SDPSR_I8_PACE_KB) at N = 16384 and 32768
set -u
python tools/i8_check.py --small "64,200,384" --big "2048" > gpurun_out/i8_pace_check.log 2>&1; tail -1 gpurun_out/i8_pace_check.log
for kb in 256 64 16 4; do
  for n in 16384 32768; do
    echo "PACE_KB=$kb N=$n"
    SDPSR_I8_PACE_KB=$kb timeout 300 python tools/i8_check.py --small "" --big "" --time $n 2>&1 | grep "8-bit digits\", \"S\": 7"
  done
done
