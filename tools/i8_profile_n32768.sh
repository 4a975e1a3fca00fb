#!/bin/bash
# ncu capture of the INT8 square at N = 32768 (config 5) on a random symmetric label-valued matrix
set -u
ncu --set full --clock-control none --import-source on -k "regex:square_i8" -s 5 -c 1 -f -o gpurun_out/prof_i8_n32768 \
    python tools/i8_check.py --small "" --big "" --time 32768 > gpurun_out/ncu_i8_n32768.log 2>&1
tail -2 gpurun_out/ncu_i8_n32768.log
