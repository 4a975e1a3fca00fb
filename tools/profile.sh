#!/bin/bash
# Run on the GPU box (gpurun): plain bench first, then the ncu passes of B200_PROFILING.md.
#   tools/profile.sh <tag>      -> gpurun_out/launches_<tag>.csv, gpurun_out/prof_*_<tag>.ncu-rep
set -u
TAG=${1:-r2}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
tail -c 300 gpurun_out/plain_${TAG}.log
# launch list of the whole command (every kernel, the library's and torch's peak probes alike)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
# the INT8 square (dominant kernel of the default path): launches 1-4 are the warm-up job, 5-8 the timed one
ncu --set full --clock-control none --import-source on -k regex:square_i8_kernel -s 5 -c 1 \
    -f -o gpurun_out/prof_i8_${TAG} $CMD > gpurun_out/ncu_i8_${TAG}.log 2>&1
echo "i8 capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:refine_fast_kernel -s 9 -c 2 \
    -f -o gpurun_out/prof_refine_${TAG} $CMD > gpurun_out/ncu_refine_${TAG}.log 2>&1
echo "refine capture rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:orbit_count_kernel|labmm_kernel|slice_labels_kernel|init_sym_kernel" -c 8 \
    -f -o gpurun_out/prof_small_${TAG} $CMD > gpurun_out/ncu_small_${TAG}.log 2>&1
echo "small-kernel capture rc=$?"
ls -la gpurun_out/ | tail -6
