#!/bin/bash
# Run on the GPU box (gpurun): plain bench first, then the ncu passes of B200_PROFILING.md.
set -u
TAG=${1:-r1}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
MINE='regex:^(void )?(<unnamed>::)?(refine_|gemm_f64|square_i8|slice_kernel|colmax|basis_small|labmv|kr_|mirror_lower|fill_kernel|rowdot|init_elem|symmetrize|basis_partial|block_max|class_|transpose_kernel|rank_brute|bitmap_|gather_|build_lut|canonical|pattern_|relabel|symcheck|clamp_kernel|pair_|labels_|qhat_|count_zero|reduce_)'
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
tail -c 400 gpurun_out/plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -k "$MINE" -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# per job: GEMM launches 1-4 are the half (lower-triangle) squares, 5 is the full A2*Q, 6 the half Q'T
ncu --set full --clock-control none --import-source on -k regex:gemm_f64_kernel -s 1 -c 1 \
    -f -o gpurun_out/prof_gemm_${TAG} $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
# the INT8 square (dominant kernel of the default path): launches 1-4 of a job
ncu --set full --clock-control none --import-source on -k regex:square_i8_kernel -s 5 -c 1 \
    -f -o gpurun_out/prof_i8_${TAG} $CMD > gpurun_out/ncu_i8.log 2>&1
echo "i8 capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:refine_fast_kernel -s 9 -c 3 \
    -f -o gpurun_out/prof_refine_${TAG} $CMD > gpurun_out/ncu_refine.log 2>&1
echo "refine capture rc=$?"
ls -la gpurun_out/ | tail -8
