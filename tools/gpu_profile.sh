#!/bin/bash
# Profiling pass run on the GPU box through gpurun (writes everything under gpurun_out/):
#   1. plain runs of the single-phase drivers (device timings, per-family table)
#   2. ncu launch list (gpu__time_duration) of a short bench.py run
#   3. ncu --set full of the dominant sparse and dense kernels
# usage: tools/gpu_profile.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python tools/prof_phase.py solve 1024 384 > $OUT/solve_$TAG.log 2>&1 || { tail -20 $OUT/solve_$TAG.log; exit 1; }
python tools/prof_phase.py dense 4096 3 > $OUT/dense_$TAG.log 2>&1 || { tail -20 $OUT/dense_$TAG.log; exit 1; }
python tools/prof_phase.py gcov 1024 2 > $OUT/gcov_$TAG.log 2>&1 || { tail -20 $OUT/gcov_$TAG.log; exit 1; }
cat $OUT/solve_$TAG.log $OUT/dense_$TAG.log $OUT/gcov_$TAG.log
BENCH="python bench.py --workload mid --steps 1 --warmup 1 --no-cpu-baseline"
$BENCH > $OUT/bench_mid_plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
python tools/prof_phase.py solve 1024 384 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"csr_block_kernel|lat_block_kernel|fine_restrict|fine_prolong|cg_update|cg_pupdate|cg_pz|jacobi0" -s 200 -c 60 \
    -o $OUT/prof_solve_$TAG -f python tools/prof_phase.py solve 1024 384 > $OUT/ncu_solve_$TAG.log 2>&1
echo "solve capture rc=$?"
python tools/prof_phase.py dense 4096 1 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"potrf_panel|dgemm_kernel|grad_reduce|kernel_matrix" -s 40 -c 30 \
    -o $OUT/prof_dense_$TAG -f python tools/prof_phase.py dense 4096 1 > $OUT/ncu_dense_$TAG.log 2>&1
echo "dense capture rc=$?"
ls -la $OUT
