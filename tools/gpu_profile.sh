#!/bin/bash
# Profiling pass run on the GPU box through gpurun (writes everything under gpurun_out/):
#   1. plain runs of the single-phase drivers (device timings, per-family table)
#   2. ncu launch list (gpu__time_duration) of one block solve (the same kernels, shapes and order as in bench.py)
#   3. ncu --set full of the fine-level tile kernels (tensor-core form) and, optionally, of the dense kernels
# usage: tools/gpu_profile.sh <tag> [dense]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
SOLVE="python tools/prof_phase.py solve 1024 384"
DENSE="python tools/prof_phase.py dense 4096 1"
timeout 200 $SOLVE > $OUT/solve_$TAG.log 2>&1 || { tail -20 $OUT/solve_$TAG.log; exit 1; }
cat $OUT/solve_$TAG.log
# launch list: skip the set-up and the warm-up solve, take one full block solve
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"tileop|cg_|reduce_partials|coarse_dense|csr_block" -s 620 -c 640 \
    --csv --log-file $OUT/launches_solve_$TAG.csv $SOLVE > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"tileop_dmma_kernel" -s 152 -c 8 \
    -o $OUT/prof_tile_$TAG -f $SOLVE > $OUT/ncu_tile_$TAG.log 2>&1
echo "tile capture rc=$?"
if [ "$2" = "dense" ]; then
    timeout 200 python tools/prof_phase.py dense 4096 3 > $OUT/dense_$TAG.log 2>&1
    timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"potrf_panel|dgemm_kernel" -s 60 -c 10 \
        -o $OUT/prof_dense_$TAG -f $DENSE > $OUT/ncu_dense_$TAG.log 2>&1
    echo "dense capture rc=$?"
fi
ls -la $OUT | tail -8
