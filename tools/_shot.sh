set -x
timeout 600 python bench.py --workload c4 --parity --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_t4_c4.json 2> gpurun_out/r2b_t4_c4.err
echo "c4 rc=$?"
SOLVE3="python tools/prof_phase.py solve 128 192 3"
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on --kernel-name-base demangled -k regex:"tileop_dmma_kernel<[256]," -s 0 -c 3 \
    --metrics smsp__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active \
    -o gpurun_out/prof_tile_r02b_3d -f $SOLVE3 > gpurun_out/r2b_t4_ncu_3d.log 2>&1
echo "ncu 3d rc=$?"
tail -5 gpurun_out/r2b_t4_ncu_3d.log
