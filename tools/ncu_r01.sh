# Round-1 profiling recipe (B200_PROFILING.md): launch list of the bench command + one full capture of the top kernels.
set -x
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/r01_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/r01_ncu1.log 2>&1
$CMD > gpurun_out/r01_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"^k_scatter$|k_aggregate_cols|k_pack|k_tile_summary|k_gather_buckets" -s 5 -c 5 -o gpurun_out/r01_prof_top $CMD > gpurun_out/r01_ncu2.log 2>&1
ls -la gpurun_out | tail -8
