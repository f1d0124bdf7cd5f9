set -x
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_extract -s 2 -c 2 -o gpurun_out/r01_prof_extract $CMD > gpurun_out/ncu2.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_aggregate|k_pack|k_sort_scatter|k_tile_summary" -s 4 -c 6 -o gpurun_out/r01_prof_others $CMD > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out
