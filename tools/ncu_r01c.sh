# full capture of the unit-path kernels of one resident build (the second of bench.py's builds)
set -x
T=${1:-r01_v3}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/${T}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on \
  -k regex:"k_unit_bounds|k_units_scatter|k_units_dedupe|k_units_expand|k_aggregate_cols|k_pack|k_tile_summary|k_gather_buckets" \
  -s 9 -c 9 -o gpurun_out/${T}_prof_top $CMD > gpurun_out/${T}_ncu2.log 2>&1
tail -3 gpurun_out/${T}_ncu2.log
ls -la gpurun_out | tail -5
