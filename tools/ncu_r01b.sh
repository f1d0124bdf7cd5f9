set -x
CMD="python tools/quick_perf.py 100 1.0 2"
$CMD > gpurun_out/r01b_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_extract_staged|k_aggregate|k_pack|k_tile_summary|k_local_sort_gather|k_msd_scatter" -s 6 -c 6 -o gpurun_out/r01b_prof $CMD > gpurun_out/r01b_ncu.log 2>&1
tail -5 gpurun_out/r01b_plain.log
ls -la gpurun_out
