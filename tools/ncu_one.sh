# usage: bash tools/ncu_one.sh <kernel-regex> <skip> <count> <out-name>   (one ncu capture of quick_perf's second build)
set -x
CMD="python tools/quick_perf.py 100 1.0 2 ${5:-0}"
$CMD > gpurun_out/$4_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$1" -s $2 -c $3 -o gpurun_out/$4 $CMD > gpurun_out/$4_ncu.log 2>&1
tail -4 gpurun_out/$4_plain.log
