# usage: bash tools/ncu_kern.sh <tag> <kernel-regex> <skip> <count>  -- full capture of selected kernels of bench.py's builds
set -x
T=$1
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/${T}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/${T} $CMD > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
