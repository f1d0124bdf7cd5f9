# usage: bash tools/bench_configs.sh <tag> [configs...]   -- one bench line per BASELINE config on ONE GPU
# (c4 at the per-GPU share of the 8-GPU run: 25 read sets; the full 200 need 2 / 4 / 8 GPUs)
T=${1:-r02}
shift
for cfg in ${@:-c1 c3 c5 c4}; do
  extra=""
  [ "$cfg" = c4 ] && extra="--genomes ${C4_GENOMES:-25}"
  python bench.py --config $cfg $extra --steps ${STEPS:-5} --warmup 3 > gpurun_out/${T}_${cfg}_bench.json 2> gpurun_out/${T}_${cfg}_bench.err
  echo "$cfg exit $?"; tail -c 600 gpurun_out/${T}_${cfg}_bench.err; head -c 400 gpurun_out/${T}_${cfg}_bench.json; echo
done
