# usage: bash tools/gpu_round_n.sh <tag> <N> [c3w c3 c4 ...]  -- N-GPU tests and bench lines under torchrun (gpurun --gpus N)
T=$1; N=$2; shift; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
nvidia-smi -L | head -8
python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/${T}_n${N}_pytest.log 2>&1; tail -3 gpurun_out/${T}_n${N}_pytest.log
for cfg in ${@:-c3w c3 c4}; do
  extra=""
  [ "$cfg" = c4 ] && extra="--genomes ${C4_GENOMES:-100}"
  $TR bench.py --gpus $N --config $cfg $extra --steps ${STEPS:-5} --warmup 3 > gpurun_out/${T}_n${N}_${cfg}_bench.json 2> gpurun_out/${T}_n${N}_${cfg}_bench.err
  echo "$cfg exit $?"; tail -c 400 gpurun_out/${T}_n${N}_${cfg}_bench.err; grep -o '"value": [0-9.]*' gpurun_out/${T}_n${N}_${cfg}_bench.json | head -2
done
$TR bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/${T}_n${N}_reference.json 2> gpurun_out/${T}_n${N}_reference.err; echo "reference exit $?"; head -c 300 gpurun_out/${T}_n${N}_reference.json
