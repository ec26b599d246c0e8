#!/bin/bash
# 8-GPU: the driver's bench line (graph replays, secondary configs incl. the data-parallel train step) + the DDP A/B
mkdir -p gpurun_out
TAG=${1:-r02i}
N=${2:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$N.log 2>&1
echo "bench exit $?"; grep '^{' gpurun_out/${TAG}_bench_n$N.log | cut -c1-300
PORT=29620
for MODE in "--reduce overlap" "--reduce deferred" "--reduce flat --graph"; do
  PORT=$((PORT+1))
  NAME=$(echo $MODE | tr -d ' -' )
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      tools/check_ddp_train.py --batch 32 --steps 6 $MODE > gpurun_out/${TAG}_ddp${N}_${NAME}.log 2>&1
  echo "$MODE exit $?"; grep '^{' gpurun_out/${TAG}_ddp${N}_${NAME}.log | cut -c1-420
done
