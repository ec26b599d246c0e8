#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02g}
N=${2:-2}
B=${3:-32}
PORT=29540
for MODE in "--reduce overlap --graph" "--reduce flat --graph" ${4:+"--reduce overlap"}; do
  PORT=$((PORT+1))
  NAME=$(echo $MODE | tr -d ' -' )
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      tools/check_ddp_train.py --batch $B --steps 6 $MODE > gpurun_out/${TAG}_ddp${N}_${NAME}.log 2>&1
  echo "$MODE exit $?"; grep '^{' gpurun_out/${TAG}_ddp${N}_${NAME}.log | cut -c1-420 || tail -5 gpurun_out/${TAG}_ddp${N}_${NAME}.log
done
tail -5 gpurun_out/${TAG}_ddp${N}_reduceoverlapgraph.log | cut -c1-300
