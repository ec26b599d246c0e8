#!/bin/bash
# data-parallel training A/B on N GPUs: flat all-reduce after the backward (eager and graphed) vs per-flow buckets
# overlapped with the backward vs the same buckets deferred to the end
mkdir -p gpurun_out
TAG=${1:-r02f}
N=${2:-2}
B=${3:-32}
if [ "${4:-}" = "tests" ]; then
  timeout 1200 python -m pytest tests/test_training.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/${TAG}_pytest_train.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_train.log; tail -5 gpurun_out/${TAG}_pytest_train.log
fi
PORT=29520
for MODE in "--reduce overlap" "--reduce deferred" "--reduce flat" "--reduce flat --graph"; do
  PORT=$((PORT+1))
  NAME=$(echo $MODE | tr -d ' -' )
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      tools/check_ddp_train.py --batch $B --steps 6 $MODE > gpurun_out/${TAG}_ddp${N}_${NAME}.log 2>&1
  echo "$MODE exit $?"; grep '^{' gpurun_out/${TAG}_ddp${N}_${NAME}.log | cut -c1-420
done
