#!/bin/bash
# 8-GPU strong-scaling line (graph replays) + the eager variant for comparison
mkdir -p gpurun_out
TAG=${1:-r02b}
N=${2:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$N.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-graph --no-secondary > gpurun_out/${TAG}_bench_n${N}_eager.log 2>&1
grep '^{' gpurun_out/${TAG}_bench_n$N.log | cut -c1-400; grep -o '"rank_ms_per_step": [^]]*]' gpurun_out/${TAG}_bench_n$N.log
grep '^{' gpurun_out/${TAG}_bench_n${N}_eager.log | cut -c1-300; grep -o '"rank_ms_per_step": [^]]*]' gpurun_out/${TAG}_bench_n${N}_eager.log
