#!/bin/bash
# 8-GPU box: the driver's bench line at N = 8 (every config, sharded), then the headline alone at N = 1 on the same box
mkdir -p gpurun_out
TAG=${1:-r02w}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n8.log 2>&1
echo "bench n8 exit $?"; grep '^{' gpurun_out/${TAG}_bench_n8.log | tail -1 | cut -c1-300; grep -o '"rank_ms_per_step": [^]]*]' gpurun_out/${TAG}_bench_n8.log
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-secondary --no-cpu-baseline --no-gpu-baseline > gpurun_out/${TAG}_bench_n1_same_box.log 2>&1
echo "bench n1 exit $?"; grep '^{' gpurun_out/${TAG}_bench_n1_same_box.log | tail -1 | cut -c1-300
