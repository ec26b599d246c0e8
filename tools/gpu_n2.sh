#!/bin/bash
# 2-GPU check of the sharded paths of bench.py (headline + every secondary config) and of its reference arm
mkdir -p gpurun_out
TAG=${1:-r02t}
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider -k "write_only_their_output" > gpurun_out/${TAG}_pytest_guard.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest_guard.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n2.log 2>&1
echo "bench n2 exit $?"; grep '^{' gpurun_out/${TAG}_bench_n2.log | tail -1 | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_n2_reference.log 2>&1
echo "reference n2 exit $?"; grep '^{' gpurun_out/${TAG}_bench_n2_reference.log | tail -1 | cut -c1-300
