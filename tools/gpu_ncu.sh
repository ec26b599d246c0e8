#!/bin/bash
# ncu --set full capture of the tcgen05 WN kernels inside the default bench workload (after a plain run exits 0).
mkdir -p gpurun_out
TAG=${1:-ncu}
SKIP=${2:-13}
COUNT=${3:-4}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_kernel|wn_tc_kernel' -s $SKIP -c $COUNT \
    -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_plain.log | cut -c1-300; tail -3 gpurun_out/${TAG}_ncu_full.log
