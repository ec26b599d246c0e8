#!/bin/bash
# ncu --set full capture of the backward-pass tensor-core kernels inside one training step (after a plain run exits 0):
# two layers' worth of launches (skinny wgrad, g_acts pair GEMM, dW_res, bias-sum wgrad, dW_in, dW_cond, dilated dgrad).
mkdir -p gpurun_out
TAG=${1:-ncu_train}
SKIP=${2:-200}
COUNT=${3:-14}
CMD="python tools/bench_train.py --steps 1 --warmup 1"
timeout 600 $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'wgrad_kernel|pair_kernel' -s $SKIP -c $COUNT \
    -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-300; tail -3 gpurun_out/${TAG}_ncu_full.log
