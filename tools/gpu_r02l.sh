#!/bin/bash
# round-2 A/B, second pass: gate hints combined (weights evict_last +/- mel_stack evict_last +/- streaming acts stores), the
# un-composed gate kernel, the STFT pair kernels' basis tiles, and the step-level effect in bench.py
mkdir -p gpurun_out
TAG=${1:-r02l}
timeout 600 python tools/bench_kernels.py --only l2_hint --hints 0,1,5,9,13,1,0 --hint-dilation 8 --seconds 2.0 \
    --out gpurun_out/${TAG}_l2_hint_ab.json > gpurun_out/${TAG}_l2_hint_ab.log 2>&1
echo "bench_kernels exit $?"; cut -c1-200 gpurun_out/${TAG}_l2_hint_ab.log | tail -20
timeout 600 python tools/bench_stft_ab.py --l2-hints 0,1,0,1 --out gpurun_out/${TAG}_stft_l2_hint_ab.json > gpurun_out/${TAG}_stft_l2_hint_ab.log 2>&1
echo "bench_stft_ab exit $?"; tail -5 gpurun_out/${TAG}_stft_l2_hint_ab.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-secondary > gpurun_out/${TAG}_bench_hint1.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/${TAG}_bench_hint1.log | cut -c1-400
