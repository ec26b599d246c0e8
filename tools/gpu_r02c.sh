#!/bin/bash
# STFT pair kernels: parity tests, A/B timing, ncu of the pair kernels
mkdir -p gpurun_out
TAG=${1:-r02c}
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 600 -p no:cacheprovider -k "stft or mel or denois or split3 or cfg5" > gpurun_out/${TAG}_pytest_stft.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_stft.log
tail -25 gpurun_out/${TAG}_pytest_stft.log
timeout 600 python tools/bench_stft_ab.py --out gpurun_out/${TAG}_stft_ab.json > gpurun_out/${TAG}_stft_ab.log 2>&1; tail -60 gpurun_out/${TAG}_stft_ab.log
timeout 600 python tools/bench_stft_ab.py --once pair > /dev/null 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stft_pair_kernel' -s 3 -c 3 -f -o gpurun_out/${TAG}_stft_prof python tools/bench_stft_ab.py --once pair > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
