#!/bin/bash
# round-2 first validation: GPU suite without -x (collect every failure and every measured SNR), smoke, both bench arms
mkdir -p gpurun_out
TAG=${1:-r02a}
rm -f gpurun_out/${TAG}_snr.tsv
WGB_SNR_LOG=gpurun_out/${TAG}_snr.tsv timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_reference.log 2>&1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/${TAG}_bench.log
timeout 900 python bench.py --no-graph --no-secondary --no-gpu-baseline --no-cpu-baseline > gpurun_out/${TAG}_bench_eager.log 2>&1
tail -15 gpurun_out/${TAG}_pytest_gpu.log; tail -4 gpurun_out/${TAG}_smoke.log
tail -1 gpurun_out/${TAG}_bench_reference.log | cut -c1-250; tail -2 gpurun_out/${TAG}_bench.log | cut -c1-3000
tail -1 gpurun_out/${TAG}_bench_eager.log | cut -c1-600
