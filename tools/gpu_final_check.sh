#!/bin/bash
# What the driver runs at round end, in one call: full GPU suite, smoke, both bench arms (default flags).
mkdir -p gpurun_out
TAG=${1:-final}
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.log 2>&1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log; tail -2 gpurun_out/${TAG}_smoke.log
tail -1 gpurun_out/${TAG}_bench_reference.log | cut -c1-250; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-400
