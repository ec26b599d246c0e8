#!/bin/bash
# round-2 final validation: full GPU suite, smoke, both bench arms, launch list + ncu --set full of the WN kernels (eager steps)
mkdir -p gpurun_out
TAG=${1:-r02s}
timeout 2400 python -m pytest tests -x -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/${TAG}_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/${TAG}_bench.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log; tail -4 gpurun_out/${TAG}_smoke.log
tail -1 gpurun_out/${TAG}_bench_reference.log | cut -c1-300; grep '^{' gpurun_out/${TAG}_bench.log | tail -1 | cut -c1-600; tail -5 gpurun_out/${TAG}_bench.log | grep -E "real|exit"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-secondary --no-graph"
timeout 600 $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_kernel|skip16_kernel' -s 13 -c 5 \
    -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
