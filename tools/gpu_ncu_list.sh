#!/bin/bash
# ncu launch list (gpu__time_duration.sum per launch) of the default bench workload, after a plain run exits 0.
mkdir -p gpurun_out
TAG=${1:-list}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-200; wc -l gpurun_out/${TAG}_launches.csv
