#!/bin/bash
# Full GPU check of HEAD: parity suite, smoke, both bench arms, then the ncu launch list and one
# --set full capture of the three tcgen05 WN kernels (gate pair kernel, residual, skip+end).
mkdir -p gpurun_out
TAG=${1:-r01c}
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.log 2>&1
timeout 900 python bench.py --breakdown > gpurun_out/${TAG}_bench.log 2>&1
# ncu: the same command first exits 0 without ncu
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1
# per WN the tcgen05 launches are g r g r ... g s (16 per flow): skip 12 -> gate, residual, gate, skip+end, gate
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pair_kernel|skip16_kernel' -s 12 -c 5 \
    -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu.log; tail -3 gpurun_out/${TAG}_smoke.log
tail -1 gpurun_out/${TAG}_bench_reference.log; tail -1 gpurun_out/${TAG}_bench.log; tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out
