#!/bin/bash
# round-2 A/B: L2 eviction-priority hints on the gate / residual kernels (same process, NVML power + clock), and how many
# 4-CTA clusters the GPU can hold (the B-operand multicast question of DESIGN.md section 8)
mkdir -p gpurun_out
TAG=${1:-r02k}
nvcc -arch=sm_100a tools/cluster_probe.cu -o /tmp/cluster_probe && /tmp/cluster_probe > gpurun_out/${TAG}_cluster_probe.json 2>&1
cat gpurun_out/${TAG}_cluster_probe.json
timeout 600 python tools/bench_kernels.py --only l2_hint --seconds 2.0 --out gpurun_out/${TAG}_l2_hint_ab.json > gpurun_out/${TAG}_l2_hint_ab.log 2>&1
echo "bench_kernels exit $?"
cut -c1-260 gpurun_out/${TAG}_l2_hint_ab.log | tail -30
