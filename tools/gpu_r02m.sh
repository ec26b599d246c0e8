#!/bin/bash
# step-level A/B of gate_l2_hint (two graphs, alternating replays in one process) + the kernel loop at d = 1 once more
mkdir -p gpurun_out
TAG=${1:-r02m}
timeout 900 python tools/bench_step_ab.py --key gate_l2_hint --values 0,1 --rounds 4 --steps 6 --out gpurun_out/${TAG}_step_ab_gate_l2_hint.json > gpurun_out/${TAG}_step_ab.log 2>&1
echo "step_ab exit $?"; tail -12 gpurun_out/${TAG}_step_ab.log | cut -c1-300
timeout 600 python tools/bench_kernels.py --only "gate_mel padded d=1 l2_hint" --seconds 2.0 --out gpurun_out/${TAG}_l2_hint_ab.json > gpurun_out/${TAG}_l2_hint_ab.log 2>&1
echo "bench_kernels exit $?"; cut -c1-200 gpurun_out/${TAG}_l2_hint_ab.log | tail -8
