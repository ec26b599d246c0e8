#!/bin/bash
# programmatic dependent launch (wgb_set_tuning "pdl"): step-level A/B at batch 1 / 8 / 64 (two graphs, alternating replays),
# then the full GPU suite with it on
mkdir -p gpurun_out
TAG=${1:-r02u}
for B in 1 8 64; do
  STEPS=$(( B == 1 ? 40 : (B == 8 ? 10 : 4) ))
  timeout 600 python tools/bench_step_ab.py --key pdl --values 0,1 --batch $B --rounds 4 --steps $STEPS --out gpurun_out/${TAG}_step_ab_pdl_b${B}.json > gpurun_out/${TAG}_step_ab_pdl_b${B}.log 2>&1
  echo "step_ab b=$B exit $?"; tail -1 gpurun_out/${TAG}_step_ab_pdl_b${B}.log | cut -c1-300
done
timeout 2400 python -m pytest tests -x -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
