#!/bin/bash
# Quick GPU check after a kernel change: kernel-level parity tests, then the bench with per-entry-point breakdown.
mkdir -p gpurun_out
TAG=${1:-quick}
KEXPR=${2:-"tc_ or infer_bf16 or forward_matches"}
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x -k "$KEXPR" > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --breakdown > gpurun_out/${TAG}_bench.log 2>&1
tail -6 gpurun_out/${TAG}_pytest.log; tail -1 gpurun_out/${TAG}_bench.log
