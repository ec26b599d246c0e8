#!/bin/bash
# GPU check of the training direction: kernel tests + gradient parity (tests/test_training.py), optional step timing.
mkdir -p gpurun_out
TAG=${1:-train}
KEXPR=${2:-""}
timeout 1200 python -m pytest tests/test_training.py -m gpu -q --timeout 900 -p no:cacheprovider -s ${KEXPR:+-k "$KEXPR"} > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -40 gpurun_out/${TAG}_pytest.log
if [ -f tools/bench_train.py ] && [ "${3:-}" = "bench" ]; then
  timeout 900 python tools/bench_train.py --breakdown > gpurun_out/${TAG}_bench.log 2>&1; tail -3 gpurun_out/${TAG}_bench.log
  timeout 900 python tools/bench_train.py --graph > gpurun_out/${TAG}_bench_graph.log 2>&1; tail -3 gpurun_out/${TAG}_bench_graph.log
  timeout 900 python tools/bench_train.py --graph --batch 8 > gpurun_out/${TAG}_bench_graph_b8.log 2>&1; tail -1 gpurun_out/${TAG}_bench_graph_b8.log
fi
