#!/bin/bash
# First GPU contact: diagnostics, parity tests in independent processes, short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; ls /root/reference >> gpurun_out/gpu.txt 2>&1
timeout 300 python tools/gpu_diag.py > gpurun_out/diag.log 2>&1; echo "diag exit $?" >> gpurun_out/diag.log
timeout 600 python -m pytest tests -m gpu -q --timeout 240 -p no:cacheprovider -k "sgemm or tc_" > gpurun_out/pytest_prims.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "fp32 or stft or mel or denoiser or weight_norm or fallback" > gpurun_out/pytest_fp32.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -k "not (sgemm or tc_ or fp32 or stft or mel or denoiser or weight_norm or fallback)" > gpurun_out/pytest_bf16.log 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.log 2>&1
tail -5 gpurun_out/diag.log; tail -15 gpurun_out/pytest_prims.log; tail -15 gpurun_out/pytest_fp32.log; tail -15 gpurun_out/pytest_bf16.log; tail -5 gpurun_out/smoke.log; tail -5 gpurun_out/bench.log
