#!/bin/bash
# ncu --set full of the STFT / mel / denoiser kernels (BASELINE configs[4]: 256 x 10 s waveforms), after a plain run.
mkdir -p gpurun_out
TAG=${1:-stft}
CMD="python tools/bench_configs.py --only cfg5 --out gpurun_out/${TAG}_cfg5.json"
timeout 600 $CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'wn_tc_kernel|reflect_pad_split|istft_overlap_add' -s 6 -c 5 -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_plain.log | cut -c1-300; tail -3 gpurun_out/${TAG}_ncu.log
