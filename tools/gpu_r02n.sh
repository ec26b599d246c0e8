#!/bin/bash
# butterfly STFT kernels: parity tests, the A/B against the tensor-core dense-basis kernels at 256 x 10 s, ncu --set full
mkdir -p gpurun_out
TAG=${1:-r02n}
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 600 -p no:cacheprovider \
    -k "fft_stft or length_sweep or stft_pair_kernels or cfg5 or test_mel_spectrogram or test_denoiser" > gpurun_out/${TAG}_pytest_fft.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/${TAG}_pytest_fft.log | cut -c1-300
timeout 600 python tools/bench_stft_ab.py --out gpurun_out/${TAG}_stft_ab.json > gpurun_out/${TAG}_stft_ab.log 2>&1
echo "stft_ab exit $?"; TAG=$TAG python - <<'PY'
import json, os
f = "gpurun_out/%s_stft_ab.json" % os.environ["TAG"]
d = json.load(open(f)) if os.path.exists(f) else {}
for k, v in d.items():
    print(k, {a: b for a, b in v.items() if not a.endswith("breakdown")} if isinstance(v, dict) else v)
PY
CMD="python tools/bench_stft_ab.py --once fft"
timeout 300 $CMD > gpurun_out/${TAG}_once.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fft_' -s 2 -c 2 -f -o gpurun_out/${TAG}_fft_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
