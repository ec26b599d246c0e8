#!/bin/bash
# butterfly STFT kernels, quick loop: the STFT-family parity tests + the A/B against the dense-basis kernels
mkdir -p gpurun_out
TAG=${1:-fftq}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 600 -p no:cacheprovider \
    -k "fft_ or length_sweep or cfg5 or test_mel_spectrogram or test_denoiser" > gpurun_out/${TAG}_pytest_fft.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${TAG}_pytest_fft.log | cut -c1-200
timeout 600 python tools/bench_stft_ab.py --out gpurun_out/${TAG}_stft_ab.json > gpurun_out/${TAG}_stft_ab.log 2>&1
echo "stft_ab exit $?"; TAG=$TAG python - <<'PY'
import json, os
d = json.load(open("gpurun_out/%s_stft_ab.json" % os.environ["TAG"]))
for k, v in d.items():
    print(k, {a: round(b, 4) for a, b in v.items() if not a.endswith("breakdown")} if isinstance(v, dict) else v)
PY
