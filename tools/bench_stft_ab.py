"""A/B of the STFT kernels at BASELINE configs[4] size (256 x 10 s waveforms): CTA-pair kernels with the unpadded
layout (csrc/stft_tc2.cu, default) vs the one-CTA kernels with the 640-bin padded layout (csrc/wn_tc.cu).

    python tools/bench_stft_ab.py [--out gpurun_out/stft_ab.json] [--once fft|pair|single]   (--once: one call each, for ncu)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import text2speech_b200 as t2s                     # noqa: E402
from text2speech_b200 import synthetic as syn      # noqa: E402
from tools.bench_configs import breakdown, timeit  # noqa: E402

DEV = torch.device("cuda:0")
STFT_FLOP = 2 * 1026 * 1024
MEL_FLOP = 2 * 80 * 513


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "stft_ab.json"))
    ap.add_argument("--once", default="")
    ap.add_argument("--l2-hints", default="", help="comma list of stft_l2_hint values to A/B on the CTA-pair kernels instead")
    args = ap.parse_args()
    cfg = syn.load_config()
    model = t2s.WaveGlow.remove_weightnorm(t2s.WaveGlow(**cfg))
    model.load_state_dict(syn.synthetic_state_dict(cfg, seed=1234, end_std=0.01))
    model = model.to(DEV).eval()
    taco = t2s.TacotronSTFT(1024, 256, 1024, 80, 22050, 0.0, 8000.0).to(DEV)
    den = t2s.Denoiser(model)
    y = syn.synthetic_waveforms(256, 220160, sr=22050, seed=5).to(DEV)
    frames = 256 * 861
    if args.once:
        taco.stft_fn.pair = den.stft.pair = args.once != "single"
        taco.stft_fn.algorithm = den.stft.algorithm = "auto" if args.once == "fft" else "gemm"
        for _ in range(2):
            taco._mel_spectrogram(y)
            den(y, 0.01)
        torch.cuda.synchronize()
        return
    out = {}
    if args.l2_hints:
        from text2speech_b200 import _lib
        for i, hint in enumerate(int(v) for v in args.l2_hints.split(",")):
            _lib.call("wgb_set_tuning", "stft_l2_hint", hint)
            rec = {"stft_l2_hint": hint}
            rec["mel_ms"], _ = timeit(lambda: taco._mel_spectrogram(y), warmup=3, iters=20)
            rec["denoiser_ms"], _ = timeit(lambda: den(y, 0.01), warmup=3, iters=20)
            out[f"run{i}_hint{hint}"] = rec
            print(json.dumps(rec), flush=True)
        _lib.call("wgb_set_tuning", "stft_l2_hint", 0)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)
        return
    from text2speech_b200 import _lib
    for name, algo, pair, fused in (("fft", "auto", True, True), ("fft_mel_16_warps", "auto", True, True),
                                    ("fft_again", "auto", True, True), ("pair", "gemm", True, True),
                                    ("pair_separate_overlap_add", "gemm", True, False), ("one_cta", "gemm", False, False)):
        _lib.call("wgb_set_tuning", "fft_mel_warps", 16 if name == "fft_mel_16_warps" else 12)
        taco.stft_fn.algorithm = den.stft.algorithm = algo
        taco.stft_fn.pair = den.stft.pair = pair
        den.stft.fused_ola = fused
        rec = {}
        med, best = timeit(lambda: taco._mel_spectrogram(y), warmup=3, iters=10)
        rec["mel_ms"] = med
        rec["mel_algorithmic_tflops"] = frames * (STFT_FLOP + MEL_FLOP) / (med * 1e-3) / 1e12
        rec["mel_breakdown"] = breakdown(lambda: taco._mel_spectrogram(y))
        med, best = timeit(lambda: den(y, 0.01), warmup=3, iters=10)
        rec["denoiser_ms"] = med
        rec["denoiser_algorithmic_tflops"] = frames * 2 * STFT_FLOP / (med * 1e-3) / 1e12
        rec["denoiser_breakdown"] = breakdown(lambda: den(y, 0.01))
        out[name] = rec
    taco.stft_fn.pair = den.stft.pair = den.stft.fused_ola = True
    taco.stft_fn.algorithm = den.stft.algorithm = "auto"
    a_fft = taco._mel_spectrogram(y)
    d_fft = den(y, 0.01)
    taco.stft_fn.algorithm = den.stft.algorithm = "gemm"
    a = taco._mel_spectrogram(y)
    d = den(y, 0.01)
    out["max_abs_mel_diff_fft_vs_pair"] = float((a - a_fft).abs().max())
    out["denoiser_snr_fft_vs_pair_db"] = float(10 * torch.log10(d.double().pow(2).sum() / (d - d_fft).double().pow(2).sum()))
    taco.stft_fn.pair = den.stft.pair = False
    out["max_abs_mel_diff"] = float((a - taco._mel_spectrogram(y)).abs().max())
    d2 = den(y, 0.01)
    out["denoiser_snr_pair_vs_one_cta_db"] = float(10 * torch.log10(d2.double().pow(2).sum() / (d - d2).double().pow(2).sum()))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
