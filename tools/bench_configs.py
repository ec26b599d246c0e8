"""Secondary measurements for the non-headline BASELINE.json configs (cfg2, cfg4, cfg5) on one B200.

    python tools/bench_configs.py [--out gpurun_out/configs.json]

CUDA-event timing, >= 3 warm-ups, median of 5; every call goes through the public module API.
The headline (cfg3, batch 64) lives in bench.py.
"""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import text2speech_b200 as t2s                     # noqa: E402
from text2speech_b200 import synthetic as syn      # noqa: E402

DEV = torch.device("cuda:0")
WN_FLOP_PER_STEP = 522190848
STFT_FLOP_PER_FRAME = 2 * 1026 * 1024


def timeit(fn, warmup=3, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return statistics.median(ms), min(ms)


def breakdown(fn):
    """Per-C-ABI-entry-point CUDA-event totals of one call of fn (after it has been warmed up)."""
    from text2speech_b200 import _lib
    raw, events = _lib.call, []

    def timed_call(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        raw(name, *a)
        e1.record()
        events.append((name, e0, e1))

    _lib.call = timed_call
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        _lib.call = raw
    agg = {}
    for name, a, b in events:
        n, ms = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, ms + a.elapsed_time(b))
    return {k: {"launches": n, "ms": round(ms, 3)} for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--only", default="", help="'cfg5': skip the WaveGlow configs (used for the ncu capture of the STFT kernels)")
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}
    cfg = syn.load_config()
    model = t2s.WaveGlow.remove_weightnorm(t2s.WaveGlow(**cfg))
    model.load_state_dict(syn.synthetic_state_dict(cfg, seed=1234, end_std=0.01))
    model = model.to(DEV).eval()
    model.mode = "bf16"
    out = []

    if args.only != "cfg5":
        wave_glow_configs(model, peaks, out)

    # ---- cfg5: mel + denoiser on 256 x 10 s waveforms
    taco = t2s.TacotronSTFT(1024, 256, 1024, 80, 22050, 0.0, 8000.0).to(DEV)
    stft_configs(model, taco, peaks, out)

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    for r in out:
        print(json.dumps(r))


def wave_glow_configs(model, peaks, out):
    # ---- cfg2: single 10 s utterance, latency + RTF
    mel = syn.synthetic_mel(1, 860, seed=0).to(DEV)
    z = syn.synthetic_z(1, 860, seed=2024).to(DEV)
    med, best = timeit(lambda: model.infer(mel, sigma=0.666, z=z))
    samples = 860 * 256
    rec = {"config": "cfg2: WaveGlow.infer 1 x 80x860 mel (10 s), bf16", "ms_median": med, "ms_best": best,
           "samples_per_s": samples / (med * 1e-3), "rtf": (med * 1e-3) / (samples / 22050),
           "wn_gemm_tflops": WN_FLOP_PER_STEP * 860 * 32 / (med * 1e-3) / 1e12}
    rec["frac_bf16_burst"] = rec["wn_gemm_tflops"] / peaks["bf16_tflops"]
    # same call replayed from a CUDA graph (removes the ~210 host launches from the critical path)
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            model.infer(mel, sigma=0.666, z=z)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            audio_g = model.infer(mel, sigma=0.666, z=z)
        gmed, gbest = timeit(g.replay)
        ref = model.infer(mel, sigma=0.666, z=z)
        rec.update({"graph_ms_median": gmed, "graph_ms_best": gbest, "graph_rtf": (gmed * 1e-3) / (samples / 22050),
                    "graph_matches_eager": bool(torch.equal(ref, audio_g))})
    except Exception as e:  # noqa: BLE001
        rec["graph_error"] = repr(e)[:300]
    out.append(rec)

    # ---- cfg4: forward direction, batch 32 x 16000 samples
    g = torch.Generator().manual_seed(1)
    wav = (0.1 * torch.randn(32, 16000, generator=g)).clamp(-1, 1).to(DEV)
    taco = t2s.TacotronSTFT(1024, 256, 1024, 80, 22050, 0.0, 8000.0).to(DEV)
    mel4 = taco.mel_spectrogram(wav)                      # [32, 80, 63]
    med, best = timeit(lambda: model((mel4, wav)))
    out.append({"config": "cfg4: WaveGlow.forward 32 x 16000 samples, bf16", "ms_median": med, "ms_best": best,
                "samples_per_s": 32 * 16000 / (med * 1e-3),
                "wn_gemm_tflops": WN_FLOP_PER_STEP * 32 * 2000 / (med * 1e-3) / 1e12})



def stft_configs(model, taco, peaks, out):
    y = syn.synthetic_waveforms(256, 220160, sr=22050, seed=5).to(DEV)
    n = y.numel()
    frames = 256 * 861
    med, best = timeit(lambda: taco.mel_spectrogram(y))
    out.append({"config": "cfg5a: TacotronSTFT.mel_spectrogram 256 x 220160 samples", "ms_median": med, "ms_best": best,
                "samples_per_s": n / (med * 1e-3), "algorithmic_gb_s": n * 5.25 / (med * 1e-3) / 1e9,
                "frac_hbm": n * 5.25 / (med * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "dense_basis_tflops": frames * (STFT_FLOP_PER_FRAME + 2 * 80 * 513) / (med * 1e-3) / 1e12})
    out[-1]["breakdown"] = breakdown(lambda: taco.mel_spectrogram(y))
    den = t2s.Denoiser(model)
    med, best = timeit(lambda: den(y, strength=0.01))
    out.append({"config": "cfg5b: Denoiser(strength 0.01) 256 x 220160 samples", "ms_median": med, "ms_best": best,
                "samples_per_s": n / (med * 1e-3), "algorithmic_gb_s": n * 8 / (med * 1e-3) / 1e9,
                "frac_hbm": n * 8 / (med * 1e-3) / 1e9 / peaks["hbm_gbs"],
                "dense_basis_tflops": frames * 2 * STFT_FLOP_PER_FRAME / (med * 1e-3) / 1e12})

    out[-1]["breakdown"] = breakdown(lambda: den(y, strength=0.01))


if __name__ == "__main__":
    main()
