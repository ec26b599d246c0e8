"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

Inputs and weights are NOT stored: they are regenerated from seeds by
``text2speech_b200.synthetic`` (same torch build on the GPU box).  Only reference OUTPUTS are
stored, as small float32 arrays.  The reference cannot travel to the GPU box (it is a read-only
mount of the build container), which is why these fixtures are committed.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import ref_harness as rh          # noqa: E402
from text2speech_b200 import synthetic as syn       # noqa: E402

RECIPES = {                       # name -> synthetic_state_dict kwargs
    "bench": dict(seed=1234, end_std=0.01, gain=1.0),
    "stress": dict(seed=4321, end_std=0.01, gain=2.0),
    # non-orthogonal invertible 1x1 convs (W^-1 != W^T, log det W != 0, two flows with det W < 0) and a stronger coupling
    "skew": dict(seed=777, end_std=0.03, gain=1.0, mix="skew"),
}
SIGMA = 0.666
TAP_FLOW = 11
TAP_LAYERS = (0, 3, 7)
TAP_STEPS = 24


def flow_channels(cfg):
    out, c = [], cfg["n_group"]
    for k in range(cfg["n_flows"]):
        if k % cfg["n_early_every"] == 0 and k > 0:
            c -= cfg["n_early_size"]
        out.append(c)
    return out


def golden_waveglow(ref_glow, feed, name, kw, out):
    cfg = syn.load_config()
    chans = flow_channels(cfg)
    sd = syn.synthetic_state_dict(cfg, **kw)
    model = rh.build_reference_waveglow(ref_glow, cfg, sd, weight_norm=False)
    bsz, frames = 2, 6
    mel = syn.synthetic_mel(bsz, frames, seed=0)
    z = syn.synthetic_z(bsz, frames, seed=2024)

    # ---- per-layer taps of WN[TAP_FLOW] via forward hooks on the reference modules
    taps = {}
    wn = model.WN[TAP_FLOW]
    hooks = []
    def tap_in(i):
        def hook(mod, inp, outp):
            taps[f"h{i}"] = inp[0].detach().clone()
        return hook

    def tap_rs(i):
        def hook(mod, inp, outp):
            taps[f"acts{i}"] = inp[0].detach().clone()
            taps[f"rs{i}"] = outp.detach().clone()
        return hook

    def tap_wn(k):
        def hook(mod, inp, outp):
            wn_out[k] = outp.detach().clone()
        return hook

    for i in TAP_LAYERS:
        hooks.append(wn.in_layers[i].register_forward_hook(tap_in(i)))
        hooks.append(wn.res_skip_layers[i].register_forward_hook(tap_rs(i)))
    wn_out = {}
    for k in (11, 5, 0):
        hooks.append(model.WN[k].register_forward_hook(tap_wn(k)))
    feed.load(z, chans)
    with torch.no_grad():
        audio = model.infer(mel, sigma=SIGMA)
    for h in hooks:
        h.remove()
    assert torch.isfinite(audio).all()
    steps = np.linspace(0, mel.shape[2] * 32 - 1, TAP_STEPS).round().astype(np.int64)
    out[f"{name}_infer_audio"] = audio.numpy()
    out[f"{name}_tap_steps"] = steps
    for key, val in taps.items():
        out[f"{name}_wn{TAP_FLOW}_{key}"] = val[0][:, steps].numpy()       # batch 0, [C, steps]
    for k, val in wn_out.items():
        out[f"{name}_wn{k}_out"] = val.numpy()

    # ---- forward direction (glow.py:207-249)
    g = torch.Generator().manual_seed(1)
    wav = (0.1 * torch.randn((bsz, frames * 256), generator=g)).clamp(-1, 1)
    with torch.no_grad():
        zf, log_s, log_det = model((mel, wav))
    out[f"{name}_fwd_z"] = zf.numpy()
    for k in (0, 5, 11):
        out[f"{name}_fwd_log_s{k}"] = log_s[k].numpy()
    out[f"{name}_fwd_log_det"] = np.array([float(v) for v in log_det], dtype=np.float64)
    # forward with audio shorter than the upsampled mel (trim branch, glow.py:216-218)
    with torch.no_grad():
        zs, _, _ = model((mel, wav[:, : frames * 256 - 64]))
    out[f"{name}_fwd_z_short"] = zs.numpy()

    if name == "bench":
        # weight-norm layout: same network but g != ||v||  (pins the fold, glow.py:294-310)
        sd_wn = syn.synthetic_state_dict(cfg, weight_norm=True, **kw)
        model_wn = rh.build_reference_waveglow(ref_glow, cfg, sd_wn, weight_norm=True)
        feed.load(z[:1, :, : 4 * 32], chans)
        with torch.no_grad():
            out["bench_wnorm_infer_audio"] = model_wn.infer(mel[:1, :, :4], sigma=SIGMA).numpy()
    return model


def golden_stft(ref_stft, ref_layers, ref_denoiser, model, out):
    dc = syn.DEFAULT_DATA_CONFIG
    fl, hop, win = dc["filter_length"], dc["hop_length"], dc["win_length"]
    y = syn.synthetic_waveforms(2, 4096, sr=dc["sampling_rate"], seed=5)
    with rh.cpu_cuda_noop(), torch.no_grad():
        stft = ref_stft.STFT(fl, hop, win)
        mag, phase = stft.transform(y)
        recon = stft.inverse(mag, phase)
        taco = ref_layers.TacotronSTFT(fl, hop, win, 80, dc["sampling_rate"], dc["mel_fmin"], dc["mel_fmax"])
        mel = taco.mel_spectrogram(y)
        den = ref_denoiser.Denoiser(model, filter_length=fl, n_overlap=4, win_length=win, mode="zeros")
        den_out = den(y, strength=0.1)
        den_out2 = den(y, strength=0.01)
        small = ref_stft.STFT(64, 16, 48)         # filter_length > win_length branch (stft.py:58-62)
        mag_s, ph_s = small.transform(y[:, :512])
        rec_s = small.inverse(mag_s, ph_s)
        nowin = ref_stft.STFT(64, 16, 64, window=None)   # window=None: no envelope division, no L/hop scale (stft.py:111)
        mag_n, ph_n = nowin.transform(y[:, :512])
        rec_n = nowin.inverse(mag_n, ph_n)
    rows = np.array([0, 1, 2, 100, 256, 511, 512, 513, 514, 700, 1024, 1025])
    out["stft_rows"] = rows
    out["stft_forward_basis_rows"] = stft.forward_basis[rows, 0].numpy()
    out["stft_inverse_basis_rows"] = stft.inverse_basis[rows, 0].numpy()
    out["stft_mag"] = mag.numpy()
    out["stft_phase"] = phase.numpy()
    out["stft_recon"] = recon.numpy()
    out["mel_basis_22050"] = taco.mel_basis.numpy()          # via torchaudio-backed shim (unpinned)
    out["mel"] = mel.numpy()
    out["denoiser_bias_spec"] = den.bias_spec.numpy()
    out["denoised_s0p1"] = den_out.numpy()
    out["denoised_s0p01"] = den_out2.numpy()
    out["small_mag"] = mag_s.numpy()
    out["small_recon"] = rec_s.numpy()
    out["nowin_mag"] = mag_n.numpy()
    out["nowin_recon"] = rec_n.numpy()
    from utils.audio_processing import window_sumsquare     # reference helper
    out["wss_17"] = window_sumsquare("hann", 17, hop_length=hop, win_length=win, n_fft=fl, dtype=np.float32)


def main():
    warnings.simplefilter("ignore")
    torch.set_num_threads(os.cpu_count() or 1)
    ref_glow, ref_denoiser, ref_stft, ref_layers, feed = rh.load()
    wg, st = {}, {}
    model = None
    for name, kw in RECIPES.items():
        m = golden_waveglow(ref_glow, feed, name, kw, wg)
        if name == "bench":
            model = m
    golden_stft(ref_stft, ref_layers, ref_denoiser, model, st)
    np.savez_compressed(os.path.join(HERE, "waveglow_golden.npz"), **wg)
    np.savez_compressed(os.path.join(HERE, "stft_golden.npz"), **st)
    for f in ("waveglow_golden.npz", "stft_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
    for k, v in {**wg, **st}.items():
        print(f"  {k:40s} {tuple(v.shape)} absmax={np.abs(v).max():.4g}")


if __name__ == "__main__":
    main()
