"""Golden fixtures for the caller-side rows: vocoder CLI, checkpoint formats, Griffin-Lim.

    python tests/golden/make_golden_cli.py          (build container only: imports /root/reference read-only)

Writes
  tiny_ckpt_ref.pt      {'model': <reference glow.WaveGlow, weight norm applied>, ...} exactly as
                        waveglow/train.py:52-60 pickles it (tiny architecture: 8 mels, 4 flows, WN 2 x 16)
  tiny_ckpt_ref_old.pt  the same network in the pre-`res_skip_layers` layout (`res_layers` / `skip_layers`)
                        that waveglow/convert_model.py migrates
  cli_golden.npz        reference outputs: infer on both tiny checkpoints (after the reference's own
                        remove_weightnorm / update_model), the CLI's int16 samples (infer -> Denoiser ->
                        * 32768 -> astype int16, waveglow/inference.py:55-62) on the full-size synthetic model,
                        and griffin_lim (utils/audio_processing.py:51-67) with seeded initial angles.
"""
from __future__ import annotations

import copy
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import ref_harness as rh          # noqa: E402
from tests.golden.make_golden import flow_channels  # noqa: E402
from text2speech_b200 import synthetic as syn       # noqa: E402

TINY_CONFIG = {"n_mel_channels": 8, "n_flows": 4, "n_group": 8, "n_early_every": 2, "n_early_size": 2,
               "WN_config": {"n_layers": 2, "n_channels": 16, "kernel_size": 3}}
SIGMA = 0.666
GL_SEED, GL_ITERS = 7, 4


def tiny_inputs():
    g = torch.Generator().manual_seed(11)
    mel = torch.randn((2, TINY_CONFIG["n_mel_channels"], 5), generator=g)
    z = torch.randn((2, 8, 5 * 32), generator=g)
    return mel, z


def detach_cache(model):
    """torch >= 2.1 cannot deepcopy the non-leaf ``weight`` attribute weight_norm leaves on a module
    (the reference's update_model deep-copies, convert_model.py:14): swap in detached tensors first."""
    for m in model.modules():
        for hook in m._forward_pre_hooks.values():
            name = getattr(hook, "name", None)
            if name is not None and isinstance(m.__dict__.get(name), torch.Tensor):
                m.__dict__[name] = m.__dict__[name].detach()
    return model


def split_res_skip(model):
    """Current-layout reference model -> the old layout convert_model.py expects (inverse of update_model)."""
    old = copy.deepcopy(detach_cache(model))
    for wn in old.WN:
        res, skip = torch.nn.ModuleList(), torch.nn.ModuleList()
        n_ch = wn.n_channels
        for i, layer in enumerate(wn.res_skip_layers):
            plain = torch.nn.utils.remove_weight_norm(copy.deepcopy(layer))       # layer cache already detached
            w, b = plain.weight.detach(), plain.bias.detach()
            parts = [(w[:n_ch], b[:n_ch], res), (w[n_ch:], b[n_ch:], skip)] if i < wn.n_layers - 1 else [(w, b, skip)]
            for wp, bp, dst in parts:
                conv = torch.nn.Conv1d(n_ch, wp.shape[0], 1)
                conv.weight = torch.nn.Parameter(wp.clone())
                conv.bias = torch.nn.Parameter(bp.clone())
                dst.append(torch.nn.utils.weight_norm(conv, name="weight"))
        wn.res_layers, wn.skip_layers = res, skip
        del wn.res_skip_layers
    return old


def main():
    warnings.simplefilter("ignore")
    ref_glow, ref_denoiser, ref_stft, ref_layers, feed = rh.load()
    import convert_model as ref_convert                      # reference waveglow/convert_model.py
    from utils.audio_processing import griffin_lim as ref_griffin_lim
    out = {}

    # ---- tiny pickled checkpoints
    torch.manual_seed(99)
    tiny = ref_glow.WaveGlow(**TINY_CONFIG)
    g = torch.Generator().manual_seed(100)
    for wn in tiny.WN:
        wn.end.weight.data.normal_(0, 0.05, generator=g)
        wn.end.bias.data.normal_(0, 0.05, generator=g)
    torch.save({"model": tiny, "iteration": 7, "learning_rate": 1e-4}, os.path.join(HERE, "tiny_ckpt_ref.pt"))
    tiny_old = split_res_skip(tiny)
    torch.save({"model": tiny_old, "iteration": 7, "learning_rate": 1e-4}, os.path.join(HERE, "tiny_ckpt_ref_old.pt"))

    mel, z = tiny_inputs()
    chans = flow_channels(TINY_CONFIG)
    for name, model in (("tiny", copy.deepcopy(detach_cache(tiny))),
                        ("tiny_old", ref_convert.update_model(detach_cache(tiny_old)))):
        model = ref_glow.WaveGlow.remove_weightnorm(model).eval()
        feed.load(z, chans)
        with torch.no_grad():
            out[f"{name}_audio"] = model.infer(mel, sigma=SIGMA).numpy()
    assert np.abs(out["tiny_audio"] - out["tiny_old_audio"]).max() < 1e-5      # the migration preserves the network

    # ---- CLI samples on the full-size synthetic model (inference.py:48-62, one mel at a time)
    cfg = syn.load_config()
    sd = syn.synthetic_state_dict(cfg, seed=1234, end_std=0.01)
    model = rh.build_reference_waveglow(ref_glow, cfg, sd, weight_norm=False)
    with rh.cpu_cuda_noop(), torch.no_grad():
        den = ref_denoiser.Denoiser(model)
        for idx, frames in enumerate((5, 6, 5)):
            m = syn.synthetic_mel(1, frames, seed=40 + idx)
            zz = syn.synthetic_z(1, frames, seed=50 + idx)
            feed.load(zz, flow_channels(cfg))
            audio = model.infer(m, sigma=SIGMA)
            plain = (audio * 32768.0).squeeze().numpy().astype("int16")
            audio_d = den(audio, 0.1) * 32768.0
            out[f"cli_pcm_{idx}"] = plain
            out[f"cli_pcm_denoised_{idx}"] = audio_d.squeeze().numpy().astype("int16")
            out[f"cli_float_denoised_{idx}"] = audio_d.squeeze().numpy()

        # ---- Griffin-Lim
        dc = syn.DEFAULT_DATA_CONFIG
        stft = ref_stft.STFT(dc["filter_length"], dc["hop_length"], dc["win_length"])
        y = syn.synthetic_waveforms(2, 4096, sr=dc["sampling_rate"], seed=5)
        mag, _ = stft.transform(y)
        np.random.seed(GL_SEED)
        out["gl_signal"] = ref_griffin_lim(mag, stft, n_iters=GL_ITERS).numpy()
    np.savez_compressed(os.path.join(HERE, "cli_golden.npz"), **out)
    for f in ("tiny_ckpt_ref.pt", "tiny_ckpt_ref_old.pt", "cli_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
    for k, v in out.items():
        print(f"  {k:28s} {tuple(v.shape)} {v.dtype} absmax={np.abs(v).max():.4g}")


if __name__ == "__main__":
    main()
