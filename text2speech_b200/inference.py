"""Vocoder CLI (reference: waveglow/inference.py): mel .pt files -> WaveGlow.infer -> Denoiser -> int16 wav.

    python -m text2speech_b200.inference -f mel_files.txt -w checkpoint.pt -o out_dir -s 0.6 [-d 0.1] [--batch 16]

Same ``main(mel_files, waveglow_path, sigma, output_dir, sampling_rate, is_fp16, denoiser_strength)``
signature and the same per-file outputs (``<name>_synthesis.wav``, int16 = trunc(audio * 32768)) as
the reference (inference.py:34-64).  Differences, all additive:

  * checkpoints: the reference unpickles a whole ``glow.WaveGlow`` module (inference.py:37); here
    ``load_waveglow`` accepts that pickle (the class path ``glow.*`` is aliased to this package's
    drop-in classes, so no reference code is needed), old res/skip-split pickles
    (convert_model.py:11-38), ``{'model': state_dict}`` / ``{'state_dict': ...}`` / bare state_dicts;
  * ``batch`` > 1 groups mels with the same number of frames into one ``infer`` call (utterances are
    independent, so the result is bit-identical to the one-at-a-time loop);
  * ``* MAX_WAV_VALUE`` + int16 conversion run in one kernel on the GPU (half the D2H bytes);
    values outside int16 saturate (numpy's ``astype('int16')`` on out-of-range floats is undefined).
"""
from __future__ import annotations

import os
import sys
from collections import OrderedDict
from typing import Dict, List, Optional

import torch

from . import _lib
from . import glow as _glow
from .convert_model import update_model
from .denoiser import Denoiser
from .mel2samp import MAX_WAV_VALUE, files_to_list
from .synthetic import load_config


def install_glow_alias() -> None:
    """Make pickles that name ``glow.WaveGlow`` / ``glow.WN`` / ``glow.Invertible1x1Conv`` (every
    published WaveGlow checkpoint; inference.py:37) resolve to this package's drop-in classes."""
    mod = sys.modules.get("glow")
    if mod is None or not hasattr(mod, "WaveGlow"):
        sys.modules["glow"] = _glow


def _model_from_state_dict(state: Dict[str, torch.Tensor], config: Optional[Dict]) -> _glow.WaveGlow:
    cfg = config if config is not None else load_config()
    model = _glow.WaveGlow(**cfg)
    if not any(k.endswith(".weight_g") for k in state):
        model = _glow.WaveGlow.remove_weightnorm(model)
    model.load_state_dict(state, strict=True)
    return model


def load_waveglow(path: str, config: Optional[Dict] = None) -> _glow.WaveGlow:
    """Load any of the checkpoint layouts listed in the module docstring; returns an eval-mode
    drop-in WaveGlow on the CPU (move it with ``.cuda()`` like the reference does)."""
    install_glow_alias()
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    obj = ckpt
    if isinstance(ckpt, dict) and not isinstance(ckpt, OrderedDict):
        obj = ckpt.get("model", ckpt.get("state_dict", ckpt))
    if isinstance(obj, torch.nn.Module):
        if not isinstance(obj, _glow.WaveGlow):
            raise TypeError(f"checkpoint holds a {type(obj).__module__}.{type(obj).__name__}, expected glow.WaveGlow")
        model = update_model(obj)
    elif isinstance(obj, dict):
        model = _model_from_state_dict(obj, config)
    else:
        raise TypeError(f"unrecognised checkpoint content: {type(obj)}")
    return model.eval()


def audio_to_int16(audio: torch.Tensor, scale: float = MAX_WAV_VALUE) -> torch.Tensor:
    """trunc(audio * scale) as int16 on the GPU (inference.py:58-62), saturating."""
    flat = audio.float().contiguous()
    out = torch.empty(flat.shape, device=flat.device, dtype=torch.int16)
    with torch.cuda.device(flat.device):
        _lib.call("wgb_audio_to_int16", flat, out, flat.numel(), float(scale), _lib.stream_ptr())
    return out


def synthesize(waveglow: _glow.WaveGlow, mels: List[torch.Tensor], sigma: float, denoiser: Optional[Denoiser] = None,
               denoiser_strength: float = 0.0, is_fp16: bool = False, batch: int = 1,
               z: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
    """mels: list of [80, F_i] tensors -> list of int16 CPU tensors [256 * F_i] (same order).
    ``z`` optionally supplies the noise per utterance ([8, 32 F_i]) for reproducible runs."""
    device = waveglow.upsample.weight.device
    order = sorted(range(len(mels)), key=lambda i: mels[i].shape[-1])
    out: List[Optional[torch.Tensor]] = [None] * len(mels)
    pos = 0
    while pos < len(order):
        group = [order[pos]]
        while (len(group) < max(1, batch) and pos + len(group) < len(order)
               and mels[order[pos + len(group)]].shape[-1] == mels[group[0]].shape[-1]):
            group.append(order[pos + len(group)])
        pos += len(group)
        mel = torch.stack([mels[i] for i in group]).to(device, non_blocking=True)
        mel = mel.half() if is_fp16 else mel
        zz = None if z is None else torch.stack([z[i] for i in group]).to(device, non_blocking=True)
        with torch.no_grad():
            audio = waveglow.infer(mel, sigma=sigma, z=zz)
            if denoiser is not None and denoiser_strength > 0:
                audio = denoiser(audio, denoiser_strength)
            pcm = audio_to_int16(audio.reshape(len(group), -1)).cpu()
        for row, i in enumerate(group):
            out[i] = pcm[row]
    return out  # type: ignore[return-value]


def main(mel_files, waveglow_path, sigma, output_dir, sampling_rate, is_fp16, denoiser_strength, batch=1,
         config_path=None):
    from scipy.io.wavfile import write
    mel_files = files_to_list(mel_files)
    config = load_config(config_path) if config_path else None
    waveglow = load_waveglow(waveglow_path, config)
    waveglow = waveglow.remove_weightnorm(waveglow) if _has_weight_norm(waveglow) else waveglow
    waveglow.cuda().eval()
    denoiser = Denoiser(waveglow).cuda() if denoiser_strength > 0 else None
    mels = [torch.load(p, map_location="cpu", weights_only=False) for p in mel_files]
    pcm = synthesize(waveglow, mels, sigma, denoiser, denoiser_strength, is_fp16, batch)
    os.makedirs(output_dir, exist_ok=True)
    paths = []
    for file_path, samples in zip(mel_files, pcm):
        file_name = os.path.splitext(os.path.basename(file_path))[0]
        audio_path = os.path.join(output_dir, "{}_synthesis.wav".format(file_name))
        write(audio_path, sampling_rate, samples.numpy())
        print(audio_path)
        paths.append(audio_path)
    return paths


def _has_weight_norm(model: _glow.WaveGlow) -> bool:
    return any(k.endswith(".weight_g") for k in model.state_dict())


if __name__ == "__main__":
    import argparse

    parser = argparse.ArgumentParser()
    parser.add_argument("-f", "--filelist_path", required=True)
    parser.add_argument("-w", "--waveglow_path", help="Path to waveglow decoder checkpoint with model")
    parser.add_argument("-o", "--output_dir", required=True)
    parser.add_argument("-s", "--sigma", default=1.0, type=float)
    parser.add_argument("--sampling_rate", default=22050, type=int)
    parser.add_argument("--is_fp16", action="store_true")
    parser.add_argument("-d", "--denoiser_strength", default=0.0, type=float,
                        help="Removes model bias. Start with 0.1 and adjust")
    parser.add_argument("--batch", default=1, type=int, help="utterances of equal length per infer call")
    parser.add_argument("-c", "--config", default=None, help="reference-format config.json (for state_dict checkpoints)")
    args = parser.parse_args()
    main(args.filelist_path, args.waveglow_path, args.sigma, args.output_dir, args.sampling_rate, args.is_fp16,
         args.denoiser_strength, args.batch, args.config)
