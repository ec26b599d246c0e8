"""Deterministic synthetic weights / inputs for the vocoding path.

There is no network for checkpoints, so tests, the golden-vector generator and bench.py all
draw the same random-init ``state_dict`` (reference key layout) from this recipe.
Every tensor has its own generator seeded from (seed, key), so the values do not depend on
construction order, on the reference being importable, or on which keys are requested.

Statistics mimic the reference's random init (PyTorch Conv1d default: U(+-1/sqrt(fan_in)) for
weight and bias, weight_norm with g=||v||; orthogonal convinv from a QR, glow.py:73-80) except
``WN.k.end`` which the reference zero-initialises (glow.py:128-131): zeros would make every
affine coupling the identity and hide WN bugs, so it is drawn N(0, end_std^2) instead.
"""
from __future__ import annotations

import json
import math
import zlib
from collections import OrderedDict
from typing import Dict, Optional

import torch

DEFAULT_WAVEGLOW_CONFIG = {          # == waveglow/config.json:27-38 of the reference
    "n_mel_channels": 80,
    "n_flows": 12,
    "n_group": 8,
    "n_early_every": 4,
    "n_early_size": 2,
    "WN_config": {"n_layers": 8, "n_channels": 512, "kernel_size": 3},
}
DEFAULT_DATA_CONFIG = {              # == waveglow/config.json:12-21
    "segment_length": 16000, "sampling_rate": 22050, "filter_length": 1024,
    "hop_length": 256, "win_length": 1024, "mel_fmin": 0.0, "mel_fmax": 8000.0,
}


def load_config(path: Optional[str] = None) -> Dict:
    """Read a reference-format config.json; returns its ``waveglow_config`` (defaults if None)."""
    if path is None:
        return json.loads(json.dumps(DEFAULT_WAVEGLOW_CONFIG))
    with open(path) as f:
        return json.load(f)["waveglow_config"]


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63))
    return g


def _uniform(shape, bound, seed, key):
    return (torch.rand(shape, generator=_gen(seed, key), dtype=torch.float32) * 2 - 1) * bound


def synthetic_state_dict(config: Optional[Dict] = None, seed: int = 1234, end_std: float = 0.01,
                         gain: float = 1.0, weight_norm: bool = False,
                         mix: str = "orthogonal") -> "OrderedDict[str, torch.Tensor]":
    """Random-init WaveGlow ``state_dict`` in the reference layout.

    weight_norm=False gives the layout after ``WaveGlow.remove_weightnorm`` (``*.weight``);
    True gives ``weight_g`` / ``weight_v`` pairs with g != ||v|| so that folding is exercised.
    mix='orthogonal' is the reference's init of the invertible 1x1 convs (glow.py:73-80: W^-1 = W^T, log|det W| = 0);
    mix='skew' scales the columns of that Q by U[0.5, 2] (a trained-model-like W: W^-1 != W^T, log|det W| != 0) and
    flips the sign of one column in every flow k with k % 5 == 3, so those flows have det W < 0 (torch.logdet, which
    glow.py:100 calls, returns NaN there; W^-1 is still well defined).
    """
    cfg = config or DEFAULT_WAVEGLOW_CONFIG
    n_mel, n_group = cfg["n_mel_channels"], cfg["n_group"]
    wn = cfg["WN_config"]
    n_layers, n_ch, ks = wn["n_layers"], wn["n_channels"], wn["kernel_size"]
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def conv(prefix, c_out, c_in, taps, wnorm):
        bound = 1.0 / math.sqrt(c_in * taps)
        w = _uniform((c_out, c_in, taps), gain * bound, seed, prefix + ".weight")
        b = _uniform((c_out,), bound, seed, prefix + ".bias")
        sd[prefix + ".bias"] = b
        if wnorm:
            scale = 0.5 + torch.rand((c_out, 1, 1), generator=_gen(seed, prefix + ".g"))
            sd[prefix + ".weight_g"] = w.reshape(c_out, -1).norm(dim=1).reshape(c_out, 1, 1) * scale
            sd[prefix + ".weight_v"] = w
        else:
            sd[prefix + ".weight"] = w

    up_k = 1024
    bound = 1.0 / math.sqrt(n_mel * up_k)
    sd["upsample.weight"] = _uniform((n_mel, n_mel, up_k), bound, seed, "upsample.weight")
    sd["upsample.bias"] = _uniform((n_mel,), bound, seed, "upsample.bias")

    n_half, n_rem = n_group // 2, n_group
    chans = []
    for k in range(cfg["n_flows"]):
        if k % cfg["n_early_every"] == 0 and k > 0:
            n_half -= cfg["n_early_size"] // 2
            n_rem -= cfg["n_early_size"]
        chans.append(n_rem)
        p = f"WN.{k}."
        for i in range(n_layers):
            conv(p + f"in_layers.{i}", 2 * n_ch, n_ch, ks, weight_norm)
        for i in range(n_layers):
            rs = 2 * n_ch if i < n_layers - 1 else n_ch
            conv(p + f"res_skip_layers.{i}", rs, n_ch, 1, weight_norm)
        for i in range(n_layers):
            conv(p + f"cond_layers.{i}", 2 * n_ch, n_mel * n_group, 1, weight_norm)
        conv(p + "start", n_ch, n_half, 1, weight_norm)
        sd[p + "end.weight"] = torch.randn((2 * n_half, n_ch, 1), generator=_gen(seed, p + "end.weight")) * end_std
        sd[p + "end.bias"] = torch.randn((2 * n_half,), generator=_gen(seed, p + "end.bias")) * end_std
    for k, c in enumerate(chans):
        q = torch.linalg.qr(torch.randn((c, c), generator=_gen(seed, f"convinv.{k}"), dtype=torch.float64))[0]
        if torch.det(q) < 0:
            q[:, 0] = -q[:, 0]
        if mix == "skew":
            scale = 0.5 + 1.5 * torch.rand((c,), generator=_gen(seed, f"convinv.{k}.scale"), dtype=torch.float64)
            q = q * scale[None, :]
            if k % 5 == 3:
                q[:, 1] = -q[:, 1]
        elif mix != "orthogonal":
            raise ValueError(f"mix must be 'orthogonal' or 'skew', got {mix!r}")
        sd[f"convinv.{k}.conv.weight"] = q.float().reshape(c, c, 1)
    return sd


def synthetic_mel(batch: int, frames: int, n_mel: int = 80, seed: int = 0) -> torch.Tensor:
    """Log-mel-like input in [log 1e-5, 2]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.clamp(-4.0 + 1.5 * torch.randn((batch, n_mel, frames), generator=g), -11.5, 2.0)


def synthetic_z(batch: int, frames: int, n_group: int = 8, seed: int = 2024) -> torch.Tensor:
    """Host-supplied noise [B, n_group, 32*frames] in the layout ``WaveGlow.forward`` returns."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn((batch, n_group, frames * 256 // n_group), generator=g)


def synthetic_waveforms(batch: int, samples: int, sr: int = 22050, seed: int = 5) -> torch.Tensor:
    """Tone + noise waveforms in [-1, 1] for the STFT / mel / denoiser workload."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    f0 = 80.0 + 320.0 * torch.rand((batch, 1), generator=g, dtype=torch.float64)
    n = torch.arange(samples, dtype=torch.float64)[None, :]
    tone = 0.5 * torch.sin(2 * math.pi * f0 * n / sr)
    noise = 0.05 * torch.randn((batch, samples), generator=g, dtype=torch.float64)
    return torch.clamp(tone + noise, -1.0, 1.0).float()


DEFAULT_POSTNET_HPARAMS = {          # == hparams.py:18,146-148 of the reference
    "n_mel_channels": 80, "postnet_embedding_dim": 512, "postnet_kernel_size": 5, "postnet_n_convolutions": 5,
}


DEFAULT_ENCODER_HPARAMS = {          # == hparams.py:111-113 of the reference
    "enc_conv_num_layers": 3, "enc_conv_kernel_size": 5, "enc_conv_channels": 512,
}


def synthetic_encoder_convs_state_dict(hparams: Optional[Dict] = None, seed: int = 78) -> "OrderedDict[str, torch.Tensor]":
    """Random state of the Encoder's conv bank (tacotron/tacotron.py:175-186) in the reference key layout."""
    hp = hparams or DEFAULT_ENCODER_HPARAMS
    n, c = hp["enc_conv_num_layers"], hp["enc_conv_channels"]
    return _conv_bn_state([c] * (n + 1), hp["enc_conv_kernel_size"], [math.sqrt(2.0)] * n, seed)


def synthetic_postnet_state_dict(hparams: Optional[Dict] = None, seed: int = 77) -> "OrderedDict[str, torch.Tensor]":
    """Random Postnet ``state_dict`` in the reference layout (tacotron/modules.py:94-130): xavier-uniform conv weights
    (gain 5/3 before a tanh, 1 for the last layer), and NON-trivial BatchNorm affine parameters / running statistics so
    that folding the BatchNorm is exercised."""
    hp = hparams or DEFAULT_POSTNET_HPARAMS
    n_mel, dim, k, n = hp["n_mel_channels"], hp["postnet_embedding_dim"], hp["postnet_kernel_size"], hp["postnet_n_convolutions"]
    return _conv_bn_state([n_mel] + [dim] * (n - 1) + [n_mel], k, [5.0 / 3.0] * (n - 1) + [1.0], seed)


def _conv_bn_state(chans, k, gains, seed) -> "OrderedDict[str, torch.Tensor]":
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for i in range(len(chans) - 1):
        c_in, c_out = chans[i], chans[i + 1]
        bound = gains[i] * math.sqrt(6.0 / (c_in * k + c_out * k))
        p = f"convolutions.{i}."
        sd[p + "0.conv.weight"] = _uniform((c_out, c_in, k), bound, seed, p + "w")
        sd[p + "0.conv.bias"] = _uniform((c_out,), 1.0 / math.sqrt(c_in * k), seed, p + "b")
        sd[p + "1.weight"] = 0.5 + torch.rand((c_out,), generator=_gen(seed, p + "g"))
        sd[p + "1.bias"] = 0.1 * torch.randn((c_out,), generator=_gen(seed, p + "beta"))
        sd[p + "1.running_mean"] = 0.1 * torch.randn((c_out,), generator=_gen(seed, p + "mean"))
        sd[p + "1.running_var"] = 0.5 + torch.rand((c_out,), generator=_gen(seed, p + "var"))
        sd[p + "1.num_batches_tracked"] = torch.tensor(1, dtype=torch.long)
    return sd
