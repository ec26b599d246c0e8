"""Golden gradients of the UNMODIFIED reference (build container only): tests/golden/train_grads_golden.npz.

    python tests/golden/make_golden_grads.py

Runs the reference ``glow.WaveGlow`` (weight-norm layout, train mode) + ``glow.WaveGlowLoss`` (waveglow/train.py:72,
116-120) on a reduced flow count (4 flows, early outputs every 2 -> n_half 4,4,3,3; WN unchanged: 8 layers x 512 ch)
and stores, per parameter, the gradient's L2 norm and 512 sampled entries (fixed seed; whole tensor when smaller),
plus the loss.  Weights / inputs are regenerated from seeds by ``text2speech_b200.synthetic``.
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

N_SAMPLES = 512
SIGMA = 1.0


def train_config():
    cfg = dict(syn.load_config())
    cfg.update(n_flows=4, n_early_every=2, n_early_size=2)
    return cfg


def train_inputs(bsz=2, frames=8):
    mel = syn.synthetic_mel(bsz, frames, seed=11)
    g = torch.Generator().manual_seed(12)
    wav = (0.1 * torch.randn((bsz, frames * 256), generator=g)).clamp(-1, 1)
    return mel, wav


def sample_index(name: str, numel: int) -> np.ndarray:
    if numel <= N_SAMPLES:
        return np.arange(numel)
    g = torch.Generator().manual_seed(abs(hash_name(name)) % (2 ** 31))
    return torch.randperm(numel, generator=g)[:N_SAMPLES].numpy()


def hash_name(name: str) -> int:
    h = 0
    for ch in name:
        h = (h * 131 + ord(ch)) % 1000003
    return h


def main():
    ref_glow = rh.load()[0]
    cfg = train_config()
    sd = syn.synthetic_state_dict(cfg, seed=777, end_std=0.05, weight_norm=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = rh.build_reference_waveglow(ref_glow, cfg, sd, weight_norm=True).train()
        mel, wav = train_inputs()
        outputs = model((mel, wav))
        loss = ref_glow.WaveGlowLoss(SIGMA)(outputs)
        loss.backward()
    out = {"loss": np.float32(loss.item())}
    for name, p in model.named_parameters():
        g = p.grad.detach().float().flatten()
        out[name + "|norm"] = np.float32(g.norm().item())
        out[name + "|samples"] = g[torch.from_numpy(sample_index(name, g.numel()))].numpy().astype(np.float32)
    path = os.path.join(HERE, "train_grads_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays; loss", float(loss))


if __name__ == "__main__":
    main()
