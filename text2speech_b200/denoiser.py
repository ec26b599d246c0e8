"""Drop-in Denoiser (reference: waveglow/denoiser.py): removes the model's bias spectrum.

``bias_spec`` comes from this package's own ``WaveGlow.infer`` on a zero (or random) mel with
sigma = 0 (denoiser.py:16-33); ``forward`` is STFT -> spectral subtraction -> ISTFT, all on the GPU
(no atan2/sin/cos: the phase enters only through Re/|X| and Im/|X|).
"""
from __future__ import annotations

import torch

from .stft import STFT


class Denoiser(torch.nn.Module):
    """Removes model bias from audio produced with waveglow."""

    def __init__(self, waveglow, filter_length=1024, n_overlap=4, win_length=1024, mode="zeros"):
        super().__init__()
        device = waveglow.upsample.weight.device
        dtype = waveglow.upsample.weight.dtype
        if not device.type == "cuda":
            raise RuntimeError("Denoiser needs the WaveGlow on a CUDA device (B200); there is no CPU fallback")
        self.stft = STFT(filter_length=filter_length, hop_length=int(filter_length / n_overlap),
                         win_length=win_length).to(device)
        if mode == "zeros":
            mel_input = torch.zeros((1, 80, 88), dtype=dtype, device=device)
        elif mode == "normal":
            mel_input = torch.randn((1, 80, 88), dtype=dtype, device=device)
        else:
            raise Exception("Mode {} if not supported".format(mode))
        with torch.no_grad():
            bias_audio = waveglow.infer(mel_input, sigma=0.0).float()
            bias_spec, _ = self.stft.transform(bias_audio)
        self.register_buffer("bias_spec", bias_spec[:, :, 0][:, :, None].contiguous())

    def forward(self, audio: torch.Tensor, strength: float = 0.1) -> torch.Tensor:
        """audio [B, N] -> denoised [B, 1, N'] with N' = hop * (N // hop)  (denoiser.py:35-40)."""
        audio = audio.to(self.bias_spec.device).float().contiguous()
        bias = self.bias_spec.reshape(-1).contiguous()
        with torch.cuda.device(audio.device):
            out = self.stft._denoised_fft(audio, bias, strength)      # stock bases: one butterfly kernel
            if out is not None:
                return out
            if self.stft._use_tc():           # subtraction inside the STFT GEMM's epilogue
                return self.stft._denoised(audio, bias, strength)
            spec, frames, cp = self.stft._spectrum(audio)
            return self.stft._synthesize(spec, frames, cp, denoise=(bias, strength))
