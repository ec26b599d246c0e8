"""Host-side signal helpers (reference: utils/audio_processing.py and the librosa calls it makes).

One-off table construction runs on the host in float64; per-call work is on the GPU (see stft.py).
``mel_filterbank`` re-implements librosa 0.6's ``filters.mel`` defaults (Slaney scale, area
normalisation) because the reference takes the table from that third-party package (layers.py:50-51).
"""
from __future__ import annotations

import math

import numpy as np
import torch
from scipy.signal import get_window


def padded_window(window: str, win_length: int, n_fft: int) -> np.ndarray:
    """Periodic window of win_length, zero-padded symmetrically to n_fft (stft.py:60-62)."""
    win = get_window(window, win_length, fftbins=True).astype(np.float64)
    left = (n_fft - win_length) // 2
    out = np.zeros(n_fft, dtype=np.float64)
    out[left: left + win_length] = win
    return out


def window_sumsquare(window, n_frames, hop_length=200, win_length=800, n_fft=800, dtype=np.float32, norm=None):
    """Sum-square window envelope, same signature/semantics as the reference (audio_processing.py:7-48).
    Host-side and only for API parity: the GPU ISTFT rebuilds the envelope per sample in its epilogue."""
    if win_length is None:
        win_length = n_fft
    assert norm is None
    n = n_fft + hop_length * (n_frames - 1)
    env = np.zeros(n, dtype=dtype)
    sq = padded_window(window, win_length, n_fft) ** 2
    for i in range(n_frames):
        lo = i * hop_length
        env[lo: min(n, lo + n_fft)] += sq[: max(0, min(n_fft, n - lo))]
    return env


def griffin_lim(magnitudes, stft_fn, n_iters=30, angles=None):
    """Griffin-Lim phase reconstruction (audio_processing.py:51-67) on the GPU STFT kernels.

    magnitudes [B, cutoff, F] (CUDA); stft_fn: this package's STFT.  ``angles`` optionally supplies the
    initial phase (the reference draws it with numpy: ``angle(exp(2j*pi*rand))``); None draws it the same
    way on the host.  Each iteration is transform -> keep phase, impose magnitudes -> inverse; the
    projection runs as one in-place kernel on the channels-last spectrum (no atan2 / cos / sin).
    """
    from . import _lib
    if not magnitudes.is_cuda:
        raise RuntimeError("griffin_lim needs CUDA tensors on a B200; there is no CPU fallback")
    magnitudes = magnitudes.float().contiguous()
    if angles is None:
        angles = np.angle(np.exp(2j * np.pi * np.random.rand(*magnitudes.size()))).astype(np.float32)
        angles = torch.from_numpy(angles)
    angles = angles.to(magnitudes.device).float().contiguous()
    b, cutoff, _ = magnitudes.shape
    with torch.cuda.device(magnitudes.device):
        signal = stft_fn.inverse(magnitudes, angles).squeeze(1)
        for _ in range(n_iters):
            spec, frames, cp = stft_fn._spectrum(signal.contiguous())
            _lib.call("wgb_spec_set_magnitude", spec, magnitudes, b, frames, cutoff, cp, _lib.stream_ptr())
            signal = stft_fn._synthesize(spec, frames, cp).squeeze(1)
    return signal


def dynamic_range_compression(x, C=1, clip_val=1e-5):
    """log(clamp(x, clip_val) * C)  (audio_processing.py:70-76)."""
    return torch.log(torch.clamp(x, min=clip_val) * C)


def dynamic_range_decompression(x, C=1):
    """exp(x) / C  (audio_processing.py:79-85)."""
    return torch.exp(x) / C


def _slaney_hz_to_mel(hz: torch.Tensor) -> torch.Tensor:
    lin = hz * (3.0 / 200.0)
    log_part = 15.0 + torch.log(torch.clamp(hz, min=1e-30) / 1000.0) * (27.0 / math.log(6.4))
    return torch.where(hz >= 1000.0, log_part, lin)


def _slaney_mel_to_hz(mel: torch.Tensor) -> torch.Tensor:
    lin = mel * (200.0 / 3.0)
    log_part = 1000.0 * torch.exp((mel - 15.0) * (math.log(6.4) / 27.0))
    return torch.where(mel >= 15.0, log_part, lin)


def mel_filterbank(sr: float, n_fft: int, n_mels: int = 80, fmin: float = 0.0, fmax: float = None) -> torch.Tensor:
    """[n_mels, n_fft//2+1] float32 triangular Slaney filterbank with 2/(f_hi - f_lo) normalisation."""
    fmax = sr / 2.0 if fmax is None else fmax
    f64 = torch.float64
    bins = torch.linspace(0.0, sr / 2.0, n_fft // 2 + 1, dtype=f64)
    lo, hi = _slaney_hz_to_mel(torch.tensor(float(fmin), dtype=f64)), _slaney_hz_to_mel(torch.tensor(float(fmax), dtype=f64))
    edges = _slaney_mel_to_hz(torch.linspace(float(lo), float(hi), n_mels + 2, dtype=f64))
    width = edges[1:] - edges[:-1]
    rising = (bins[None, :] - edges[:-2, None]) / width[:-1, None]
    falling = (edges[2:, None] - bins[None, :]) / width[1:, None]
    fb = torch.clamp(torch.minimum(rising, falling), min=0.0)
    fb = fb * (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return fb.float()
