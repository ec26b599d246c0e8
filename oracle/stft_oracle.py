"""CPU oracle: conv-basis STFT, mel front end and Denoiser (TEST INFRASTRUCTURE).

Restates ``utils/stft.py``, ``utils/layers.py:42-79``, ``utils/audio_processing.py`` and
``waveglow/denoiser.py`` of the reference as plain functions on CPU tensors (no ``.cuda()``
round trips, no module state).  ``mel_filterbank`` restates librosa 0.6.0's
``filters.mel`` (third-party, pinned at waveglow/requirements.txt:5, absent from the
reference tree): PARITY UNPINNED for that one table, see oracle/__init__.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from scipy.signal import get_window

Tensor = torch.Tensor


def _centered(window: np.ndarray, size: int) -> np.ndarray:
    """librosa.util.pad_center: zero-pad ``window`` symmetrically to ``size`` (stft.py:61)."""
    lpad = (size - len(window)) // 2
    return np.pad(window, (lpad, size - len(window) - lpad))


def stft_bases(filter_length: int, hop_length: int, win_length: int, window: str = "hann"):
    """(forward_basis [2*cutoff,1,L], inverse_basis [2*cutoff,1,L]) as float32 tensors.

    stft.py:45-66: rows 0..cutoff-1 are Re(DFT), the next cutoff rows Im(DFT) (so they carry
    -sin), both multiplied by the periodic window; the inverse basis is
    pinv(scale * [Re;Im]).T times the same window, scale = L / hop.
    """
    scale = filter_length / hop_length
    dft = np.fft.fft(np.eye(filter_length))
    cutoff = filter_length // 2 + 1
    stacked = np.vstack([dft[:cutoff].real, dft[:cutoff].imag])
    fwd = torch.from_numpy(stacked[:, None, :]).float()
    inv = torch.from_numpy(np.linalg.pinv(scale * stacked).T[:, None, :]).float()
    if window is not None:
        assert filter_length >= win_length                                    # stft.py:58
        win = torch.from_numpy(_centered(get_window(window, win_length, fftbins=True), filter_length)).float()
        fwd = fwd * win
        inv = inv * win
    return fwd.float(), inv.float()


def stft_transform(y: Tensor, forward_basis: Tensor, hop_length: int):
    """y [B,N] -> (magnitude, phase), each [B, cutoff, N//hop + 1]   (stft.py:71-99)."""
    length = forward_basis.shape[2]
    padded = F.pad(y[:, None, None, :], (length // 2, length // 2, 0, 0), mode="reflect")[:, 0]
    spec = F.conv1d(padded, forward_basis, stride=hop_length)
    cutoff = length // 2 + 1
    re, im = spec[:, :cutoff], spec[:, cutoff:]
    return torch.sqrt(re * re + im * im), torch.atan2(im, re)


def window_sumsquare(window: str, n_frames: int, hop_length: int, win_length: int, n_fft: int) -> np.ndarray:
    """Sum of hop-shifted squared windows, float32 (audio_processing.py:7-48)."""
    n = n_fft + hop_length * (n_frames - 1)
    env = np.zeros(n, dtype=np.float32)
    win_sq = _centered(get_window(window, win_length, fftbins=True) ** 2, n_fft)
    for i in range(n_frames):
        s = i * hop_length
        env[s: min(n, s + n_fft)] += win_sq[: max(0, min(n_fft, n - s))]
    return env


def stft_inverse(magnitude: Tensor, phase: Tensor, inverse_basis: Tensor, hop_length: int,
                 win_length: int, window: str = "hann") -> Tensor:
    """(magnitude, phase) [B,cutoff,F] -> [B,1,hop*(F-1)]   (stft.py:101-130)."""
    length = inverse_basis.shape[2]
    spec = torch.cat([magnitude * torch.cos(phase), magnitude * torch.sin(phase)], dim=1)
    out = F.conv_transpose1d(spec, inverse_basis, stride=hop_length)
    if window is not None:
        env = window_sumsquare(window, magnitude.shape[-1], hop_length, win_length, length)
        nz = torch.from_numpy(np.where(env > np.finfo(env.dtype).tiny)[0])    # stft.py:118
        env_t = torch.from_numpy(env)
        out[:, :, nz] = out[:, :, nz] / env_t[nz]
        out = out * (float(length) / hop_length)                              # stft.py:125
    return out[:, :, length // 2: out.shape[2] - length // 2]                 # stft.py:127-128


# ----------------------------------------------------------------------------- mel


def _hz_to_mel_slaney(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mel = f / f_sp
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mel)


def _mel_to_hz_slaney(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr: float, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """librosa 0.6.0 ``filters.mel(sr, n_fft, n_mels, fmin, fmax)`` (htk=False, norm=1): triangular
    filters on the Slaney mel scale, each scaled by 2 / (f_hi - f_lo).  Called positionally at
    layers.py:50-51.  Returns float64 [n_mels, n_fft//2+1]; the reference casts to float32."""
    fft_f = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_f[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0.0, np.minimum(lower, upper))
    weights *= (2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels]))[:, None]
    return weights


def mel_spectrogram(y: Tensor, forward_basis: Tensor, mel_basis: Tensor, hop_length: int) -> Tensor:
    """TacotronSTFT.mel_spectrogram (layers.py:63-79): log(clamp(mel_basis @ |STFT|, 1e-5))."""
    assert float(y.min()) >= -1 and float(y.max()) <= 1                       # layers.py:72-73
    mag, _ = stft_transform(y, forward_basis, hop_length)
    return torch.log(torch.clamp(torch.matmul(mel_basis, mag), min=1e-5))     # audio_processing.py:76


# ----------------------------------------------------------------------------- denoiser


def denoiser_bias_spec(bias_audio: Tensor, forward_basis: Tensor, hop_length: int) -> Tensor:
    """|STFT(bias_audio)| first frame, [1, cutoff, 1] (denoiser.py:29-33); ``bias_audio`` is
    ``infer(zeros[1,80,88], sigma=0)`` computed by the caller."""
    mag, _ = stft_transform(bias_audio.float(), forward_basis, hop_length)
    return mag[:, :, 0][:, :, None]


def denoise(audio: Tensor, bias_spec: Tensor, strength: float, forward_basis: Tensor,
            inverse_basis: Tensor, hop_length: int, win_length: int) -> Tensor:
    """Denoiser.forward (denoiser.py:35-40): spectral subtraction then ISTFT -> [B,1,N]."""
    mag, phase = stft_transform(audio.float(), forward_basis, hop_length)
    mag = torch.clamp(mag - bias_spec * strength, 0.0)
    return stft_inverse(mag, phase, inverse_basis, hop_length, win_length)


# ----------------------------------------------------------------------------- callers


def griffin_lim_initial_angles(shape, seed: int) -> Tensor:
    """The reference's random initial phase (audio_processing.py:59-61) under ``np.random.seed(seed)``."""
    np.random.seed(seed)
    return torch.from_numpy(np.angle(np.exp(2j * np.pi * np.random.rand(*shape))).astype(np.float32))


def griffin_lim(magnitudes: Tensor, angles: Tensor, forward_basis: Tensor, inverse_basis: Tensor, hop_length: int,
                win_length: int, n_iters: int = 30) -> Tensor:
    """Griffin-Lim (audio_processing.py:51-67): inverse with the given angles, then n_iters rounds of
    transform -> keep the phase -> inverse with the target magnitudes.  Returns [B, hop*(F-1)]."""
    signal = stft_inverse(magnitudes, angles, inverse_basis, hop_length, win_length).squeeze(1)
    for _ in range(n_iters):
        _, angles = stft_transform(signal, forward_basis, hop_length)
        signal = stft_inverse(magnitudes, angles, inverse_basis, hop_length, win_length).squeeze(1)
    return signal


def pcm16(audio: Tensor, max_wav_value: float = 32768.0):
    """The CLI's output conversion (waveglow/inference.py:58-62): (audio * MAX_WAV_VALUE) -> int16 by C
    truncation.  Returns (int16 array, in_range mask); out-of-range products are undefined behaviour in
    the reference (numpy astype), so parity is asserted on the in-range samples only."""
    scaled = (audio.float() * max_wav_value).squeeze().numpy()
    in_range = (scaled > -32768.0) & (scaled < 32768.0)
    return np.trunc(np.clip(scaled, -32768.0, 32767.0)).astype(np.int16), in_range
