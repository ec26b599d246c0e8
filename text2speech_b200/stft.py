"""Drop-in conv-basis STFT (reference: utils/stft.py) running on sm_100a kernels.

Same constructor, buffers (``forward_basis``, ``inverse_basis`` [2*cutoff, 1, L]) and methods
(``transform`` / ``inverse`` / ``forward``) as the reference.  Both dense-basis contractions are
GEMMs through the C ABI (A = overlapping frames of the reflect-padded signal, row stride = hop);
padding, magnitude/phase, recombination, overlap-add and window-sum normalisation are fused,
coalesced CUDA kernels.  No CPU path.

Conscious deviation: the reference's ``transform`` returns CPU tensors because of its own
``.cuda()...cpu()`` round trip (stft.py:85-89); here results stay on the input's CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .audio_processing import padded_window


def _round4(n: int) -> int:
    return (n + 3) // 4 * 4


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def _split3(w: torch.Tensor) -> torch.Tensor:
    """fp32 [N, K] -> bf16 [N, 3K] = [W_hi | W_hi | W_lo], the weight side of wgb_tc_gemm_split3."""
    hi = w.bfloat16()
    lo = (w - hi.float()).bfloat16()
    return torch.cat([hi, hi, lo], dim=1).contiguous()


class STFT(torch.nn.Module):
    def __init__(self, filter_length=800, hop_length=200, win_length=800, window="hann"):
        super().__init__()
        self.filter_length = filter_length
        self.hop_length = hop_length
        self.win_length = win_length
        self.window = window
        self.forward_transform = None
        length, cutoff = filter_length, filter_length // 2 + 1
        scale = filter_length / hop_length
        # rows 0..cutoff-1: cos(2 pi k n / L); rows cutoff..: -sin(2 pi k n / L)   (stft.py:46-51)
        kn = np.outer(np.arange(cutoff), np.arange(length)) * (2.0 * np.pi / length)
        basis = np.concatenate([np.cos(kn), -np.sin(kn)], axis=0)
        inverse = np.linalg.pinv(scale * basis).T                                  # stft.py:54-55
        if window is not None:
            assert filter_length >= win_length
            win = padded_window(window, win_length, filter_length).astype(np.float32)
        else:
            win = np.ones(length, dtype=np.float32)
        fwd = torch.from_numpy(basis.astype(np.float32)) * torch.from_numpy(win)
        inv = torch.from_numpy(inverse.astype(np.float32)) * torch.from_numpy(win)
        self.register_buffer("forward_basis", fwd[:, None, :].contiguous().float())
        self.register_buffer("inverse_basis", inv[:, None, :].contiguous().float())
        self._pack = None
        # 'tc': split-bf16 tcgen05 GEMMs (fp32-grade accuracy, needs filter_length % 256 == 0 and hop % 8 == 0);
        # 'fp32': CUDA-core FP32 GEMM (any shape).  'auto' picks 'tc' when the shape allows.
        self.precision = "auto"
        # fused mel / denoiser paths: CTA-pair kernels with the unpadded spectrum layout (csrc/stft_tc2.cu) when True,
        # the one-CTA kernels with the 640-bin padded layout (csrc/wn_tc.cu) when False (kept for A/B)
        self.pair = True
        # Denoiser: inverse GEMM with the overlap-add inside (wgb_tc2_istft_ola) instead of GEMM + overlap-add kernel
        self.fused_ola = True
        # 'auto': the fused mel / denoiser paths run as butterflies (csrc/fft.cu: 1024-point real FFTs, one warp per frame)
        # whenever the basis buffers are still the constructor's real-DFT pair, filter_length is 1024 and precision is
        # 'auto', else as the dense-basis contractions; 'gemm' forces the contractions (A/B, validation); 'fft' insists
        self.algorithm = "auto"
        self._fft = None

    # ------------------------------------------------------------------ packed constants
    @property
    def cutoff(self) -> int:
        return self.filter_length // 2 + 1

    def _use_tc(self) -> bool:
        ok = self.filter_length % 256 == 0 and self.hop_length % 8 == 0
        if self.precision == "tc" and not ok:
            raise RuntimeError("tensor-core STFT needs filter_length % 256 == 0 and hop_length % 8 == 0")
        return ok and self.precision in ("auto", "tc")

    def _packed(self, device):
        tc = self._use_tc()
        key = (str(device), tc, self.forward_basis.data_ptr(), self.forward_basis._version,
               self.inverse_basis.data_ptr(), self.inverse_basis._version)
        if self._pack is None or self._pack[0] != key:
            cutoff, length = self.cutoff, self.filter_length
            cp = _round_up(cutoff, 128) if tc else _round4(cutoff)
            fwd = torch.zeros(2 * cp, length, dtype=torch.float32)
            fb = self.forward_basis.detach().float().cpu()[:, 0]
            fwd[:cutoff] = fb[:cutoff]
            fwd[cp: cp + cutoff] = fb[cutoff:]
            inv = torch.zeros(length, 2 * cp, dtype=torch.float32)          # W[n][k] = inverse_basis[k][n]
            ib = self.inverse_basis.detach().float().cpu()[:, 0]
            inv[:, :cutoff] = ib[:cutoff].t()
            inv[:, cp: cp + cutoff] = ib[cutoff:].t()
            fwd_paired = None
            if tc:
                # Re/Im-paired row order for the fused-epilogue GEMMs: pass p = Re rows of bins 128p..128p+127, then Im
                order = torch.cat([torch.cat([torch.arange(128 * p, 128 * (p + 1)), cp + torch.arange(128 * p, 128 * (p + 1))])
                                   for p in range(cp // 128)])
                fwd_paired = _split3(fwd[order]).to(device)
                fwd, inv = _split3(fwd), _split3(inv)
            if self.window is not None:
                sq = torch.from_numpy(padded_window(self.window, self.win_length, length) ** 2).double().to(device)
            else:
                sq = None          # window=None: no envelope division and no L/hop scale (stft.py:111-125)
            pair_pack = None
            if tc:
                # CTA-pair kernels: no padded bins.  Im of bins 0 and L/2 is exactly zero (stft.py:46-51), so the Re row of
                # bin L/2 rides in the Im slot of bin 0: pass p = Re rows of bins 128p..128p+127, then their Im rows
                half = length // 2
                re, im = fb[:cutoff], fb[cutoff:]
                paired = torch.empty(length, length, dtype=torch.float32)
                for p in range(length // 256):
                    paired[256 * p: 256 * p + 128] = re[128 * p: 128 * (p + 1)]
                    paired[256 * p + 128: 256 * (p + 1)] = im[128 * p: 128 * (p + 1)]
                paired[128] = re[half]
                # inverse basis with the matching K order: [Re 0..L/2-1 | Re L/2 | Im 1..L/2-1]
                inv2 = torch.empty(length, length, dtype=torch.float32)
                inv2[:, :half] = ib[:half].t()
                inv2[:, half] = ib[half]
                inv2[:, half + 1:] = ib[cutoff + 1: cutoff + half].t()
                ola = None
                taps = length // self.hop_length
                if self.hop_length % 256 == 0 and length % self.hop_length == 0 and taps <= 8:
                    # inverse GEMM with the overlap-add inside (wgb_tc2_istft_ola): tap j = inverse-basis samples
                    # j*hop .. (j+1)*hop-1 as a [hop, 3L] split-bf16 block; envelope per set of covering frames, summed
                    # like the reference's host loop (float32 += float64, frames ascending = taps descending)
                    hop = self.hop_length
                    w_ola = torch.cat([_split3(inv2[j * hop: (j + 1) * hop]) for j in range(taps)], dim=1).contiguous()
                    env = None
                    if self.window is not None:
                        sq64 = padded_window(self.window, self.win_length, length).astype(np.float64) ** 2
                        env = np.zeros((1 << taps, hop), dtype=np.float32)
                        for mask in range(1 << taps):
                            for j in reversed(range(taps)):
                                if (mask >> j) & 1:
                                    env[mask] = (env[mask].astype(np.float64) + sq64[j * hop: (j + 1) * hop]).astype(np.float32)
                        env = torch.from_numpy(env).to(device)
                    ola = (w_ola.to(device), env)
                pair_pack = (_split3(paired).to(device), _split3(inv2).to(device), ola)
            self._pack = (key, fwd.to(device), inv.to(device), sq, cp, fwd_paired, pair_pack)
        return self._pack[1:5]

    def _paired_basis(self, device):
        self._packed(device)
        return self._pack[5]

    def _pair_pack(self, device):
        """(forward basis, inverse basis, overlap-add pack) in the unpadded layout of the CTA-pair kernels: the bases as
        split-bf16 [L][3L]; the pack = (tap-sliced inverse basis [hop][taps*3L], envelope table [2^taps][hop] or None
        for window=None), or None when hop % 256 != 0."""
        self._packed(device)
        return self._pack[6]

    def _fft_pack(self, device):
        """(window fp32 [L] on device, overlap-add envelope table [2^taps][hop] or None) for the butterfly kernels, or None
        when they do not apply.  They assume what stft.py:46-60 builds: forward_basis = window * [cos; -sin](2 pi k n / L)
        and inverse_basis = its pseudo-inverse, which for this basis is the inverse real DFT c_k / (L scale) * [cos; -sin]
        (c = 1 for bins 0 and L/2, else 2) times the window.  Buffers that were edited or loaded from elsewhere are compared
        with those closed forms (1e-6 of the largest entry); anything else goes to the dense-basis kernels."""
        if self.algorithm == "gemm" or self.precision != "auto" or self.filter_length != 1024:
            if self.algorithm == "fft":
                raise RuntimeError("the butterfly STFT kernels need filter_length 1024 and precision 'auto'")
            return None
        key = (str(device), self.forward_basis.data_ptr(), self.forward_basis._version,
               self.inverse_basis.data_ptr(), self.inverse_basis._version)
        if self._fft is None or self._fft[0] != key:
            length, cutoff = self.filter_length, self.cutoff
            if self.window is not None:
                win = padded_window(self.window, self.win_length, length).astype(np.float32)
            else:
                win = np.ones(length, dtype=np.float32)
            kn = (np.outer(np.arange(cutoff), np.arange(length)) % length) * (2.0 * np.pi / length)
            basis = np.concatenate([np.cos(kn), -np.sin(kn)], axis=0)
            c = np.full((cutoff, 1), 2.0)
            c[0] = c[-1] = 1.0
            inv = np.concatenate([c, c], axis=0) * basis / (length * (length / self.hop_length))
            fb = self.forward_basis.detach().float().cpu().numpy()[:, 0]
            ib = self.inverse_basis.detach().float().cpu().numpy()[:, 0]
            stock = (fb.shape == basis.shape and ib.shape == inv.shape
                     and float(np.abs(fb - basis.astype(np.float32) * win).max()) <= 1e-6
                     and float(np.abs(ib - inv.astype(np.float32) * win).max()) <= 1e-6 * float(np.abs(inv).max()))
            pack = None
            if stock:
                env = None
                taps = length // self.hop_length
                if self.window is not None and length % self.hop_length == 0 and taps <= 8:
                    # window-sum envelope per set of covering frames, summed like the reference's host loop
                    # (audio_processing.py:45-47: float32 += float64, frames ascending = taps descending)
                    hop = self.hop_length
                    sq64 = win.astype(np.float64) ** 2
                    tab = np.zeros((1 << taps, hop), dtype=np.float32)
                    for mask in range(1 << taps):
                        for j in reversed(range(taps)):
                            if (mask >> j) & 1:
                                tab[mask] = (tab[mask].astype(np.float64) + sq64[j * hop: (j + 1) * hop]).astype(np.float32)
                    env = torch.from_numpy(tab).to(device)
                pack = (torch.from_numpy(win).to(device), env)
            self._fft = (key, pack)
        if self._fft[1] is None and self.algorithm == "fft":
            raise RuntimeError("the butterfly STFT kernels need the constructor's basis buffers")
        return self._fft[1]

    def _denoised_fft(self, y: torch.Tensor, bias_spec: torch.Tensor, strength: float):
        """Denoiser.forward as ONE butterfly kernel (wgb_fft_denoise), or None when that path does not apply."""
        pack = self._fft_pack(y.device)
        if pack is None or self.hop_length * 4 != self.filter_length or y.shape[1] <= self.filter_length // 2:
            return None
        _lib.require_b200(y.device)
        b, n = y.shape
        frames = n // self.hop_length + 1
        out = torch.empty((b, 1, self.hop_length * (frames - 1)), device=y.device, dtype=torch.float32)
        _lib.call("wgb_fft_denoise", y, pack[0], bias_spec, float(strength), pack[1] if self.window is not None else None,
                  out, b, n, self.hop_length, _lib.stream_ptr())
        return out

    def _mel_fft(self, y: torch.Tensor, mel_pack, n_mel: int, clip: float, range_flag=None):
        """TacotronSTFT.mel_spectrogram as ONE butterfly kernel (wgb_fft_stft_mel), or None when that path does not
        apply.  mel_pack = (slots int32 [S, 32], piece weights fp32 [S, 2, 32, 4], S, bins_used) from TacotronSTFT._mel_slots."""
        pack = self._fft_pack(y.device)
        if pack is None or mel_pack is None or y.shape[1] <= self.filter_length // 2:
            return None
        _lib.require_b200(y.device)
        b, n = y.shape
        out = torch.empty((b, n_mel, n // self.hop_length + 1), device=y.device, dtype=torch.float32)
        _lib.call("wgb_fft_stft_mel", y, pack[0], mel_pack[0], mel_pack[2], mel_pack[1], mel_pack[3], out, b, n,
                  self.hop_length, n_mel, float(clip), range_flag, _lib.stream_ptr())
        return out

    def _use_pair(self) -> bool:
        return self.pair and self._use_tc() and self.filter_length <= 1024      # per-bin tables of L/2 + 1 entries in smem

    # ------------------------------------------------------------------ device pipeline pieces
    def _spectrum(self, y: torch.Tensor):
        """y [B,N] fp32 cuda -> spec [B, F, 2cp] (Re | Im), F = N // hop + 1."""
        _lib.require_b200(y.device)
        fwd, _, _, cp = self._packed(y.device)
        b, n = y.shape
        length, hop = self.filter_length, self.hop_length
        frames = n // hop + 1
        s = _lib.stream_ptr()
        spec = torch.empty((b, frames, 2 * cp), device=y.device, dtype=torch.float32)
        if self._use_tc():
            # frames are overlapping rows (stride = hop) of the padded signal, read by TMA; operands split hi/lo
            hi, lo, ld_pad = self._padded_split(y)
            _lib.call("wgb_tc_gemm_split3", hi, lo, fwd, None, spec, b, frames, 2 * cp, length, hop, ld_pad, s)
            return spec, frames, cp
        if hop % 4 != 0 or length % 4 != 0:
            raise RuntimeError("hop_length and filter_length must be multiples of 4")
        ld_pad = _round4(n + length)
        ypad = torch.empty((b, ld_pad), device=y.device, dtype=torch.float32)
        _lib.call("wgb_stft_reflect_pad", y, ypad, b, n, length // 2, ld_pad, s)
        _lib.call("wgb_sgemm_f32", ypad, fwd, None, spec, 0, b, frames, 2 * cp, length,
                  hop, ld_pad, length, 2 * cp, frames * 2 * cp, 0, 0, s)
        return spec, frames, cp

    def _padded_split(self, y: torch.Tensor, whole_hops: bool = False, range_flag=None):
        """reflect-padded signal as bf16 hi / lo parts [B, ld_pad] (operands of the tensor-core STFT GEMMs).
        whole_hops: pitch = a whole number of hops, so that the frames of the whole batch form ONE row axis (frame r of
        utterance b = flat row b * ld_pad / hop + r) for the CTA-pair kernels."""
        b, n = y.shape
        ld_pad = _round_up(n + self.filter_length, self.hop_length if whole_hops else 8)
        hi = torch.empty((b, ld_pad), device=y.device, dtype=torch.bfloat16)
        lo = torch.empty_like(hi)
        if range_flag is not None:        # int32 device flag, zeroed by the caller: set when a sample is outside [-1, 1]
            _lib.call("wgb_stft_reflect_pad_split_check", y, hi, lo, b, n, self.filter_length // 2, ld_pad, range_flag,
                      _lib.stream_ptr())
        else:
            _lib.call("wgb_stft_reflect_pad_split", y, hi, lo, b, n, self.filter_length // 2, ld_pad, _lib.stream_ptr())
        return hi, lo, ld_pad

    def _magnitude_cl(self, y: torch.Tensor):
        """y [B,N] -> |STFT| channels-last fp32 [B, F, cp] straight from the GEMM epilogue (tensor-core path only)."""
        _lib.require_b200(y.device)
        _, _, _, cp = self._packed(y.device)
        b, n = y.shape
        frames = n // self.hop_length + 1
        hi, lo, ld_pad = self._padded_split(y)
        mag = torch.empty((b, frames, cp), device=y.device, dtype=torch.float32)
        _lib.call("wgb_tc_stft_mag", hi, lo, self._paired_basis(y.device), mag, b, frames, cp, self.filter_length,
                  self.hop_length, ld_pad, _lib.stream_ptr())
        return mag, frames, cp

    def _mel_fused(self, y: torch.Tensor, mel_table: torch.Tensor, n_mel: int, clip: float, range_flag=None) -> torch.Tensor:
        """y [B,N] -> log-mel [B, n_mel, F] in one GEMM kernel (TacotronSTFT.mel_spectrogram, layers.py:63-79)."""
        _lib.require_b200(y.device)
        _, _, _, cp = self._packed(y.device)
        b, n = y.shape
        frames = n // self.hop_length + 1
        out = torch.empty((b, n_mel, frames), device=y.device, dtype=torch.float32)
        if isinstance(mel_table, tuple):          # (table [L/2 + 1], n_pass): CTA-pair kernel, unpadded layout
            hi, lo, ld_pad = self._padded_split(y, whole_hops=True, range_flag=range_flag)
            _lib.call("wgb_tc2_stft_mel", hi, lo, self._pair_pack(y.device)[0], mel_table[0], out, b, frames,
                      ld_pad // self.hop_length, self.filter_length, self.hop_length, mel_table[1], n_mel, float(clip),
                      _lib.stream_ptr())
            return out
        hi, lo, ld_pad = self._padded_split(y, range_flag=range_flag)
        _lib.call("wgb_tc_stft_mel", hi, lo, self._paired_basis(y.device), mel_table, out, b, frames, cp,
                  self.filter_length, self.hop_length, ld_pad, n_mel, float(clip), _lib.stream_ptr())
        return out

    def _denoised(self, y: torch.Tensor, bias_spec: torch.Tensor, strength: float) -> torch.Tensor:
        """Denoiser.forward on the tensor-core path: STFT GEMM with the spectral subtraction in its epilogue (writes
        the inverse GEMM's bf16 hi / lo operands), inverse GEMM, overlap-add."""
        _lib.require_b200(y.device)
        _, inv, win_sq, cp = self._packed(y.device)
        b, n = y.shape
        length, hop = self.filter_length, self.hop_length
        frames = n // hop + 1
        s = _lib.stream_ptr()
        if self._use_pair():                      # CTA-pair kernels, unpadded layout: K of the inverse GEMM = L
            hi, lo, ld_pad = self._padded_split(y, whole_hops=True)
            fwd2, inv2, ola = self._pair_pack(y.device)
            out = torch.empty((b, 1, hop * (frames - 1)), device=y.device, dtype=torch.float32)
            if ola is not None and self.fused_ola:
                # inverse-basis GEMM with the overlap-add, envelope, scale and trim in it: the spectra carry taps - 1
                # zero guard rows per utterance (the frames before the first / after the last one)
                rp = frames + length // hop - 1
                shi = torch.empty((b, rp, length), device=y.device, dtype=torch.bfloat16)
                slo = torch.empty_like(shi)
                shi[:, frames:].zero_()
                slo[:, frames:].zero_()
                _lib.call("wgb_tc2_stft_denoise", hi, lo, fwd2, bias_spec, float(strength), shi, slo, b, frames,
                          ld_pad // hop, length, hop, rp, s)
                _lib.call("wgb_tc2_istft_ola", shi, slo, ola[0], ola[1], out, b, frames, length, hop, s)
                return out
            shi = torch.empty((b * frames, length), device=y.device, dtype=torch.bfloat16)
            slo = torch.empty_like(shi)
            _lib.call("wgb_tc2_stft_denoise", hi, lo, fwd2, bias_spec, float(strength), shi, slo, b, frames, ld_pad // hop,
                      length, hop, frames, s)
            fr = torch.empty((b, frames, length), device=y.device, dtype=torch.float32)
            _lib.call("wgb_tc2_gemm_split3", shi, slo, inv2, fr, b * frames, length, length, s)
            _lib.call("wgb_istft_overlap_add", fr, win_sq, out, b, frames, length, hop, s)
            return out
        hi, lo, ld_pad = self._padded_split(y)
        shi = torch.empty((b * frames, 2 * cp), device=y.device, dtype=torch.bfloat16)
        slo = torch.empty_like(shi)
        _lib.call("wgb_tc_stft_denoise", hi, lo, self._paired_basis(y.device), bias_spec, float(strength), shi, slo, b,
                  frames, self.cutoff, cp, length, hop, ld_pad, s)
        fr = torch.empty((b, frames, length), device=y.device, dtype=torch.float32)
        _lib.call("wgb_tc_gemm_split3", shi, slo, inv, None, fr, 1, b * frames, length, 2 * cp, 2 * cp,
                  b * frames * 2 * cp, s)
        out = torch.empty((b, 1, hop * (frames - 1)), device=y.device, dtype=torch.float32)
        _lib.call("wgb_istft_overlap_add", fr, win_sq, out, b, frames, length, hop, s)
        return out

    def _synthesize(self, spec: torch.Tensor, frames: int, cp: int, denoise=None) -> torch.Tensor:
        """spec [B, F, 2cp] -> [B, 1, hop*(F-1)] (stft.py:105-128).  denoise = (bias_spec [cutoff], strength): apply
        the Denoiser's spectral subtraction on the way (denoiser.py:36-38), fused with the operand split."""
        _, inv, win_sq, _ = self._packed(spec.device)
        b = spec.shape[0]
        length, hop = self.filter_length, self.hop_length
        s = _lib.stream_ptr()
        fr = torch.empty((b, frames, length), device=spec.device, dtype=torch.float32)
        if denoise is not None and not self._use_tc():
            _lib.call("wgb_denoise_scale", spec, denoise[0], float(denoise[1]), b * frames, self.cutoff, cp, s)
        if self._use_tc():
            hi = torch.empty((b * frames, 2 * cp), device=spec.device, dtype=torch.bfloat16)
            lo = torch.empty_like(hi)
            if denoise is not None:
                _lib.call("wgb_denoise_scale_split", spec, denoise[0], float(denoise[1]), hi, lo, b * frames, self.cutoff,
                          cp, s)
            else:
                _lib.call("wgb_split_bf16", spec, hi, lo, spec.numel(), s)
            _lib.call("wgb_tc_gemm_split3", hi, lo, inv, None, fr, 1, b * frames, length, 2 * cp, 2 * cp,
                      b * frames * 2 * cp, s)
        else:
            _lib.call("wgb_sgemm_f32", spec, inv, None, fr, 0, 1, b * frames, length, 2 * cp,
                      2 * cp, 0, 2 * cp, length, 0, 0, 0, s)
        out = torch.empty((b, 1, hop * (frames - 1)), device=spec.device, dtype=torch.float32)
        _lib.call("wgb_istft_overlap_add", fr, win_sq, out, b, frames, length, hop, s)
        return out

    # ------------------------------------------------------------------ reference API
    def transform(self, input_data: torch.Tensor):
        if not input_data.is_cuda:
            raise RuntimeError("STFT.transform needs a CUDA tensor on a B200; there is no CPU fallback")
        self.num_samples = input_data.size(1)
        y = input_data.float().contiguous()
        with torch.cuda.device(y.device):
            spec, frames, cp = self._spectrum(y)
            b = y.shape[0]
            mag = torch.empty((b, self.cutoff, frames), device=y.device, dtype=torch.float32)
            phase = torch.empty_like(mag)
            _lib.call("wgb_stft_polar", spec, mag, phase, None, b, frames, self.cutoff, cp, _lib.stream_ptr())
        return mag, phase

    def inverse(self, magnitude: torch.Tensor, phase: torch.Tensor):
        if not magnitude.is_cuda:
            raise RuntimeError("STFT.inverse needs CUDA tensors on a B200; there is no CPU fallback")
        _lib.require_b200(magnitude.device)
        with torch.cuda.device(magnitude.device):
            _, _, _, cp = self._packed(magnitude.device)
            b, cutoff, frames = magnitude.shape
            spec = torch.empty((b, frames, 2 * cp), device=magnitude.device, dtype=torch.float32)
            _lib.call("wgb_stft_recombine", magnitude.float().contiguous(), phase.float().contiguous(), spec, b, frames,
                      cutoff, cp, _lib.stream_ptr())
            return self._synthesize(spec, frames, cp)

    def forward(self, input_data):
        self.magnitude, self.phase = self.transform(input_data)
        return self.inverse(self.magnitude, self.phase)
