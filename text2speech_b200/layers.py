"""Drop-in TacotronSTFT (reference: utils/layers.py:42-79): STFT -> mel matmul -> log-clamp on the GPU.

``LinearNorm`` / ``ConvNorm`` (layers.py:8-39) belong to Tacotron-2 and are out of scope.
"""
from __future__ import annotations

import torch

from . import _lib
from .audio_processing import dynamic_range_compression, dynamic_range_decompression, mel_filterbank
from .stft import STFT, _round4


class TacotronSTFT(torch.nn.Module):
    def __init__(self, filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=80, sampling_rate=44800,
                 mel_fmin=0.0, mel_fmax=8000.0):
        super().__init__()
        self.n_mel_channels = n_mel_channels
        self.sampling_rate = sampling_rate
        self.stft_fn = STFT(filter_length, hop_length, win_length)
        self.register_buffer("mel_basis", mel_filterbank(sampling_rate, filter_length, n_mel_channels, mel_fmin, mel_fmax))
        self._mel_pack = None
        self.fused = True          # False: keep the mel matmul / log as separate kernels (A/B and validation)

    def spectral_normalize(self, magnitudes):
        return dynamic_range_compression(magnitudes)

    def spectral_de_normalize(self, magnitudes):
        return dynamic_range_decompression(magnitudes)

    def _mel_packed(self, device, cp):
        """(mel basis [n_mel, k_used] fp32 on device, k_used): only the leading bins that carry any weight take part
        in the matmul (bins above mel_fmax are exactly zero: 372 of 513 at 22.05 kHz / 8 kHz)."""
        key = (str(device), cp, self.mel_basis.data_ptr(), self.mel_basis._version)
        if self._mel_pack is None or self._mel_pack[0] != key:
            basis = self.mel_basis.detach().float().cpu()
            nz = torch.nonzero(basis.abs().sum(0))
            k_used = min(cp, _round4(int(nz.max()) + 1 if nz.numel() else 4))
            w = torch.zeros((self.n_mel_channels, k_used), dtype=torch.float32)
            n = min(k_used, basis.shape[1])
            w[:, :n] = basis[:, :n]
            self._mel_pack = (key, w.to(device), k_used)
        return self._mel_pack[1], self._mel_pack[2]

    def _mel_table(self, device, cp):
        """Sparse form of mel_basis for the fused kernel: per bin {first filter index, weight in it, weight in the next
        filter, 0}.  Triangular filters overlap pairwise, so a bin feeds at most two ADJACENT filters; returns None if
        this basis does not have that structure (a hand-edited mel_basis) or has more than 80 filters."""
        key = (str(device), cp, self.mel_basis.data_ptr(), self.mel_basis._version)
        cached = getattr(self, "_mel_tab", None)
        if cached is None or cached[0] != key:
            basis = self.mel_basis.detach().float().cpu()
            n_mel, n_bins = basis.shape
            tab = torch.zeros((cp, 4), dtype=torch.float32)
            ok = n_mel <= 80 and n_bins <= cp
            for k in range(min(n_bins, cp)) if ok else ():
                nz = torch.nonzero(basis[:, k]).flatten()
                if nz.numel() == 0:
                    continue
                m0 = int(nz[0])
                if nz.numel() > 2 or (nz.numel() == 2 and int(nz[1]) != m0 + 1):
                    ok = False
                    break
                tab[k, 0], tab[k, 1] = float(m0), basis[m0, k]
                if nz.numel() == 2:
                    tab[k, 2] = basis[m0 + 1, k]
            # passes (of 128 bins) of the CTA-pair kernel that hold a bin with non-zero weight; its layout has exactly
            # L/2 + 1 entries and the last bin (L/2) rides in pass 0
            nz = torch.nonzero(tab[: max(cp - 1, 1), 1:3].abs().sum(1)).flatten()
            n_pass = max(1, -(-(int(nz.max()) + 1) // 128)) if nz.numel() else 1
            self._mel_tab = (key, tab.to(device) if ok else None, n_pass)
        return self._mel_tab[1]

    def _mel_table_pair(self, device):
        """(table [L/2 + 1] float4, n_pass) for wgb_tc2_stft_mel, or None when the basis lacks the structure that kernel
        streams over (a bin feeds at most two ADJACENT filters, and over the bins that carry weight -- the last bin L/2
        included -- the first-filter index never decreases) or the filter is longer than its
        shared-memory table (L/2 > 512); the one-CTA kernel, which scatters into per-row accumulators, takes those."""
        cutoff = self.stft_fn.cutoff
        if cutoff - 1 > 512:
            return None
        tab = self._mel_table(device, cutoff)
        if tab is None:
            return None
        cached = getattr(self, "_mel_pair_ok", None)
        if cached is None or cached[0] is not tab:
            t = tab.detach().cpu()
            has = (t[:, 1] != 0) | (t[:, 2] != 0)
            idx = t[:, 0][has].long()                  # first-filter index of the bins with weight, ascending bins
            ok = idx.numel() > 0 and bool(((idx[1:] - idx[:-1]) >= 0).all())
            self._mel_pair_ok = (tab, ok)
            cached = self._mel_pair_ok
        return (tab, self._mel_tab[2]) if cached[1] else None

    def _mel_slots(self, device):
        """mel_basis for wgb_fft_stft_mel: (slots int32 [S, 32], piece weights fp32 [S, 2, 32, 4], S, bins the pieces with
        weight reach), or None when the basis does not fit the kernel's tables.  A filter's non-zero span is covered by
        pieces of 8 bins that start at multiples of 4 (two aligned 16-byte reads of |X| each; weights zero outside the span);
        whole filters are dealt to the 32 lanes of a warp, most pieces first onto the least loaded lane, and lane l's
        slots [q][l] list its filters' pieces in order, packed as first bin / 4 | (filter to emit after this piece + 1) << 8
        | (starts a filter) << 16.  Filters without any weight get one all-zero piece (they emit log(clip), like the
        reference's matmul); unused slots have zero weights and emit nothing."""
        key = (str(device), self.mel_basis.data_ptr(), self.mel_basis._version)
        cached = getattr(self, "_mel_slots_pack", None)
        if cached is None or cached[0] != key:
            basis = self.mel_basis.detach().float().cpu()
            n_mel, n_bins = basis.shape
            pack = None
            if n_mel <= 128 and n_bins == self.stft_fn.cutoff and n_bins == 513:
                filters = []                                   # (number of pieces, filter, [(bin4, weights[8]), ...])
                padded = torch.zeros((n_mel, 544))
                padded[:, :n_bins] = basis
                bins_used = 1
                for m in range(n_mel):
                    nz = torch.nonzero(basis[m]).flatten()
                    if nz.numel() == 0:
                        filters.append((1, m, [(0, torch.zeros(8))]))
                        continue
                    lo, hi = int(nz[0]), int(nz[-1]) + 1
                    pieces = []
                    for start in range(lo // 4 * 4, hi, 8):
                        w = padded[m, start: start + 8].clone()
                        w[: max(lo - start, 0)] = 0.0
                        if hi - start < 8:
                            w[hi - start:] = 0.0
                        pieces.append((start // 4, w))
                        bins_used = max(bins_used, start + 8)
                    filters.append((len(pieces), m, pieces))
                lanes, load = [[] for _ in range(32)], [0] * 32
                for cnt, m, pieces in sorted(filters, key=lambda f: (-f[0], f[1])):
                    i = min(range(32), key=lambda j: (load[j], j))
                    for q, (bin4, w) in enumerate(pieces):
                        lanes[i].append((bin4 | ((m + 1 if q == cnt - 1 else 0) << 8) | ((1 if q == 0 else 0) << 16), w))
                    load[i] += cnt
                per_lane = max(load)
                if per_lane <= 16:
                    table = torch.zeros((per_lane, 32), dtype=torch.int32)
                    weights = torch.zeros((per_lane, 2, 32, 4), dtype=torch.float32)
                    for i, lst in enumerate(lanes):
                        for q, (word, w) in enumerate(lst):
                            table[q, i] = word
                            weights[q, 0, i], weights[q, 1, i] = w[:4], w[4:]
                    pack = (table.contiguous().to(device), weights.contiguous().to(device), per_lane, min(bins_used, n_bins))
            self._mel_slots_pack = (key, pack)
            cached = self._mel_slots_pack
        return cached[1]

    def mel_spectrogram(self, y: torch.Tensor) -> torch.Tensor:
        """y [B, T] in [-1, 1] (CUDA) -> log-mel [B, n_mel_channels, T // hop + 1].  The reference's two input asserts
        (layers.py:72-73: min(y) >= -1, max(y) <= 1) raise AssertionError here too; on the fused tensor-core path they are
        one flag written by the reflect-pad pass and read once, instead of two device-wide reductions with a host sync each."""
        if not y.is_cuda:
            raise RuntimeError("TacotronSTFT.mel_spectrogram needs a CUDA tensor on a B200; there is no CPU fallback")
        with torch.cuda.device(y.device):
            if self.fused and self.stft_fn._use_tc() and y.dtype == torch.float32:
                flag = torch.zeros(1, device=y.device, dtype=torch.int32)
                out = self._mel_spectrogram(y, flag)
                if int(flag.item()) != 0:
                    assert torch.min(y.data) >= -1                             # layers.py:72-73
                    assert torch.max(y.data) <= 1
                    raise AssertionError("input outside [-1, 1]")              # NaN input
                return out
            assert torch.min(y.data) >= -1                                     # layers.py:72-73
            assert torch.max(y.data) <= 1
            return self._mel_spectrogram(y)

    def _mel_spectrogram(self, y: torch.Tensor, range_flag=None) -> torch.Tensor:
        y = y.float().contiguous()
        b = y.shape[0]
        s = _lib.stream_ptr()
        if self.fused:                        # stock STFT bases: FFT, |X|, filterbank and log-clamp in one butterfly kernel
            out = self.stft_fn._mel_fft(y, self._mel_slots(y.device), self.n_mel_channels, 1e-5, range_flag)
            if out is not None:
                return out
        if self.stft_fn._use_tc():
            cp = self.stft_fn._packed(y.device)[3]
            table = None
            if self.fused:
                table = self._mel_table_pair(y.device) if self.stft_fn._use_pair() else None
                if table is None:
                    table = self._mel_table(y.device, cp)
            if table is not None:             # STFT GEMM with |X|, the mel filterbank and log-clamp in its epilogue
                return self.stft_fn._mel_fused(y, table, self.n_mel_channels, 1e-5, range_flag)
            if range_flag is not None:        # un-fused fallback: the flag would stay unwritten -> check the slow way
                assert torch.min(y.data) >= -1 and torch.max(y.data) <= 1
            mag_cl, frames, cp = self.stft_fn._magnitude_cl(y)   # |X| straight from the STFT GEMM's epilogue
        else:
            spec, frames, cp = self.stft_fn._spectrum(y)
            mag_cl = torch.empty((b, frames, cp), device=y.device, dtype=torch.float32)
            _lib.call("wgb_stft_polar", spec, None, None, mag_cl, b, frames, self.stft_fn.cutoff, cp, s)
        raw = torch.empty((b, frames, self.n_mel_channels), device=y.device, dtype=torch.float32)
        w_mel, k_used = self._mel_packed(y.device, cp)
        _lib.call("wgb_sgemm_f32", mag_cl, w_mel, None, raw, 0, 1, b * frames,
                  self.n_mel_channels, k_used, cp, 0, k_used, self.n_mel_channels, 0, 0, 0, s)
        out = torch.empty((b, self.n_mel_channels, frames), device=y.device, dtype=torch.float32)
        _lib.call("wgb_mel_log", raw, out, b, frames, self.n_mel_channels, 1e-5, s)
        return out
