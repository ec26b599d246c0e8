"""Host-side packing of the CTA-pair STFT kernels (csrc/stft_tc2.cu), emulated on the CPU from the PACKED operands:
the unpadded spectrum layout (Re row of bin L/2 in the Im slot of bin 0), the matching inverse-basis K order and the
mel table / pass count must reproduce the oracle's STFT, mel spectrogram and denoiser."""
import numpy as np
import torch

import oracle
from tests import util
from text2speech_b200 import synthetic as syn
from text2speech_b200.layers import TacotronSTFT
from text2speech_b200.stft import STFT

DC = syn.DEFAULT_DATA_CONFIG
CPU = torch.device("cpu")


def _unsplit(w3):
    k = w3.shape[1] // 3
    assert torch.equal(w3[:, :k], w3[:, k:2 * k])
    return w3[:, :k].double() + w3[:, 2 * k:].double()


def _frames(y, length, hop):
    padded = torch.nn.functional.pad(y[:, None], (length // 2, length // 2), mode="reflect")[:, 0]
    return padded.unfold(1, length, hop).double()                         # [B, F, L]


def _spectrum_from_paired(s, length):
    """paired GEMM output [.., L] -> (re [.., L/2+1], im [.., L/2+1]) as the kernels' epilogues read it."""
    half = length // 2
    re = torch.zeros(s.shape[:-1] + (half + 1,), dtype=s.dtype)
    im = torch.zeros_like(re)
    for p in range(length // 256):
        re[..., 128 * p: 128 * (p + 1)] = s[..., 256 * p: 256 * p + 128]
        im[..., 128 * p: 128 * (p + 1)] = s[..., 256 * p + 128: 256 * (p + 1)]
    re[..., half] = im[..., 0]                                            # Nyquist rides in the Im slot of bin 0
    im[..., 0] = 0.0
    return re, im


def test_pair_layout_reproduces_transform_mel_and_denoiser(golden):
    length, hop = DC["filter_length"], DC["hop_length"]
    stft = STFT(length, hop, DC["win_length"])
    fwd3, inv3, ola = stft._pair_pack(CPU)
    assert fwd3.shape == (length, 3 * length) and inv3.shape == (length, 3 * length) and fwd3.dtype == torch.bfloat16
    y = syn.synthetic_waveforms(2, 4096, sr=DC["sampling_rate"], seed=5)
    fr = _frames(y, length, hop)
    re, im = _spectrum_from_paired(fr @ _unsplit(fwd3).t(), length)
    mag = torch.sqrt(re * re + im * im).permute(0, 2, 1)
    assert util.rel_l2(mag, golden["stft_mag"]) < 5e-6          # hi + lo of the bf16 split keep ~16 bits of the basis

    # mel: table of L/2 + 1 entries, only the first n_pass passes of 128 bins are computed
    taco = TacotronSTFT(length, hop, DC["win_length"], 80, DC["sampling_rate"], DC["mel_fmin"], DC["mel_fmax"])
    table, n_pass = taco._mel_table_pair(CPU)
    assert table.shape == (length // 2 + 1, 4) and n_pass == 3            # bins above 371 carry no weight at 8 kHz / 22.05 kHz
    mag_used = mag.permute(0, 2, 1).clone()
    mag_used[..., 128 * n_pass: length // 2] = 0.0                        # passes the kernel never runs (Nyquist is in pass 0)
    acc = torch.zeros(mag_used.shape[:2] + (82,), dtype=torch.float64)
    for k in range(length // 2 + 1):
        m0 = int(table[k, 0])
        acc[..., m0] += float(table[k, 1]) * mag_used[..., k]
        acc[..., m0 + 1] += float(table[k, 2]) * mag_used[..., k]
    mel = torch.log(torch.clamp(acc[..., :80], min=1e-5)).permute(0, 2, 1)
    assert float((mel - torch.from_numpy(golden["mel"]).double()).abs().max()) < 1e-4

    # denoiser: epilogue columns [Re 0..L/2-1 | Re L/2 | Im 1..L/2-1] against the packed inverse basis
    bias = torch.from_numpy(golden["denoiser_bias_spec"]).double().reshape(-1)
    for strength, key in ((0.1, "denoised_s0p1"), (0.01, "denoised_s0p01")):
        m = torch.sqrt(re * re + im * im)
        m2 = torch.clamp(m - bias * strength, min=0.0)
        g = torch.where(m > 0, m2 / m.clamp_min(1e-300), torch.zeros_like(m))
        re2 = torch.where(m > 0, re * g, m2)
        im2 = im * g
        half = length // 2
        cols = torch.cat([re2[..., :half], re2[..., half:half + 1], im2[..., 1:half]], dim=-1)
        assert cols.shape[-1] == length
        out_fr = cols @ _unsplit(inv3).t()                                # [B, F, L]
        frames = out_fr.shape[1]
        total = hop * (frames - 1) + length
        out = torch.zeros(out_fr.shape[0], total, dtype=torch.float64)
        for f in range(frames):
            out[:, f * hop: f * hop + length] += out_fr[:, f]
        env = torch.from_numpy(oracle.window_sumsquare("hann", frames, hop, DC["win_length"], length)).double()
        nz = env > np.finfo(np.float32).tiny
        out[:, nz] = out[:, nz] / env[nz]
        out = out * (length / hop)
        out = out[:, length // 2: total - length // 2]
        assert util.rel_l2(out[:, None], golden[key]) < 1e-5
        # the same through the overlap-add GEMM (wgb_tc2_istft_ola): block q = sum_j frame[q - j] W_j over zero-guarded
        # rows, envelope row picked by the set of covering frames, scale, blocks inside the trimmed L/2 dropped
        w_ola, env = ola
        taps = length // hop
        assert w_ola.shape == (hop, taps * 3 * length) and env.shape == (1 << taps, hop)
        rp = frames + taps - 1
        guarded = torch.zeros(cols.shape[0], rp, length, dtype=torch.float64)
        guarded[:, :frames] = cols
        blocks = torch.zeros(cols.shape[0], rp, hop, dtype=torch.float64)
        for j in range(taps):
            wj = _unsplit(w_ola[:, j * 3 * length: (j + 1) * 3 * length])            # [hop, L]
            shifted = torch.zeros_like(guarded)
            shifted[:, j:] = guarded[:, : rp - j]                                    # row q reads frame q - j
            blocks += shifted @ wj.t()
        got = torch.zeros(cols.shape[0], hop * (frames - 1), dtype=torch.float64)
        for q in range(taps // 2, frames - 2 + taps - taps // 2 + 1):
            mask = sum(1 << j for j in range(taps) if 0 <= q - j < frames)
            e = env[mask].double()
            blk = torch.where(e > np.finfo(np.float32).tiny, blocks[:, q] / e, blocks[:, q]) * (length / hop)
            got[:, (q - taps // 2) * hop: (q - taps // 2 + 1) * hop] = blk
        assert util.rel_l2(got[:, None], golden[key]) < 1e-5
        assert util.rel_l2(got, out) < 1e-6


def test_pair_layout_nyquist_weight_is_honoured():
    """A filterbank that reaches the Nyquist bin (fmax = sr / 2): its weight sits in the table's last entry, which the
    kernel applies to the value in the Im slot of bin 0, and all four passes run."""
    taco = TacotronSTFT(1024, 256, 1024, 80, 22050, 0.0, 11025.0)
    table, n_pass = taco._mel_table_pair(CPU)
    basis = taco.mel_basis
    assert n_pass == 4
    for k in (0, 1, 371, 511, 512):
        m0 = int(table[k, 0])
        assert float(table[k, 1]) == float(basis[m0, k])
        assert float(table[k, 2]) == (float(basis[m0 + 1, k]) if m0 + 1 < 80 else 0.0)


def _stream_mel(table, mag, n_pass, n_mel=80, clip=1e-5):
    """The kernel's streaming form of the filterbank (csrc/stft_tc2.cu EPI_MEL) for one row of magnitudes [L/2 + 1]:
    two running sums, a filter is emitted when the table's first-filter index moves past it, the Nyquist bin joins last."""
    cp = table.shape[0] - 1
    out = np.full(n_mel, np.nan)
    cur, a0, a1 = 0, 0.0, 0.0
    for k in range(128 * n_pass):
        m0 = int(table[k, 0]) if (table[k, 1] != 0 or table[k, 2] != 0) else -1
        if m0 > cur:
            out[cur] = np.log(max(a0, clip))
            if m0 == cur + 1:
                a0 = a1
            else:
                out[cur + 1] = np.log(max(a1, clip))
                out[cur + 2: m0] = np.log(clip)
                a0 = 0.0
            a1, cur = 0.0, m0
        a0 += float(table[k, 1]) * mag[k]
        a1 += float(table[k, 2]) * mag[k]
    mn = int(table[cp, 0]) if (table[cp, 1] != 0 or table[cp, 2] != 0) else -1
    for m in range(cur, n_mel):
        v = a0 if m == cur else (a1 if m == cur + 1 else 0.0)
        if m == mn:
            v += float(table[cp, 1]) * mag[cp]
        if m == mn + 1 and mn >= 0:
            v += float(table[cp, 2]) * mag[cp]
        out[m] = np.log(max(v, clip))
    return out


def test_streaming_filterbank_equals_dense_matmul():
    g = np.random.default_rng(0)
    for fmax, sr in ((8000.0, 22050), (11025.0, 22050), (7600.0, 16000), (8000.0, 44800)):
        taco = TacotronSTFT(1024, 256, 1024, 80, sr, 0.0, fmax)
        pair = taco._mel_table_pair(CPU)
        assert pair is not None, (fmax, sr)
        table, n_pass = pair
        mag = np.abs(g.standard_normal(513)) + 0.01
        want = np.log(np.maximum(taco.mel_basis.double().numpy() @ mag, 1e-5))
        got = _stream_mel(table.double().numpy(), mag, n_pass)
        assert not np.isnan(got).any() and np.abs(got - want).max() < 1e-6, (fmax, sr)
    # a hand-edited basis without the structure is refused (the one-CTA kernel takes it)
    taco = TacotronSTFT(1024, 256, 1024, 80, 22050, 0.0, 8000.0)
    taco.mel_basis[5, 300] = 0.1
    assert taco._mel_table_pair(CPU) is None


def test_ola_envelope_table_equals_reference_window_sum():
    """env[mask] of the overlap-add GEMM = audio_processing.py:7-48's window_sumsquare at the positions whose covering
    frames are that set -- bit for bit (float32 running sum of float64 squares, frames ascending)."""
    stft = STFT(1024, 256, 1024)
    _, _, (w_ola, env) = stft._pair_pack(CPU)
    frames, hop, taps = 9, 256, 4
    wss = oracle.window_sumsquare("hann", frames, hop, 1024, 1024)           # [hop * (frames - 1) + L]
    for q in range(frames + taps - 1):
        mask = sum(1 << j for j in range(taps) if 0 <= q - j < frames)
        assert np.array_equal(env[mask].numpy(), wss[q * hop: (q + 1) * hop]), q
    nowin = STFT(1024, 256, 1024, window=None)
    assert nowin._pair_pack(CPU)[2][1] is None
