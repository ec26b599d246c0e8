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


# ------------------------------------------------------------------------------------ butterfly path (csrc/fft.cu)

def _fft_mel_emulated(y, win, slots, weights, n_mel, hop, clip=1e-5):
    """What wgb_fft_stft_mel computes, from ITS operands: reflect-indexed frames x window -> real FFT -> |X| (zeros past
    bin 512) -> per lane, the slots [q][lane] in order: sum restarts on a filter's first piece, a piece = 8 bins from
    4 * (word & 255) times weights[q, :, lane], the running sum is the filter's value after its last piece -> log(clamp)."""
    fr = _frames(y, 1024, hop) * win.double()
    mag = torch.fft.rfft(fr, dim=-1).abs()                                 # [B, F, 513]
    mag = torch.cat([mag, torch.zeros(mag.shape[:2] + (15,), dtype=mag.dtype)], dim=-1)      # the kernel's 528-entry buffer
    out = torch.full((mag.shape[0], n_mel, mag.shape[1]), float("nan"), dtype=torch.float64)
    for lane in range(32):
        total = torch.zeros(mag.shape[:2], dtype=torch.float64)
        for q in range(slots.shape[0]):
            word = int(slots[q, lane])
            bin4, emit, first = word & 255, ((word >> 8) & 255) - 1, word >> 16
            if first:
                total = torch.zeros_like(total)
            total = total + (mag[..., 4 * bin4: 4 * bin4 + 8] * weights[q, :, lane].reshape(8).double()).sum(-1)
            if emit >= 0:
                assert torch.isnan(out[:, emit]).all()                     # every filter is emitted exactly once
                out[:, emit] = total
    return torch.log(torch.clamp(out, min=clip))


def _fft_denoise_emulated(y, win, env, bias, strength, hop, scale):
    """What wgb_fft_denoise computes: per frame rfft -> clamp(|X| - bias * strength, 0) with the phase kept -> irfft x
    window / scale; output block q (hop samples of the padded time line) = taps j = 3..0 of frames q - j, normalised with
    the envelope row of the set of existing frames, scaled by L / hop; blocks 2 .. frames are kept (L/2 trim)."""
    length = 1024
    fr = _frames(y, length, hop) * win.double()
    X = torch.fft.rfft(fr, dim=-1)
    mag = X.abs()
    g = torch.where(mag > 0, torch.clamp(mag - bias.double() * strength, min=0.0) / mag.clamp_min(1e-300), torch.zeros_like(mag))
    X = X * g
    x = torch.fft.irfft(X, n=length, dim=-1) * win.double() / scale       # [B, F, L]
    B, F, _ = x.shape
    out = torch.zeros(B, 1, hop * (F - 1), dtype=torch.float64)
    for q in range(2, F + 1):
        acc = torch.zeros(B, hop, dtype=torch.float64)
        mask = 0
        for j in (3, 2, 1, 0):                                            # oldest frame first, as the register pipeline
            r = q - j
            if 0 <= r < F:
                acc = acc + x[:, r, j * hop: (j + 1) * hop]
                mask |= 1 << j
        if env is not None:
            e = env[mask].double()
            acc = torch.where(e > np.finfo(np.float32).tiny, acc / e, acc) * scale
        out[:, 0, (q - 2) * hop: (q - 1) * hop] = acc
    return out


def mb_check(taco):
    return taco.mel_basis.abs().sum(0)


def test_fft_path_operands_reproduce_the_oracle():
    """Host side of the butterfly path on the CPU: the stock-basis test, the window / envelope tables of STFT._fft_pack
    and the banded mel rows of TacotronSTFT._mel_rows, pushed through a torch.fft emulation of the kernels' arithmetic,
    reproduce the oracle's conv-basis mel spectrogram and denoiser (ragged length, window shorter than the filter,
    window=None)."""
    y = syn.synthetic_waveforms(2, 256 * 11 + 77, sr=DC["sampling_rate"], seed=9)
    bias = torch.rand(513, generator=torch.Generator().manual_seed(1)) * 0.05
    for window, win_length, hop in (("hann", 1024, 256), ("hann", 800, 256), (None, 1024, 256), ("hann", 1024, 200)):
        stft = STFT(1024, hop, win_length, window=window)
        pack = stft._fft_pack(CPU)
        assert pack is not None
        win, env = pack
        fwd, inv = oracle.stft_bases(1024, hop, win_length, window=window)
        if window is not None:
            taco = TacotronSTFT(1024, hop, win_length, 80, DC["sampling_rate"], DC["mel_fmin"], DC["mel_fmax"])
            slots, weights, per_lane, bins_used = taco._mel_slots(CPU)
            assert bins_used == 376 and int(torch.nonzero(mb_check(taco)).max()) < bins_used     # 8 kHz at 22.05 kHz: bin 371
            assert slots.dtype == torch.int32 and slots.shape == (per_lane, 32) and per_lane <= 16
            assert weights.shape == (per_lane, 2, 32, 4)
            assert int((slots & 255).max()) * 4 + 8 <= 528                # the last piece stays inside the |X| buffer
            emitted = ((slots >> 8) & 255).flatten()
            assert sorted((emitted[emitted > 0] - 1).tolist()) == list(range(80))
            used = int(((weights.abs().sum((1, 3)) > 0) | (((slots >> 8) & 255) > 0)).sum())
            assert per_lane <= 2 + used // 32                              # balanced across the lanes
            mb = torch.from_numpy(oracle.mel_filterbank(DC["sampling_rate"], 1024, 80, DC["mel_fmin"], DC["mel_fmax"])).float()
            want = oracle.mel_spectrogram(y, fwd, mb, hop)
            got = _fft_mel_emulated(y, win, slots, weights, 80, hop)
            assert float((got - want).abs().max()) <= 2e-4, (window, win_length, hop)
        if hop * 4 == 1024:
            mag, phase = oracle.stft_transform(y, fwd, hop)
            want = oracle.stft_inverse(torch.clamp(mag - bias.reshape(1, 513, 1) * 0.5, 0.0), phase, inv, hop, win_length,
                                       window=window)
            assert (env is None) == (window is None)
            got = _fft_denoise_emulated(y, win, env, bias, 0.5, hop, 1024 / hop)
            assert got.shape == want.shape and util.snr_db(got, want) >= 100.0, (window, win_length)


def test_fft_path_refuses_edited_or_foreign_bases():
    """The butterfly kernels assume the constructor's real-DFT pair: an edited buffer, a basis of another filter length, a
    forced precision or algorithm = 'gemm' all leave the call on the dense-basis kernels; the reference's own bases
    (np.fft.fft(eye) / pinv, what a reference checkpoint would carry) are recognised as stock."""
    stft = STFT(1024, 256, 1024)
    assert stft._fft_pack(CPU) is not None
    fwd, inv = oracle.stft_bases(1024, 256, 1024)                          # the reference's construction (stft.py:45-66)
    stft.forward_basis.copy_(fwd)
    stft.inverse_basis.copy_(inv)
    assert stft._fft_pack(CPU) is not None
    stft.inverse_basis[5, 0, 100] += 1e-4
    assert stft._fft_pack(CPU) is None
    stft = STFT(1024, 256, 1024)
    stft.precision = "fp32"
    assert stft._fft_pack(CPU) is None
    stft.precision, stft.algorithm = "auto", "gemm"
    assert stft._fft_pack(CPU) is None
    assert STFT(512, 128, 512)._fft_pack(CPU) is None
    stft = STFT(1024, 256, 1024)
    with torch.no_grad():
        stft.forward_basis[3] *= 2.0
    stft.algorithm = "fft"
    try:
        stft._fft_pack(CPU)
        raise AssertionError("algorithm='fft' must insist")
    except RuntimeError:
        pass
