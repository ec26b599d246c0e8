"""Pin the CPU oracle against outputs of the UNMODIFIED reference (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

import oracle
from oracle.waveglow_oracle import algorithmic_flop_per_group_step, flow_channels
from tests import util
from text2speech_b200 import synthetic as syn

DC = syn.DEFAULT_DATA_CONFIG


@pytest.mark.parametrize("recipe", ["bench", "stress", "skew"])
def test_infer_matches_reference(golden, recipe):
    mel, z, _ = util.golden_inputs()
    taps = {}
    with torch.no_grad():
        audio = oracle.waveglow_infer(util.state_dict(recipe), mel, z, util.SIGMA, taps)
    assert audio.shape == (2, 6 * 256)
    assert util.rel_l2(audio, golden[f"{recipe}_infer_audio"]) < 1e-5
    steps = golden[f"{recipe}_tap_steps"]
    sub = taps[11]
    for i in (0, 3, 7):
        assert util.rel_l2(sub[f"h{i}"][0][:, steps], golden[f"{recipe}_wn11_h{i}"]) < 1e-5
        assert util.rel_l2(sub[f"acts{i}"][0][:, steps], golden[f"{recipe}_wn11_acts{i}"]) < 1e-5
    for k in (11, 5, 0):
        assert util.rel_l2(taps[k]["wn_out"], golden[f"{recipe}_wn{k}_out"]) < 1e-5


@pytest.mark.parametrize("recipe", ["bench", "stress", "skew"])
def test_forward_matches_reference(golden, recipe):
    mel, _, wav = util.golden_inputs()
    with torch.no_grad():
        z, log_s, log_det = oracle.waveglow_forward(util.state_dict(recipe), mel, wav)
        zs, _, _ = oracle.waveglow_forward(util.state_dict(recipe), mel, wav[:, :-64])
    assert util.rel_l2(z, golden[f"{recipe}_fwd_z"]) < 1e-6
    assert util.rel_l2(zs, golden[f"{recipe}_fwd_z_short"]) < 1e-6
    for k in (0, 5, 11):
        assert util.rel_l2(log_s[k], golden[f"{recipe}_fwd_log_s{k}"]) < 1e-6
    got = np.array([float(v) for v in log_det])
    want = golden[f"{recipe}_fwd_log_det"]
    # orthogonal recipes: |want| ~ 1e-4 (absolute); 'skew': hundreds (relative), NaN where det W < 0 (torch.logdet)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.all(np.abs(got[ok] - want[ok]) <= 1e-6 + 1e-6 * np.abs(want[ok]))
    if recipe == "skew":
        assert int(np.isnan(want).sum()) == 2 and np.nanmin(np.abs(want)) > 50.0


def test_weight_norm_fold_matches_reference(golden):
    mel, z, _ = util.golden_inputs()
    sd = util.state_dict("bench", weight_norm=True)
    with torch.no_grad():
        audio = oracle.waveglow_infer(sd, mel[:1, :, :4], z[:1, :, :128], util.SIGMA)
    assert util.rel_l2(audio, golden["bench_wnorm_infer_audio"]) < 1e-5


def test_forward_then_infer_is_identity():
    """Free self-check of the flow: infer(z = forward(x)) == x at sigma = 1."""
    mel, _, wav = util.golden_inputs(1, 4)
    sd = util.state_dict("stress")
    with torch.no_grad():
        z, _, _ = oracle.waveglow_forward(sd, mel, wav)
        back = oracle.waveglow_infer(sd, mel, z, 1.0)
    assert util.rel_l2(back, wav) < 1e-4


def test_flop_count_and_channels():
    assert flow_channels(12, 8, 4, 2) == [8, 8, 8, 8, 6, 6, 6, 6, 4, 4, 4, 4]
    assert algorithmic_flop_per_group_step(util.state_dict("bench")) == 522302368


def test_stft_bases_and_transform(golden):
    fwd, inv = oracle.stft_bases(DC["filter_length"], DC["hop_length"], DC["win_length"])
    rows = golden["stft_rows"]
    assert np.allclose(fwd[rows, 0].numpy(), golden["stft_forward_basis_rows"], atol=1e-7)
    assert np.allclose(inv[rows, 0].numpy(), golden["stft_inverse_basis_rows"], atol=1e-9)
    y = syn.synthetic_waveforms(2, 4096, sr=DC["sampling_rate"], seed=5)
    mag, phase = oracle.stft_transform(y, fwd, DC["hop_length"])
    assert util.rel_l2(mag, golden["stft_mag"]) < 1e-6
    rec = oracle.stft_inverse(mag, phase, inv, DC["hop_length"], DC["win_length"])
    assert rec.shape == (2, 1, 4096)
    assert util.rel_l2(rec, golden["stft_recon"]) < 1e-6
    assert np.array_equal(oracle.window_sumsquare("hann", 17, 256, 1024, 1024), golden["wss_17"])
    # independent pin: torch.stft agrees with the conv-basis transform
    ref = torch.stft(y, 1024, 256, 1024, torch.hann_window(1024), center=True, pad_mode="reflect",
                     return_complex=True).abs()
    assert float((ref - mag).abs().max()) < 1e-3


def test_small_stft_window_shorter_than_filter(golden):
    fwd, inv = oracle.stft_bases(64, 16, 48)
    y = syn.synthetic_waveforms(2, 4096, sr=DC["sampling_rate"], seed=5)[:, :512]
    mag, phase = oracle.stft_transform(y, fwd, 16)
    assert util.rel_l2(mag, golden["small_mag"]) < 1e-6
    rec = oracle.stft_inverse(mag, phase, inv, 16, 48)
    assert util.rel_l2(rec, golden["small_recon"]) < 1e-5


def test_stft_without_window(golden):
    """window=None (stft.py:56-62, :111-125): rectangular bases, no window-sum division and no L/hop scale."""
    fwd, inv = oracle.stft_bases(64, 16, 64, window=None)
    y = syn.synthetic_waveforms(2, 4096, sr=DC["sampling_rate"], seed=5)[:, :512]
    mag, phase = oracle.stft_transform(y, fwd, 16)
    assert util.rel_l2(mag, golden["nowin_mag"]) < 1e-6
    rec = oracle.stft_inverse(mag, phase, inv, 16, 64, window=None)
    assert util.rel_l2(rec, golden["nowin_recon"]) < 1e-5


def test_mel_filterbank_and_mel(golden):
    mb = oracle.mel_filterbank(DC["sampling_rate"], 1024, 80, DC["mel_fmin"], DC["mel_fmax"])
    # parity unpinned (librosa 0.6 absent): cross-check with torchaudio's independent Slaney bank
    assert np.allclose(mb.astype(np.float32), golden["mel_basis_22050"], atol=2e-7)
    assert (mb[:, 372:] == 0).all()          # bins above fmax carry no weight
    fwd, _ = oracle.stft_bases(1024, 256, 1024)
    y = syn.synthetic_waveforms(2, 4096, sr=DC["sampling_rate"], seed=5)
    mel = oracle.mel_spectrogram(y, fwd, torch.from_numpy(mb).float(), 256)
    assert mel.shape == (2, 80, 17)
    assert float((mel - torch.from_numpy(golden["mel"])).abs().max()) < 1e-4


def test_denoiser(golden):
    fwd, inv = oracle.stft_bases(1024, 256, 1024)
    sd = util.state_dict("bench")
    with torch.no_grad():
        bias_audio = oracle.waveglow_infer(sd, torch.zeros(1, 80, 88), torch.zeros(1, 8, 88 * 32), 0.0)
    bias = oracle.denoiser_bias_spec(bias_audio, fwd, 256)
    assert bias.shape == (1, 513, 1)
    assert util.rel_l2(bias, golden["denoiser_bias_spec"]) < 1e-4
    y = syn.synthetic_waveforms(2, 4096, sr=DC["sampling_rate"], seed=5)
    for s, key in ((0.1, "denoised_s0p1"), (0.01, "denoised_s0p01")):
        out = oracle.denoise(y, torch.from_numpy(golden["denoiser_bias_spec"]), s, fwd, inv, 256, 1024)
        assert out.shape == (2, 1, 4096)
        assert util.rel_l2(out, golden[key]) < 1e-5


def test_oracle_full_utterance_matches_reference():
    """The oracle at BASELINE configs[1] size (80x860 mel, T = 27 520) against the unmodified reference's audio
    (tests/golden/make_golden_full.py)."""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "full_utterance_golden.npz")
    with np.load(path) as f:
        want = f["bench_full_infer_audio"]
    torch.set_num_threads(os.cpu_count() or 1)
    mel, z = syn.synthetic_mel(1, 860, seed=0), syn.synthetic_z(1, 860, seed=2024)
    with torch.no_grad():
        got = oracle.waveglow_infer(util.state_dict("bench"), mel, z, util.SIGMA)
    assert got.shape == (1, 220160)
    assert util.rel_l2(got, want) < 1e-5
