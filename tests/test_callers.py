"""Caller-side rows of the hot path: vocoder CLI (waveglow/inference.py), checkpoint formats
(pickled glow.WaveGlow modules, convert_model.py), Griffin-Lim (utils/audio_processing.py:51-67).

Golden files come from the unmodified reference (tests/golden/make_golden_cli.py).  CPU tests pin the
oracle and the host logic (checkpoint loading / migration needs no GPU); `gpu` tests run the product."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle
from tests import util
from text2speech_b200 import synthetic as syn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DC = syn.DEFAULT_DATA_CONFIG
SIGMA = 0.666
GL_SEED, GL_ITERS = 7, 4


@pytest.fixture(scope="module")
def cli_golden():
    with np.load(os.path.join(GOLD, "cli_golden.npz")) as f:
        return {k: f[k] for k in f.files}


def tiny_inputs():
    g = torch.Generator().manual_seed(11)
    mel = torch.randn((2, 8, 5), generator=g)
    z = torch.randn((2, 8, 5 * 32), generator=g)
    return mel, z


# ------------------------------------------------------------------------------------ CPU: oracle + host logic

def test_oracle_griffin_lim_matches_reference(cli_golden):
    fwd, inv = oracle.stft_bases(DC["filter_length"], DC["hop_length"], DC["win_length"])
    y = syn.synthetic_waveforms(2, 4096, sr=DC["sampling_rate"], seed=5)
    mag, _ = oracle.stft_transform(y, fwd, DC["hop_length"])
    angles = oracle.griffin_lim_initial_angles(tuple(mag.shape), GL_SEED)
    sig = oracle.griffin_lim(mag, angles, fwd, inv, DC["hop_length"], DC["win_length"], n_iters=GL_ITERS)
    assert sig.shape == (2, 4096)
    assert util.rel_l2(sig, cli_golden["gl_signal"]) < 1e-4


def test_oracle_cli_samples_match_reference(cli_golden):
    """infer -> Denoiser(0.1) -> *32768 -> int16 of the oracle against the reference's CLI arithmetic."""
    sd = util.state_dict("bench")
    fwd, inv = oracle.stft_bases(1024, 256, 1024)
    with torch.no_grad():
        bias_audio = oracle.waveglow_infer(sd, torch.zeros(1, 80, 88), torch.zeros(1, 8, 88 * 32), 0.0)
        bias = oracle.denoiser_bias_spec(bias_audio, fwd, 256)
        for idx, frames in enumerate((5, 6, 5)):
            mel, z = syn.synthetic_mel(1, frames, seed=40 + idx), syn.synthetic_z(1, frames, seed=50 + idx)
            audio = oracle.waveglow_infer(sd, mel, z, SIGMA)
            pcm, ok = oracle.pcm16(audio)
            gold = cli_golden[f"cli_pcm_{idx}"]
            assert ok.mean() > 0.5
            assert np.abs(pcm[ok].astype(np.int32) - gold[ok].astype(np.int32)).max() <= 1
            den = oracle.denoise(audio, bias, 0.1, fwd, inv, 256, 1024)
            pcm_d, ok_d = oracle.pcm16(den)
            gold_f = cli_golden[f"cli_float_denoised_{idx}"]
            assert util.rel_l2(den.squeeze() * 32768.0, gold_f) < 1e-4
            close = ok_d & (np.abs(gold_f - np.round(gold_f)) > 0.05)      # away from truncation boundaries
            assert np.abs(pcm_d[close].astype(np.int32) - cli_golden[f"cli_pcm_denoised_{idx}"][close].astype(np.int32)).max() <= 1


@pytest.mark.parametrize("name", ["tiny_ckpt_ref.pt", "tiny_ckpt_ref_old.pt"])
def test_reference_pickles_load_without_the_reference(name, cli_golden):
    """A checkpoint pickled by the reference (class path glow.WaveGlow, train.py:52-60) loads into the drop-in
    classes with no reference code importable; old res/skip layouts are migrated (convert_model.py:11-38);
    the loaded network equals the reference's (CPU oracle on its state_dict vs the reference's own infer)."""
    assert not any("reference" in p for p in sys.path)
    from text2speech_b200 import glow as drop_in
    from text2speech_b200.inference import load_waveglow
    sys.modules.pop("glow", None)
    model = load_waveglow(os.path.join(GOLD, name))
    assert isinstance(model, drop_in.WaveGlow) and isinstance(model.WN[0], drop_in.WN)
    assert not hasattr(model.WN[0], "res_layers") and len(model.WN[0].res_skip_layers) == 2
    keys = set(model.state_dict())
    assert "WN.0.res_skip_layers.0.weight_g" in keys and "WN.3.end.weight" in keys
    model = drop_in.WaveGlow.remove_weightnorm(model)
    assert "WN.0.res_skip_layers.0.weight" in model.state_dict()
    mel, z = tiny_inputs()
    with torch.no_grad():
        got = oracle.waveglow_infer(model.state_dict(), mel, z, SIGMA)
    assert util.rel_l2(got, cli_golden["tiny_audio" if "old" not in name else "tiny_old_audio"]) < 1e-5


def test_update_model_passes_current_layout_through():
    from text2speech_b200 import glow as drop_in
    from text2speech_b200.convert_model import update_model
    m = drop_in.WaveGlow(8, 2, 8, 2, 2, {"n_layers": 2, "n_channels": 16, "kernel_size": 3})
    assert update_model(m) is m


def test_state_dict_checkpoint_layouts(tmp_path):
    from text2speech_b200.inference import load_waveglow
    sd = util.state_dict("bench")
    for i, payload in enumerate(({"model": dict(sd)}, {"state_dict": dict(sd)}, sd)):
        path = tmp_path / f"ck{i}.pt"
        torch.save(payload, path)
        model = load_waveglow(str(path))
        assert torch.equal(model.state_dict()["WN.3.in_layers.2.weight"], sd["WN.3.in_layers.2.weight"])
    sd_wn = util.state_dict("bench", weight_norm=True)
    torch.save({"model": dict(sd_wn)}, tmp_path / "wn.pt")
    model = load_waveglow(str(tmp_path / "wn.pt"))
    assert "WN.0.start.weight_g" in model.state_dict()


def test_files_to_list_and_constants(tmp_path):
    from text2speech_b200 import mel2samp
    p = tmp_path / "files.txt"
    p.write_text("a.pt\nb c.pt \n")
    assert mel2samp.files_to_list(str(p)) == ["a.pt", "b c.pt"]
    assert mel2samp.MAX_WAV_VALUE == 32768.0


# ------------------------------------------------------------------------------------ GPU: product

@pytest.mark.gpu
@pytest.mark.parametrize("name", ["tiny_ckpt_ref.pt", "tiny_ckpt_ref_old.pt"])
def test_gpu_infer_from_reference_pickle(name, cli_golden):
    from text2speech_b200 import glow as drop_in
    from text2speech_b200.inference import load_waveglow
    model = drop_in.WaveGlow.remove_weightnorm(load_waveglow(os.path.join(GOLD, name))).cuda().eval()
    model.mode = "fp32"                     # 8 mels / 16 channels: outside the tcgen05 specialisation
    mel, z = tiny_inputs()
    got = model.infer(mel.cuda(), sigma=SIGMA, z=z.cuda()).cpu()
    assert util.rel_l2(got, cli_golden["tiny_audio"]) < 3e-5
    model.mode = "bf16"
    with pytest.raises(RuntimeError, match="specialised"):
        model.infer(mel.cuda(), sigma=SIGMA, z=z.cuda())


@pytest.mark.gpu
def test_gpu_cli_end_to_end(tmp_path, cli_golden):
    """python -m text2speech_b200.inference semantics: mel .pt files in, <name>_synthesis.wav int16 out."""
    from scipy.io.wavfile import read
    from text2speech_b200 import inference as cli
    sd = util.state_dict("bench")
    torch.save({"model": dict(sd)}, tmp_path / "waveglow.pt")
    names, mels, zs = [], [], []
    for idx, frames in enumerate((5, 6, 5)):
        mel = syn.synthetic_mel(1, frames, seed=40 + idx)[0]
        torch.save(mel, tmp_path / f"utt{idx}.pt")
        names.append(str(tmp_path / f"utt{idx}.pt"))
        mels.append(mel)
        zs.append(syn.synthetic_z(1, frames, seed=50 + idx)[0])
    (tmp_path / "mels.txt").write_text("\n".join(names) + "\n")
    # file-level run (random noise: checks plumbing, naming, dtype, length)
    paths = cli.main(str(tmp_path / "mels.txt"), str(tmp_path / "waveglow.pt"), SIGMA, str(tmp_path / "out"), 22050,
                     False, 0.1, batch=2)
    assert [os.path.basename(p) for p in paths] == [f"utt{i}_synthesis.wav" for i in range(3)]
    for p, frames in zip(paths, (5, 6, 5)):
        rate, data = read(p)
        assert rate == 22050 and data.dtype == np.int16 and data.shape == (frames * 256,)
    # sample-level parity with host-supplied noise, FP32 validation mode, batched (2 equal-length mels share a call)
    model = cli.load_waveglow(str(tmp_path / "waveglow.pt")).cuda().eval()
    model.mode = "fp32"
    den = cli.Denoiser(model).cuda()
    plain = cli.synthesize(model, mels, SIGMA, batch=2, z=zs)
    deno = cli.synthesize(model, mels, SIGMA, den, 0.1, batch=2, z=zs)
    for idx in range(3):
        gold_f = cli_golden[f"cli_float_denoised_{idx}"]
        gold = cli_golden[f"cli_pcm_{idx}"].astype(np.int32)
        got = plain[idx].numpy().astype(np.int32)
        ok = (np.abs(got) < 32767)
        assert ok.mean() > 0.5 and np.abs(got[ok] - gold[ok]).max() <= 1
        got_d = deno[idx].numpy().astype(np.int32)
        ok_d = (np.abs(gold_f) < 32766) & (np.abs(gold_f - np.round(gold_f)) > 0.2)
        assert np.abs(got_d[ok_d] - cli_golden[f"cli_pcm_denoised_{idx}"].astype(np.int32)[ok_d]).max() <= 1
        sat = gold_f > 32768
        assert (got_d[sat] == 32767).all()          # documented deviation: saturate instead of wrapping
    # BF16 mode end to end (infer -> denoise) stays within the north-star SNR
    model.mode = "bf16"
    deno16 = cli.synthesize(model, mels, SIGMA, den, 0.1, batch=1, z=zs)
    for idx in range(3):
        gold_f = cli_golden[f"cli_float_denoised_{idx}"]
        ok = np.abs(gold_f) < 32000
        assert util.snr_db(deno16[idx].numpy()[ok].astype(np.float64), gold_f[ok]) > util.MIN_SNR_DB


@pytest.mark.gpu
def test_gpu_griffin_lim(cli_golden):
    import text2speech_b200 as t2s
    from text2speech_b200.audio_processing import griffin_lim
    stft = t2s.STFT(DC["filter_length"], DC["hop_length"], DC["win_length"]).cuda()
    y = syn.synthetic_waveforms(2, 4096, sr=DC["sampling_rate"], seed=5).cuda()
    mag, _ = stft.transform(y)
    angles = oracle.griffin_lim_initial_angles(tuple(mag.shape), GL_SEED)
    sig = griffin_lim(mag, stft, n_iters=GL_ITERS, angles=angles)
    assert sig.shape == (2, 4096)
    assert util.rel_l2(sig.cpu(), cli_golden["gl_signal"]) < 1e-3
    # more iterations never increase the spectral inconsistency much (projection property)
    def inconsistency(s):
        m2, _ = stft.transform(s)
        return float((m2 - mag).norm() / mag.norm())
    assert inconsistency(griffin_lim(mag, stft, n_iters=12, angles=angles)) <= inconsistency(sig) * 1.05


@pytest.mark.gpu
def test_gpu_audio_to_int16():
    from text2speech_b200.inference import audio_to_int16
    x = torch.tensor([0.0, 0.5, -0.5, 0.99999, -1.0, 1.5, -1.5, 1e-5, -1e-5, 0.25001, 3.0517578125e-05 * 7.9], device="cuda")
    got = audio_to_int16(x).cpu().numpy()
    want = np.trunc(np.clip(x.cpu().numpy().astype(np.float32) * np.float32(32768.0), -32768, 32767)).astype(np.int16)
    assert np.array_equal(got, want)
    big = torch.randn(100003, device="cuda") * 0.3
    got = audio_to_int16(big).cpu().numpy()
    want = np.trunc(np.clip(big.cpu().numpy() * np.float32(32768.0), -32768, 32767)).astype(np.int16)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_mel2samp_batch(tmp_path):
    from scipy.io.wavfile import write
    from text2speech_b200.mel2samp import Mel2Samp
    y = syn.synthetic_waveforms(3, 16000, sr=22050, seed=9)
    names = []
    for i in range(3):
        write(tmp_path / f"w{i}.wav", 22050, (y[i] * 32767).numpy().astype(np.int16))
        names.append(str(tmp_path / f"w{i}.wav"))
    (tmp_path / "train.txt").write_text("\n".join(names) + "\n")
    ds = Mel2Samp(str(tmp_path / "train.txt"), **{k: DC[k] for k in ("segment_length", "filter_length", "hop_length",
                                                                    "win_length", "sampling_rate", "mel_fmin", "mel_fmax")})
    mel, audio = ds[0]
    assert mel.shape == (80, 16000 // 256 + 1) and audio.shape == (16000,) and float(audio.abs().max()) <= 1.0
    fwd, _ = oracle.stft_bases(1024, 256, 1024)
    mb = torch.from_numpy(oracle.mel_filterbank(22050, 1024, 80, 0.0, 8000.0)).float()
    want = oracle.mel_spectrogram(audio[None], fwd, mb, 256)[0]
    assert float((mel.cpu() - want).abs().max()) < 2e-3
    batch = ds.mel_batch(torch.stack([audio, audio]) * 32768.0)
    assert torch.allclose(batch[0], mel, atol=1e-5) and torch.equal(batch[0], batch[1])


@pytest.mark.gpu
def test_gpu_mel2samp_cli_writes_reference_format(tmp_path):
    """python -m text2speech_b200.mel2samp (mel2samp.py:110-142): wav list -> <name>.pt mel files, batched by length."""
    import json
    from scipy.io.wavfile import write
    from text2speech_b200 import mel2samp
    y = syn.synthetic_waveforms(3, 9000, sr=22050, seed=12)
    names = []
    for i, n in enumerate((9000, 7000, 9000)):
        write(tmp_path / f"a{i}.wav", 22050, (y[i, :n] * 32767).numpy().astype(np.int16))
        names.append(str(tmp_path / f"a{i}.wav"))
    (tmp_path / "files.txt").write_text("\n".join(names) + "\n")
    cfg = {"data_config": dict(training_files=str(tmp_path / "files.txt"), **DC)}
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    written = mel2samp.main(str(tmp_path / "files.txt"), str(tmp_path / "config.json"), str(tmp_path / "mels"), batch=8)
    assert sorted(os.path.basename(p) for p in written) == ["a0.wav.pt", "a1.wav.pt", "a2.wav.pt"]
    fwd, _ = oracle.stft_bases(1024, 256, 1024)
    mb = torch.from_numpy(oracle.mel_filterbank(22050, 1024, 80, 0.0, 8000.0)).float()
    for i, n in enumerate((9000, 7000, 9000)):
        mel = torch.load(tmp_path / "mels" / f"a{i}.wav.pt")
        audio = torch.from_numpy((y[i, :n] * 32767).numpy().astype(np.int16)).float() / 32768.0
        want = oracle.mel_spectrogram(audio[None], fwd, mb, 256)[0]
        assert mel.shape == (80, n // 256 + 1) and float((mel - want).abs().max()) < 2e-3
