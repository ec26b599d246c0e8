"""Tacotron-2 Postnet (the non-autoregressive convs that feed the mel to the vocoder).

Golden output from the unmodified reference module (tests/golden/make_golden_postnet.py).  CPU: oracle vs golden,
state_dict layout, BatchNorm folding.  GPU: the tcgen05 implicit-GEMM path and the FP32 validation path."""
import os

import numpy as np
import pytest
import torch

import oracle
from tests import util
from text2speech_b200 import synthetic as syn

HP = syn.DEFAULT_POSTNET_HPARAMS
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "postnet_golden.npz")


@pytest.fixture(scope="module")
def golden_out():
    with np.load(GOLD) as f:
        return torch.from_numpy(f["postnet_out"])


def test_oracle_matches_reference(golden_out):
    sd = syn.synthetic_postnet_state_dict(HP, seed=77)
    with torch.no_grad():
        out = oracle.postnet(sd, syn.synthetic_mel(2, 37, seed=3))
    assert out.shape == golden_out.shape and util.rel_l2(out, golden_out) < 1e-6


def test_state_dict_layout_and_folding():
    from text2speech_b200.postnet import Postnet
    m = Postnet(HP)
    sd = syn.synthetic_postnet_state_dict(HP, seed=77)
    assert set(m.state_dict()) == set(sd) and all(m.state_dict()[k].shape == sd[k].shape for k in sd)
    m.load_state_dict(sd)
    m.eval()
    m.mode = "fp32"
    layers = m._packed(torch.device("cpu"))                       # packing is pure tensor work: runs without a GPU
    assert len(layers) == 5 and layers[0]["w"].shape == (5, 512, 80) and layers[4]["w"].shape == (5, 80, 512)
    # folded conv == conv + batch_norm(eval) for the first layer
    x = syn.synthetic_mel(1, 9, seed=1)
    w = layers[0]["w"].permute(1, 2, 0)[:, :80]                    # back to [c_out, c_in, taps]
    got = torch.nn.functional.conv1d(x, w, layers[0]["b"], padding=2)
    p = "convolutions.0."
    want = torch.nn.functional.batch_norm(torch.nn.functional.conv1d(x, sd[p + "0.conv.weight"], sd[p + "0.conv.bias"], padding=2),
                                          sd[p + "1.running_mean"], sd[p + "1.running_var"], sd[p + "1.weight"], sd[p + "1.bias"],
                                          training=False, eps=1e-5)
    assert util.rel_l2(got, want) < 1e-5
    with pytest.raises(RuntimeError):
        m.train()(x)                                               # inference-only
    with pytest.raises(RuntimeError):
        m.eval()(x)                                                # CPU tensor: no fallback


@pytest.mark.gpu
def test_gpu_postnet_matches_reference(golden_out):
    from text2speech_b200.postnet import Postnet
    m = Postnet(HP)
    m.load_state_dict(syn.synthetic_postnet_state_dict(HP, seed=77))
    m = m.cuda().eval()
    mel = syn.synthetic_mel(2, 37, seed=3).cuda()
    m.mode = "fp32"
    out32 = m(mel).cpu()
    assert out32.shape == golden_out.shape and util.rel_l2(out32, golden_out) <= 1e-5
    m.mode = "bf16"
    out16 = m(mel).cpu()
    assert util.snr_db(out16, golden_out) >= util.MIN_SNR_DB
    # longer, ragged length and odd batch: bf16 path against the fp32 path
    mel2 = syn.synthetic_mel(3, 333, seed=4).cuda()
    a = m(mel2)
    m.mode = "fp32"
    b = m(mel2)
    assert a.shape == (3, 80, 333) and util.snr_db(a.cpu(), b.cpu()) >= util.MIN_SNR_DB


@pytest.mark.gpu
def test_gpu_conv1d_taps_exact_integers():
    """wgb_tc_conv1d with small-integer operands: exact against an fp64 conv, dilation 1 and 3, relu / none."""
    from text2speech_b200 import _lib
    g = torch.Generator().manual_seed(9)
    B, T, C, N, taps = 2, 300, 128, 256, 5
    a = torch.randint(-3, 4, (B, T, C), generator=g).float()
    w = torch.randint(-2, 3, (N, C, taps), generator=g).float()
    bias = torch.randint(-5, 6, (N,), generator=g).float()
    for dil, act in ((1, 0), (3, 2)):
        want = torch.nn.functional.conv1d(a.permute(0, 2, 1).double(), w.double(), bias.double(), dilation=dil,
                                          padding=dil * (taps - 1) // 2).permute(0, 2, 1)
        if act == 2:
            want = want.clamp_min(0)
        out = torch.empty((B, T, N), device="cuda", dtype=torch.float32)
        wp = w.permute(0, 2, 1).reshape(N, taps * C).contiguous()
        _lib.call("wgb_tc_conv1d", a.cuda().bfloat16(), wp.cuda().bfloat16(), bias.cuda(), out, 0, B, T, N, C, taps, dil,
                  act, _lib.stream_ptr())
        torch.cuda.synchronize()
        assert torch.equal(out.cpu().double(), want)


# ---------------------------------------------------------------------------------------------------- Encoder conv bank
ENC_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encoder_golden.npz")


@pytest.fixture(scope="module")
def encoder_golden():
    with np.load(ENC_GOLD) as f:
        return torch.from_numpy(f["conv_bank_out"]).permute(0, 2, 1).contiguous()     # -> [B, 512, T]


def _encoder_input():
    from tests.golden.make_golden_encoder import encoder_input
    return encoder_input()


def test_encoder_convs_oracle_matches_reference(encoder_golden):
    """tacotron/tacotron.py:211-217 of the unmodified reference (captured at the LSTM's input) vs the oracle."""
    sd = syn.synthetic_encoder_convs_state_dict(seed=78)
    with torch.no_grad():
        out = oracle.encoder_convs(sd, _encoder_input())
    assert out.shape == encoder_golden.shape and util.rel_l2(out, encoder_golden) < 1e-6


def test_encoder_convs_state_dict_layout():
    from text2speech_b200.postnet import EncoderConvs
    m = EncoderConvs(syn.DEFAULT_ENCODER_HPARAMS)
    sd = syn.synthetic_encoder_convs_state_dict(seed=78)
    assert set(m.state_dict()) == set(sd) and all(m.state_dict()[k].shape == sd[k].shape for k in sd)
    m.load_state_dict(sd)
    assert m.acts == [2, 2, 2] and len(m.convolutions) == 3
    with pytest.raises(RuntimeError):
        m.train()(_encoder_input())                               # inference-only


@pytest.mark.gpu
def test_gpu_encoder_convs_match_reference(encoder_golden):
    from text2speech_b200.postnet import EncoderConvs
    m = EncoderConvs(syn.DEFAULT_ENCODER_HPARAMS)
    m.load_state_dict(syn.synthetic_encoder_convs_state_dict(seed=78))
    m = m.cuda().eval()
    x = _encoder_input().cuda()
    m.mode = "fp32"
    out32 = m(x).cpu()
    assert out32.shape == encoder_golden.shape and util.rel_l2(out32, encoder_golden) <= 1e-5
    m.mode = "bf16"
    out16 = m(x).cpu()
    assert util.snr_db(out16, encoder_golden) >= util.MIN_SNR_DB
    assert float(out16.min()) >= 0.0                              # relu in the GEMM epilogue
