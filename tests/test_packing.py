"""Host logic without a GPU: weight packing + the kernels' algorithmic restructuring, emulated in
torch from the PACKED weights, must reproduce the oracle (and hence the reference)."""
import pytest
import torch

import oracle
from tests import emulate, util
from text2speech_b200.packing import PackedWaveGlow, gate_row_order


@pytest.fixture(scope="module")
def packed():
    sd = util.state_dict("stress")
    return PackedWaveGlow(sd, 12, 8, 512, 8, "bf16", torch.device("cpu"))


def test_gate_row_order_pairs_tanh_with_sigmoid():
    order = gate_row_order(512)
    assert order.shape == (1024,) and sorted(order.tolist()) == list(range(1024))
    for p in range(4):
        blk = order[p * 256:(p + 1) * 256]
        assert torch.equal(blk[128:], blk[:128] + 512)


def test_upsample_gemm_matches_conv_transpose(packed):
    mel, _, _ = util.golden_inputs(2, 5)
    sd = oracle.folded_state(util.state_dict("stress"))
    up = oracle.upsample_spect(sd, mel)[:, :, : 5 * 256]
    want = oracle.regroup_spect(up, 8).permute(0, 2, 1)
    pk32 = PackedWaveGlow(util.state_dict("stress"), 12, 8, 512, 8, "fp32", torch.device("cpu"))
    got = emulate.upsample(pk32, mel)                       # fp32 layout (K = 4 x 80): exact repacking
    assert got.shape == want.shape == (2, 160, 640)
    assert util.rel_l2(got, want) < 1e-5
    got16 = emulate.upsample(packed, mel)                   # bf16 layout (K = 4 x 128, bf16 operands)
    assert got16.shape == want.shape
    assert util.rel_l2(got16, want) < 5e-3


def test_infer_emulation_fp32_matches_oracle(packed):
    mel, z, _ = util.golden_inputs(1, 3)
    with torch.no_grad():
        want = oracle.waveglow_infer(util.state_dict("stress"), mel, z, util.SIGMA)
        # bf16 weights, fp32 activations: isolates the packing from activation rounding
        got = emulate.infer(packed, mel, z, util.SIGMA, round_bf16=False)
    assert util.snr_db(got, want) > 40.0


def test_infer_emulation_bf16_predicts_snr(packed):
    mel, z, _ = util.golden_inputs(1, 3)
    with torch.no_grad():
        want = oracle.waveglow_infer(util.state_dict("stress"), mel, z, util.SIGMA)
        got = emulate.infer(packed, mel, z, util.SIGMA, round_bf16=True)
    assert util.snr_db(got, want) > util.MIN_SNR_DB


def test_forward_emulation_matches_oracle(packed):
    mel, _, wav = util.golden_inputs(1, 3)
    with torch.no_grad():
        zr, lsr, ldr = oracle.waveglow_forward(util.state_dict("stress"), mel, wav)
        z, ls, ld = emulate.forward(packed, mel, wav, round_bf16=False)
    assert util.snr_db(z, zr) > 40.0
    for k in (0, 5, 11):
        assert util.rel_l2(ls[k], lsr[k]) < 2e-2
        assert abs(ld[k] - float(ldr[k])) < 1e-3


def test_exact_packing_with_fp32_weights():
    """Same emulation with the weights left in fp32: must match the oracle to fp32 round-off,
    proving the permutations / folds are exact (not merely 'close in bf16')."""
    sd = util.state_dict("stress")
    pk = PackedWaveGlow(sd, 12, 8, 512, 8, "bf16", torch.device("cpu"))
    from text2speech_b200 import packing
    st = packing.folded(sd)
    for k, fl in enumerate(pk.flows):           # swap bf16 tensors for their fp32 originals
        p = f"WN.{k}."
        w_rs = [st[p + f"res_skip_layers.{i}.weight"] for i in range(8)]
        b_rs = [st[p + f"res_skip_layers.{i}.bias"] for i in range(8)]
        fl["w_skip"] = packing.pack_skip(w_rs, b_rs, 512)[0]
        for i in range(8):
            fl["w_gate"][i] = packing.pack_gate(st[p + f"in_layers.{i}.weight"], st[p + f"in_layers.{i}.bias"],
                                                st[p + f"cond_layers.{i}.weight"], st[p + f"cond_layers.{i}.bias"])[0]
            if i < 7:
                fl["w_res"][i] = w_rs[i][:512, :, 0]
    pk.w_up = packing.pack_upsample(st["upsample.weight"], st["upsample.bias"], 8, pk.up_ld_tap)[0]
    pk.mode = "exact"                           # emulate.upsample: do not round the im2col rows
    mel, z, _ = util.golden_inputs(1, 3)
    with torch.no_grad():
        want = oracle.waveglow_infer(sd, mel, z, util.SIGMA)
        got = emulate.infer(pk, mel, z, util.SIGMA, round_bf16=False)
    assert util.rel_l2(got, want) < 2e-5


def test_packed_mix_with_non_orthogonal_weights():
    """'skew' recipe (W^-1 != W^T, log det W != 0, det W < 0 in flows 3 and 8): the packed inverse / forward mixes and
    log_det_W through the kernels' data layout (emulated from the packed weights) against the oracle.  With the
    orthogonal recipes a packer that returned W^T for W^-1 would pass; here it cannot."""
    import math
    sd = util.state_dict("skew")
    pk = PackedWaveGlow(sd, 12, 8, 512, 8, "bf16", torch.device("cpu"))
    mel, z, wav = util.golden_inputs(1, 3)
    with torch.no_grad():
        want = oracle.waveglow_infer(sd, mel, z, util.SIGMA)
        got = emulate.infer(pk, mel, z, util.SIGMA, round_bf16=False)
        assert util.snr_db(got, want) > 40.0
        zr, lsr, ldr = oracle.waveglow_forward(sd, mel, wav)
        zf, ls, ld = emulate.forward(pk, mel, wav, round_bf16=False)
        assert util.snr_db(zf, zr) > 40.0
    for k in range(12):
        w, r = float(ldr[k]), ld[k]
        assert math.isnan(r) == math.isnan(w) == (k in (3, 8))
        if not math.isnan(w):
            assert abs(w) > 10.0 and abs(r - w) <= 1e-5 * abs(w)
    for fl in pk.flows:                          # transposing instead of inverting is now visibly wrong
        c = 2 * fl["n_half"]
        assert util.rel_l2(fl["w_mix_inv"][:c, :c], fl["w_mix"][:c, :c].t()) > 0.3


def test_cond_mel_composition_matches_cond_layer_of_upsampled_mel():
    """pack_cond_mel: cond_layers[i](regroup(upsample(mel))) == per-phase (W_cond U_phase) . stack(mel[f-j]) + bias
    (the algebra behind wgb_tc2_wn_gate_mel; glow.py:252-258 + :161)."""
    from text2speech_b200.packing import pack_cond_mel, pack_gate
    st = oracle.folded_state(util.state_dict("stress"))
    k, i, bsz, frames = 7, 3, 2, 5
    p = f"WN.{k}."
    wg, bg = pack_gate(st[p + f"in_layers.{i}.weight"], st[p + f"in_layers.{i}.bias"],
                       st[p + f"cond_layers.{i}.weight"], st[p + f"cond_layers.{i}.bias"])
    v, b_add = pack_cond_mel(wg[:, 1536:], st["upsample.weight"], st["upsample.bias"], 8)
    assert v.shape == (32, 1024, 320) and b_add.shape == (1024,)
    mel, _, _ = util.golden_inputs(bsz, frames)
    stack = torch.zeros(bsz, frames, 4, 80)
    for j in range(4):
        stack[:, j:, j] = mel[:, :, : frames - j].permute(0, 2, 1)
    stack = stack.reshape(bsz, frames, 320)
    got = torch.einsum("pok,bfk->bfpo", v, stack).reshape(bsz, frames * 32, 1024) + (bg + b_add)
    up = oracle.upsample_spect(st, mel)[:, :, : frames * 256]
    cond = oracle.regroup_spect(up, 8)
    want = torch.nn.functional.conv1d(cond, st[p + f"cond_layers.{i}.weight"], st[p + f"cond_layers.{i}.bias"])
    want = (want + st[p + f"in_layers.{i}.bias"][None, :, None])[:, gate_row_order(512)].permute(0, 2, 1)
    assert util.rel_l2(got, want) < 1e-5


def test_skip_end_composition_matches_skip_sum_then_end():
    """pack_skip_end16: end(sum_i skip_i) == (W_end W_skip) . acts + folded bias, hi + lo parts sum to the product."""
    from text2speech_b200.packing import pack_end, pack_skip, pack_skip_end16
    st = oracle.folded_state(util.state_dict("stress"))
    k, bsz, t = 6, 2, 40
    p = f"WN.{k}."
    w_rs = [st[p + f"res_skip_layers.{i}.weight"] for i in range(8)]
    b_rs = [st[p + f"res_skip_layers.{i}.bias"] for i in range(8)]
    w_skip, b_skip = pack_skip(w_rs, b_rs, 512)
    w_end, b_end = st[p + "end.weight"], st[p + "end.bias"]
    _, _, b_fold = pack_end(w_end, b_end, b_skip)
    w16 = pack_skip_end16(w_skip, w_end)
    assert w16.shape == (16, 4096) and w16.dtype == torch.bfloat16
    comp = (w16[:8].double() + w16[8:].double())
    rows = w_end.shape[0]
    exact = w_end[:, :, 0].double() @ w_skip.double()
    assert float((comp[:rows] - exact).abs().max() / exact.abs().max()) < 2e-5 and float(comp[rows:].abs().max()) == 0.0
    g = torch.Generator().manual_seed(3)
    acts = torch.rand(8, bsz, 512, t, generator=g) * 2 - 1
    total = 0
    for i in range(8):
        lo = 0 if i == 7 else 512
        total = total + torch.nn.functional.conv1d(acts[i], w_rs[i][lo: lo + 512], b_rs[i][lo: lo + 512])
    want = torch.nn.functional.conv1d(total, w_end, b_end)                                   # [B, 2n_half, T]
    a_cat = acts.permute(1, 3, 0, 2).reshape(bsz, t, 4096).double()                          # K index = layer*512 + c
    got = (a_cat @ comp.t())[:, :, :rows] + b_fold[:rows].double()
    assert util.rel_l2(got.permute(0, 2, 1), want) < 1e-5


def _x_stack_rows(x, n_half):
    """Host restatement of wgb_x_stack's row layout (fp32 values of the bf16 entries)."""
    bsz, t, _ = x.shape
    base = 8 - 2 * n_half
    a0 = torch.zeros(bsz, t, 4)
    a0[:, :, :n_half] = x[:, :, base: base + n_half]
    taps = torch.zeros(bsz, t, 3, 4)
    ind = torch.zeros(bsz, t, 3)
    for tap in range(3):
        s = tap - 1
        lo, hi = max(0, -s), min(t, t - s)
        taps[:, lo:hi, tap] = a0[:, lo + s: hi + s]
        ind[:, lo:hi, tap] = 1.0
    flat = taps.reshape(bsz, t, 12)
    a_hi = flat.bfloat16().float()
    a_lo = (flat - a_hi).bfloat16().float()
    rows = torch.zeros(bsz, t, 64)
    rows[:, :, 0:12], rows[:, :, 12:24], rows[:, :, 24:36] = a_hi, a_lo, a_hi
    rows[:, :, 36:39], rows[:, :, 39:42] = ind, ind
    return rows


def test_start_folded_into_first_in_layer():
    """pack_gate0 + the x_stack layout: in_layers[0](start(a0)) == W0 . x_stack row (edges included), packed order."""
    from text2speech_b200.packing import pack_gate0
    st = oracle.folded_state(util.state_dict("stress"))
    for k in (10, 6, 1):                                           # n_half 2, 3, 4
        p = f"WN.{k}."
        w_in0, b_in0 = st[p + "in_layers.0.weight"], st[p + "in_layers.0.bias"]
        w_start, b_start = st[p + "start.weight"][:, :, 0], st[p + "start.bias"]
        n_half = w_start.shape[1]
        g = torch.Generator().manual_seed(k)
        x = torch.randn(2, 37, 8, generator=g)
        base = 8 - 2 * n_half
        a0 = x[:, :, base: base + n_half].permute(0, 2, 1)                                   # [B, n_half, T]
        h0 = torch.nn.functional.conv1d(a0, w_start[:, :, None], b_start)
        want = torch.nn.functional.conv1d(h0, w_in0, None, padding=1)[:, gate_row_order(512)].permute(0, 2, 1)
        w0 = pack_gate0(w_in0, w_start, b_start)
        assert w0.shape == (1024, 64) and w0.dtype == torch.bfloat16
        got = _x_stack_rows(x, n_half).double() @ w0.double().t()
        assert util.rel_l2(got, want) < 3e-5, k


@pytest.mark.parametrize("cond_path", ["mel", "cond"])
def test_engine_host_sequence_on_emulated_kernels(monkeypatch, cond_path):
    """The REAL engine.infer / engine.forward (kernel sequencing, padded frame layout and guard rows, composed
    conditioning operands, first-layer fold, fused next-flow start / 1x1 mix) run on the CPU with every entry point
    replaced by a torch stand-in of its contract (tests/emulate_engine.py), against the oracle."""
    from tests import emulate_engine
    from text2speech_b200 import engine
    calls = emulate_engine.install(monkeypatch)
    from text2speech_b200 import synthetic as syn
    cfg = dict(syn.load_config())
    cfg.update(n_flows=4, n_early_every=2, n_early_size=2)        # n_half 4, 4, 3, 3: keeps the CPU composition cheap
    sd = syn.synthetic_state_dict(cfg, seed=1234, end_std=0.01)
    pk = PackedWaveGlow(sd, 4, 8, 512, 8, "bf16", torch.device("cpu"), compose_cond=(cond_path == "mel"))
    pk.cond_path = cond_path
    mel, z, wav = util.golden_inputs(2, 6)
    with torch.no_grad():
        audio = engine.infer(pk, mel, z, util.SIGMA)
        want = oracle.waveglow_infer(sd, mel, z, util.SIGMA)
        assert audio.shape == want.shape and util.snr_db(audio, want) >= util.MIN_SNR_DB
        if cond_path == "mel":
            assert "wgb_tc2_wn_gate_mel0" in calls and "wgb_x_stack" in calls and "wgb_tc2_wn_gate" not in calls
            assert calls.count("wgb_wn_start_padded") == 1                   # later flows: fused into the skip+end kernel
        else:
            assert "wgb_tc2_wn_gate_mel" not in calls and calls.count("wgb_tc2_wn_gate") == 32
        # forward direction, audio shorter than 256 * frames (trimmed / partial last frame)
        zf, log_s, log_det = engine.forward(pk, mel, wav[:, :-72])
        zw, lsw, ldw = oracle.waveglow_forward(sd, mel, wav[:, :-72])
        assert util.snr_db(zf, zw) >= util.MIN_SNR_DB
        assert util.snr_db(log_s[-1], lsw[-1]) >= 25.0
        assert all(abs(float(a) - float(b)) <= 1e-3 for a, b in zip(log_det, ldw))
