"""Training direction: gradients of WaveGlowLoss(WaveGlow.forward) w.r.t. every reference
parameter, pinned to the UNMODIFIED reference's autograd (tests/golden/train_grads_golden.npz, written by
tests/golden/make_golden_grads.py), plus the kernels behind them, the fused Adam step and the flat-bucket all-reduce.
"""
from __future__ import annotations

import os
import warnings

import numpy as np
import pytest
import torch

from tests import util
from tests.golden.make_golden_grads import SIGMA, sample_index, train_config, train_inputs
from text2speech_b200 import synthetic as syn

HERE = os.path.dirname(os.path.abspath(__file__))
DEV = "cuda:0"


@pytest.fixture(scope="module")
def golden_grads():
    with np.load(os.path.join(HERE, "golden", "train_grads_golden.npz")) as f:
        return {k: f[k] for k in f.files}


def _train_state():
    return syn.synthetic_state_dict(train_config(), seed=777, end_std=0.05, weight_norm=True)


def _compare(named_grads, golden, tol_big, tol_small, min_fraction=1.0):
    """Sampled entries of every gradient vs the reference's.  Returns the worst relative error per parameter kind."""
    worst = {}
    bad = []
    for name, g in named_grads:
        g = g.detach().float().cpu().flatten()
        ref = torch.from_numpy(golden[name + "|samples"])
        got = g[torch.from_numpy(sample_index(name, g.numel()))]
        scale = float(golden[name + "|norm"]) / max(g.numel(), 1) ** 0.5        # rms of the whole gradient
        err = float((got - ref).norm() / max(float(ref.norm()), 1e-3 * scale * len(ref) ** 0.5, 1e-30))
        kind = name.split(".")[-2] + "." + name.split(".")[-1] if name.startswith("WN") else name
        worst[kind] = max(worst.get(kind, 0.0), err)
        tol = tol_small if g.numel() <= 4096 else tol_big
        if err > tol:
            bad.append((name, err))
    return worst, bad


def test_oracle_gradients_match_reference(golden_grads):
    """The CPU oracle under torch autograd reproduces the reference's parameter gradients (pins the checker)."""
    from oracle import waveglow_oracle as wo
    sd = {k: v.clone().requires_grad_(True) for k, v in _train_state().items()}
    mel, wav = train_inputs()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        z, log_s, log_det = wo.waveglow_forward(sd, mel, wav)
        loss = wo.waveglow_loss(z, log_s, log_det, SIGMA)
        loss.backward()
    assert abs(float(loss) - float(golden_grads["loss"])) <= 1e-5 * abs(float(golden_grads["loss"])) + 1e-6
    worst, bad = _compare([(k, v.grad) for k, v in sd.items()], golden_grads, 2e-4, 2e-4)
    assert not bad, f"oracle gradients differ from the reference's: {bad[:5]} (worst per kind: {worst})"


def test_allreduce_gradients_is_identity_without_process_group():
    from text2speech_b200 import training

    class Fake:
        def gather_grads(self):
            return torch.ones(8)

    assert training.allreduce_gradients(Fake()) == 1.0


def test_per_step_packing_matches_inference_packing():
    """training._pack_flow (device-side, every step) builds the same gate / residual / composed-skip operands as
    packing.PackedWaveGlow (host-side, once) and the mirrored-tap transposes the data gradients need."""
    from text2speech_b200 import packing, training
    cfg = train_config()
    sd = packing.folded(syn.synthetic_state_dict(cfg, seed=5, end_std=0.05, weight_norm=True))
    k = 2                                                          # n_half = 3
    f = training._pack_flow(sd, k)
    p = f"WN.{k}."
    for i in (0, 7):
        wg, bg = packing.pack_gate(sd[p + f"in_layers.{i}.weight"], sd[p + f"in_layers.{i}.bias"],
                                   sd[p + f"cond_layers.{i}.weight"], sd[p + f"cond_layers.{i}.bias"])
        assert torch.equal(f["w_gate"][i], wg.bfloat16()) and torch.equal(f["b_gate"][i], bg)
        w_in = sd[p + f"in_layers.{i}.weight"].bfloat16()
        wt = f["wt_in"][i].reshape(512, 3, 1024)                   # [c_in][tap'][c_out], tap' = 2 - tap
        assert torch.equal(wt[:, 0], w_in[:, :, 2].t()) and torch.equal(wt[:, 2], w_in[:, :, 0].t())
        w_cond = sd[p + f"cond_layers.{i}.weight"][:, :, 0].bfloat16()
        assert torch.equal(f["wt_cond"][:640, i * 1024:(i + 1) * 1024], w_cond.t())
    assert float(f["wt_cond"][640:].abs().max()) == 0.0
    w_rs = [sd[p + f"res_skip_layers.{i}.weight"] for i in range(8)]
    b_rs = [sd[p + f"res_skip_layers.{i}.bias"] for i in range(8)]
    w_skip, b_skip = packing.pack_skip(w_rs, b_rs, 512)
    w16 = packing.pack_skip_end16(w_skip, sd[p + "end.weight"])
    got16 = f["w_skip16"].float()
    assert util.rel_l2(got16[:8] + got16[8:], w16.float()[:8] + w16.float()[8:]) <= 5e-6   # hi + lo = the composed product (fp32 here, fp64 there)
    _, _, b_fold = packing.pack_end(sd[p + "end.weight"], sd[p + "end.bias"], b_skip)
    assert util.rel_l2(f["b_end"], b_fold) <= 1e-6
    assert torch.equal(f["wt_rs"][3], w_rs[3][:, :, 0].t().bfloat16()) and f["n_half"] == 3


def test_effective_weights_cover_every_parameter():
    """The flat (name, tensor) list handed to the autograd node names every weight / bias of the reference layout once,
    and every model parameter (weight_g / weight_v included) is reachable from it through autograd."""
    import text2speech_b200 as t2s
    from text2speech_b200 import training
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = t2s.WaveGlow(**train_config())
    names, tensors = training.effective_weights(m)
    folded = set(syn.synthetic_state_dict(train_config(), weight_norm=False))
    assert len(names) == len(set(names)) and set(names) == folded
    total = sum((t.float() ** 2).sum() for t in tensors)
    total.backward()
    missing = [n for n, p in m.named_parameters() if p.grad is None]
    assert not missing, missing[:5]
    # weight norm is applied with autograd-visible ops: w = g v / ||v||
    conv = m.WN[0].in_layers[0]
    w = tensors[names.index("WN.0.in_layers.0.weight")]
    v, g = conv.weight_v.detach(), conv.weight_g.detach()
    want = v * (g / v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1))
    assert util.rel_l2(w.detach(), want) <= 1e-6


def test_training_host_sequence_on_emulated_kernels(monkeypatch, golden_grads):
    """The real host code of the training direction (per-step packing, kernel sequencing, buffer slicing, parameter
    algebra, autograd wiring) run on the CPU with every C-ABI entry point replaced by a torch stand-in of its
    contract (tests/emulate_train.py): loss and every parameter gradient against the unmodified reference's."""
    import text2speech_b200 as t2s
    from tests import emulate_train
    from text2speech_b200 import training
    calls = emulate_train.install(monkeypatch)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = t2s.WaveGlow(**train_config())
    m.load_state_dict(_train_state())
    m.train()
    mel, wav = train_inputs()
    outputs = training.forward_autograd(m, mel, wav)
    loss = t2s.WaveGlowLoss(SIGMA)(outputs)
    loss.backward()
    assert abs(float(loss.detach()) - float(golden_grads["loss"])) <= 2e-3 * abs(float(golden_grads["loss"]))
    named = [(n, p.grad) for n, p in m.named_parameters()]
    assert all(g is not None for _, g in named)
    worst, bad = _compare(named, golden_grads, 1.5e-2, 1.5e-2)
    assert not bad, f"{len(bad)} gradients off: {bad[:8]} (worst per kind: {worst})"
    # the sequence itself: 4 flows x 8 layers
    assert calls.count("wgb_tc2_wn_gate_train") == 32 and calls.count("wgb_tc2_wn_res_taps") == 32
    assert calls.count("wgb_tc_gemm_seg") == 4 and calls.count("wgb_upsample_wgrad") == 1
    assert calls.index("wgb_coupling_bwd") > calls.index("wgb_flow_to_z")


def test_training_refuses_cpu():
    """No CPU fallback in the training direction either."""
    import text2speech_b200 as t2s
    from text2speech_b200 import train as t2s_train
    from text2speech_b200.training import FusedAdam
    with pytest.raises(RuntimeError):
        FusedAdam([torch.nn.Parameter(torch.zeros(4))])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = t2s.WaveGlow(**train_config()).train()
    mel, wav = train_inputs()
    with pytest.raises(RuntimeError):
        m((mel, wav))
    with pytest.raises(ValueError):
        t2s_train.train(1, 0, "", "out", 1, 1e-4, 1.0, 10, 2, 1234, "", waveglow_config=train_config(), data_config={},
                        fp16_run=True)


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def lib():
    from text2speech_b200 import _lib
    _lib.require_b200(torch.device(DEV))
    return _lib


@pytest.mark.gpu
@pytest.mark.parametrize("ca,cb,taps,dil,b,t", [(1024, 512, 3, 4, 3, 200), (512, 512, 1, 1, 2, 333), (1024, 640, 1, 1, 2, 128),
                                                (1024, 512, 3, 128, 2, 300), (64, 512, 1, 1, 2, 1100), (192, 512, 1, 1, 1, 700)])
def test_tc_wgrad_matches_torch(lib, ca, cb, taps, dil, b, t):
    g = torch.Generator().manual_seed(ca + cb + taps)
    gy = torch.randn((b, t, ca), generator=g).bfloat16()
    x = torch.randn((b, t, cb), generator=g).bfloat16()
    want = torch.zeros(taps, ca, cb, dtype=torch.float64)
    for tap in range(taps):
        sh = (tap - (taps - 1) // 2) * dil
        xs = torch.zeros_like(x, dtype=torch.float64)
        if sh >= 0:
            xs[:, : t - sh] = x[:, sh:].double() if sh < t else 0
        else:
            xs[:, -sh:] = x[:, : t + sh].double()
        want[tap] = torch.einsum("btm,btn->mn", gy.double(), xs)
    dw = torch.full((taps, ca, cb), 7.0, device=DEV)
    lib.call("wgb_tc_wgrad", gy.to(DEV), x.to(DEV), dw, b, t, ca, cb, taps, dil, 0, lib.stream_ptr())
    assert util.rel_l2(dw.cpu(), want) <= 1e-4
    lib.call("wgb_tc_wgrad", gy.to(DEV), x.to(DEV), dw, b, t, ca, cb, taps, dil, 1, lib.stream_ptr())
    assert util.rel_l2(dw.cpu(), 2 * want) <= 1e-4


@pytest.mark.gpu
def test_tc_gemm_seg_taps_residual_and_two_operands(lib):
    g = torch.Generator().manual_seed(5)
    b, t, c, n, d = 2, 300, 1024, 512, 8
    a = torch.randn((b, t, c), generator=g).bfloat16()
    w = (torch.randn((n, 3 * c), generator=g) / 40).bfloat16()
    res = torch.randn((b, t, n), generator=g).bfloat16()
    want = res.double().clone()
    for s in range(3):
        sh = (s - 1) * d
        a_s = torch.zeros((b, t, c), dtype=torch.float64)
        if sh >= 0:
            a_s[:, : t - sh] = a[:, sh:].double()
        else:
            a_s[:, -sh:] = a[:, : t + sh].double()
        want += a_s @ w[:, s * c:(s + 1) * c].double().t()
    out = res.to(DEV).clone()
    lib.call("wgb_tc_gemm_seg", a.to(DEV), None, 3, 0, w.to(DEV), None, out, out, 1, b, t, n, c, -d, d, 0, 0, lib.stream_ptr())
    assert util.rel_l2(out.float().cpu(), want) <= 4e-3                     # bf16 output rounding
    # two operands (segments [a0 | a1]), fp32 output accumulated in place
    a1 = torch.randn((b, t, c), generator=g).bfloat16()
    w2 = (torch.randn((256, 2 * c), generator=g) / 40).bfloat16()
    acc0 = torch.randn((b, t, 256), generator=g)
    want2 = acc0.double() + a.double() @ w2[:, :c].double().t() + a1.double() @ w2[:, c:].double().t()
    acc = acc0.to(DEV).clone()
    lib.call("wgb_tc_gemm_seg", a.to(DEV), a1.to(DEV), 2, 0b10, w2.to(DEV), None, acc, acc, 0, b, t, 256, c, 0, 0, 0, 0,
             lib.stream_ptr())
    assert util.rel_l2(acc.cpu(), want2) <= 1e-5
    # the same product with the two operands stacked as planes of one tensor
    acc = acc0.to(DEV).clone()
    lib.call("wgb_tc_gemm_seg", torch.stack([a, a1]).to(DEV), None, 2, 0, w2.to(DEV), None, acc, acc, 0, b, t, 256, c, 0, 0, 0, 1,
             lib.stream_ptr())
    assert util.rel_l2(acc.cpu(), want2) <= 1e-5


@pytest.mark.gpu
def test_gate_train_stores_tanh_and_sigmoid(lib):
    from text2speech_b200.packing import gate_row_order
    g = torch.Generator().manual_seed(6)
    b, t = 2, 200
    h = torch.randn((b, t, 512), generator=g).bfloat16()
    cond = torch.randn((b, t, 640), generator=g).bfloat16()
    w = (torch.randn((1024, 2176), generator=g) / 50).bfloat16()
    bias = torch.randn(1024, generator=g) * 0.1
    order = gate_row_order(512)
    acts = torch.empty((b, t, 512), device=DEV, dtype=torch.bfloat16)
    acts2 = torch.empty_like(acts)
    ts = torch.empty((b, t, 1024), device=DEV, dtype=torch.bfloat16)
    args = (h.to(DEV), cond.to(DEV), w[order].contiguous().to(DEV), bias[order].contiguous().to(DEV))
    lib.call("wgb_tc2_wn_gate_train", *args, acts, ts, b, t, 2, lib.stream_ptr())
    lib.call("wgb_tc2_wn_gate", *args, acts2, b, t, 2, lib.stream_ptr())
    assert torch.equal(acts, acts2)
    hp = torch.zeros((b, t + 4, 512), dtype=torch.float64)
    hp[:, 2:-2] = h.double()
    u = bias.double() + cond.double() @ w[:, 1536:].double().t()
    for tap in range(3):
        u = u + hp[:, 2 * tap: 2 * tap + t] @ w[:, tap * 512:(tap + 1) * 512].double().t()
    want = torch.cat([torch.tanh(u[..., :512]), torch.sigmoid(u[..., 512:])], dim=-1)
    assert util.rel_l2(ts.float().cpu(), want) <= 4e-3


@pytest.mark.gpu
def test_pointwise_backward_kernels(lib):
    g = torch.Generator().manual_seed(8)
    rows, nh = 1000, 3
    c, base = 2 * nh, 8 - 2 * nh
    s = lib.stream_ptr()
    # gate backward
    ga = torch.randn((rows, 512), generator=g).bfloat16()
    tt = torch.tanh(torch.randn((rows, 512), generator=g)).bfloat16()
    ss = torch.sigmoid(torch.randn((rows, 512), generator=g)).bfloat16()
    ts = torch.cat([tt, ss], dim=1).to(DEV).contiguous()
    db = torch.empty(1024, device=DEV)
    lib.call("wgb_gate_bwd", ga.to(DEV), ts, db, rows, 512, s)
    want = torch.cat([ga.float() * ss.float() * (1 - tt.float() ** 2), ga.float() * tt.float() * ss.float() * (1 - ss.float())], 1)
    assert util.rel_l2(ts.float().cpu(), want) <= 4e-3
    assert util.rel_l2(db.cpu(), ts.float().cpu().double().sum(0)) <= 1e-5
    # coupling backward
    b, t = 2, 500
    g_x = torch.randn((rows, 8), generator=g)
    x_mix = torch.randn((rows, 8), generator=g)
    log_s = 0.3 * torch.randn((b, nh, t), generator=g)
    g_ls = torch.randn((b, nh, t), generator=g)
    w_end_t = torch.zeros(512, 8)
    w_end_t[:, :c] = torch.randn((512, c), generator=g)
    gx = g_x.to(DEV).clone()
    g_out = torch.empty((rows, 8), device=DEV)
    g_skip = torch.empty((rows, 512), device=DEV, dtype=torch.bfloat16)
    stack = torch.empty((rows, 64), device=DEV, dtype=torch.bfloat16)
    lib.call("wgb_coupling_bwd", gx, x_mix.to(DEV), log_s.to(DEV), g_ls.to(DEV), w_end_t.to(DEV), g_out, g_skip, stack, b, t,
             512, nh, s)
    ls_rows = log_s.permute(0, 2, 1).reshape(rows, nh)
    gls_rows = g_ls.permute(0, 2, 1).reshape(rows, nh)
    ga1p = g_x[:, base + nh:]
    want_out = torch.zeros(rows, 8)
    want_out[:, :nh] = ga1p
    want_out[:, nh:c] = ga1p * x_mix[:, base + nh:] * ls_rows.exp() + gls_rows
    want_gx = g_x.clone()
    want_gx[:, base + nh:] = ga1p * ls_rows.exp()
    assert util.rel_l2(g_out.cpu(), want_out) <= 1e-6
    assert util.rel_l2(gx.cpu(), want_gx) <= 1e-6
    assert util.rel_l2(g_skip.float().cpu(), want_out @ w_end_t.t()) <= 4e-3
    # the [rows, 64] stack: hi + lo parts reproduce g_out / x_mix to fp32 accuracy, column 32 is one
    st = stack.float().cpu()
    assert util.rel_l2(st[:, 0:8] + st[:, 8:16], want_out) <= 2e-5
    assert util.rel_l2(st[:, 16:24] + st[:, 24:32], x_mix) <= 2e-5
    assert torch.equal(st[:, 32], torch.ones(rows)) and float(st[:, 33:].abs().max()) == 0.0
    # the same reductions the backward pass runs on the tensor cores: stack^T b -> rows 0..7 + 8..15 = g_out^T b,
    # row 32 = column sums of b
    bb = torch.randn((rows, 512), generator=g).bfloat16()
    prod = torch.empty((1, 64, 512), device=DEV)
    lib.call("wgb_tc_wgrad", stack.view(b, t, 64), bb.to(DEV).view(b, t, 512), prod, b, t, 64, 512, 1, 1, 0, s)
    prod = prod[0].cpu()
    assert util.rel_l2(prod[0:8] + prod[8:16], want_out.double().t() @ bb.double()) <= 1e-5
    assert util.rel_l2(prod[32], bb.double().sum(0)) <= 1e-5
    c8 = torch.empty(8, device=DEV)
    lib.call("wgb_colsum8_f32", g_out, c8, rows, 0, s)
    assert util.rel_l2(c8.cpu(), want_out.double().sum(0)) <= 1e-5
    # WN.start backward
    w_start = torch.randn((512, nh), generator=g)
    gx2 = g_x.to(DEV).clone()
    lib.call("wgb_start_bwd", gx2, bb.to(DEV), w_start.to(DEV), rows, 512, nh, s)
    want2 = g_x.clone()
    want2[:, base: base + nh] += (bb.double() @ w_start.double()).float()
    assert util.rel_l2(gx2.cpu(), want2) <= 1e-5
    # 1x1 mix backward
    w = torch.zeros(8, 8)
    w[:c, :c] = torch.randn((c, c), generator=g)
    gx3 = g_x.to(DEV).clone()
    dw = torch.empty((8, 8), device=DEV)
    lib.call("wgb_mix_bwd", gx3, x_mix.to(DEV), w.to(DEV), dw, rows, c, s)
    want3 = g_x.clone()
    want3[:, base:] = g_x[:, base:] @ w[:c, :c]
    assert util.rel_l2(gx3.cpu(), want3) <= 1e-5
    assert util.rel_l2(dw[:c, :c].cpu(), g_x[:, base:].double().t() @ x_mix[:, base:].double()) <= 1e-5


@pytest.mark.gpu
def test_upsample_wgrad_matches_autograd(lib):
    g = torch.Generator().manual_seed(9)
    b, f, t = 2, 5, 150                                   # t < 32 f: the trimmed branch (glow.py:216-218)
    mel = torch.randn((b, 80, f), generator=g)
    w = torch.randn((80, 80, 1024), generator=g, requires_grad=True)
    bias = torch.zeros(80, requires_grad=True)
    g_cond = torch.zeros((b, t, 768))
    g_cond[:, :, :640] = torch.randn((b, t, 640), generator=g)
    up = torch.nn.functional.conv_transpose1d(mel, w, bias, stride=256)[:, :, : t * 8]
    cond = up.reshape(b, 80, t, 8).permute(0, 2, 1, 3).reshape(b, t, 640)
    (cond * g_cond[:, :, :640]).sum().backward()
    dw = torch.empty((80, 80, 1024), device=DEV)
    db = torch.empty(80, device=DEV)
    lib.call("wgb_upsample_wgrad", mel.to(DEV), g_cond.to(DEV), dw, db, b, 80, f, t, 768, 1024, 256, 8, lib.stream_ptr())
    assert util.rel_l2(dw.cpu(), w.grad) <= 1e-5
    assert util.rel_l2(db.cpu(), bias.grad) <= 1e-5


@pytest.fixture(scope="module")
def train_model(lib):
    import text2speech_b200 as t2s
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = t2s.WaveGlow(**train_config())
    m.load_state_dict(_train_state())
    return m.to(DEV).train()


@pytest.mark.gpu
def test_parameter_gradients_match_reference(train_model, golden_grads):
    """criterion(model((mel, audio))).backward() on the drop-in model vs the unmodified reference's autograd."""
    import text2speech_b200 as t2s
    m = train_model
    m.zero_grad()
    mel, wav = train_inputs()
    outputs = m((mel.to(DEV), wav.to(DEV)))
    loss = t2s.WaveGlowLoss(SIGMA)(outputs)
    loss.backward()
    assert abs(float(loss) - float(golden_grads["loss"])) <= 2e-3 * abs(float(golden_grads["loss"]))
    named = [(n, p.grad) for n, p in m.named_parameters()]
    assert all(g is not None for _, g in named), [n for n, g in named if g is None][:5]
    worst, bad = _compare(named, golden_grads, 1.5e-2, 1.5e-2)
    print("worst relative error per parameter kind:", {k: round(v, 4) for k, v in sorted(worst.items())})
    assert not bad, f"{len(bad)} gradients off: {bad[:8]} (worst per kind: {worst})"


@pytest.mark.gpu
def test_eval_forward_is_unchanged_and_train_forward_agrees(train_model):
    m = train_model
    mel, wav = train_inputs()
    z_train, log_s_train, log_det_train = m((mel.to(DEV), wav.to(DEV)))
    m.eval()
    with torch.no_grad():
        z_eval, log_s_eval, log_det_eval = m((mel.to(DEV), wav.to(DEV)))
    m.train()
    assert not z_eval.requires_grad and z_train.requires_grad
    assert util.snr_db(z_train.detach().cpu(), z_eval.cpu()) >= 40.0
    assert util.snr_db(log_s_train[-1].detach().cpu(), log_s_eval[-1].cpu()) >= 40.0
    assert abs(float(log_det_train[0]) - float(log_det_eval[0])) <= 1e-3


@pytest.mark.gpu
def test_fused_adam_matches_torch_adam(lib):
    from text2speech_b200.training import FusedAdam
    g = torch.Generator().manual_seed(10)
    shapes = [(7, 3), (130,), (5, 5, 5)]
    a = [torch.randn(s, generator=g).to(DEV).requires_grad_(True) for s in shapes]
    b = [p.detach().clone().requires_grad_(True) for p in a]
    opt_a = FusedAdam(a, lr=1e-2)
    opt_b = torch.optim.Adam(b, lr=1e-2)
    for step in range(5):
        for pa, pb in zip(a, b):
            gr = torch.randn(pa.shape, generator=g).to(DEV)
            pa.grad = gr.clone()
            pb.grad = gr.clone()
        opt_a.step()
        opt_b.step()
    for pa, pb in zip(a, b):
        assert util.rel_l2(pa.detach().cpu(), pb.detach().cpu()) <= 1e-6


@pytest.mark.gpu
def test_step_along_the_gradient_lowers_the_loss_as_predicted(lib):
    """Size-independent property: loss(theta - eps g) - loss(theta) = -eps |g|^2 to first order.  eps is chosen so
    that the predicted drop is 3 % of the loss; the realised drop must be 0.5x .. 1.5x of it."""
    import text2speech_b200 as t2s
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = t2s.WaveGlow(**train_config())
    m.load_state_dict(_train_state())
    m = m.to(DEV).train()
    crit = t2s.WaveGlowLoss(SIGMA)
    mel, wav = train_inputs()
    mel, wav = mel.to(DEV), wav.to(DEV)
    loss0 = crit(m((mel, wav)))
    loss0.backward()
    g2 = sum(float(p.grad.double().pow(2).sum()) for p in m.parameters())
    predicted = 0.03 * abs(float(loss0.detach()))
    eps = predicted / g2
    with torch.no_grad():
        for p in m.parameters():
            p -= eps * p.grad
    loss1 = crit(m((mel, wav)))
    drop = float(loss0.detach()) - float(loss1.detach())
    assert 0.5 * predicted <= drop <= 1.5 * predicted, (float(loss0), float(loss1), predicted)


@pytest.mark.gpu
def test_training_loop_with_fused_adam_and_flat_allreduce(lib):
    """train.py:108-124 in miniature (zero_grad / forward / loss / backward / all-reduce / step): every parameter
    moves by at most lr per step (Adam's bound) and the loss stays finite."""
    import text2speech_b200 as t2s
    from text2speech_b200.training import FusedAdam, allreduce_gradients
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = t2s.WaveGlow(**train_config())
    m.load_state_dict(_train_state())
    m = m.to(DEV).train()
    before = [p.detach().clone() for p in m.parameters()]
    lr = 1e-6
    opt = FusedAdam(m.parameters(), lr=lr)
    crit = t2s.WaveGlowLoss(SIGMA)
    mel, wav = train_inputs()
    mel, wav = mel.to(DEV), wav.to(DEV)
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss = crit(m((mel, wav)))
        loss.backward()
        scale = allreduce_gradients(opt)
        opt.step(grad_scale=scale, gathered=True)
        losses.append(float(loss.detach()))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses      # a small enough step lowers the NLL
    moved = max(float((p.detach() - q).abs().max()) for p, q in zip(m.parameters(), before))
    assert 0.5 * lr <= moved <= 3.2 * lr, moved


def test_packed_weight_cache_sees_raw_kernel_updates():
    """FusedAdam rewrites the parameters with a raw kernel behind torch's version counters and data pointers; the packed
    weights of WaveGlow are keyed on a generation counter that every such step bumps (CPU: the key logic only)."""
    import text2speech_b200 as t2s
    from text2speech_b200 import packing
    cfg = train_config()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = t2s.WaveGlow(**cfg)
    sig0 = m._signature()
    assert m._signature() == sig0
    packing.bump_param_generation()
    assert m._signature() != sig0
    import copy
    m._pack_cache = {"dummy": (sig0, object())}
    assert copy.deepcopy(m)._pack_cache == {}                 # copies / pickles carry parameters only


@pytest.mark.gpu
def test_infer_after_training_steps_uses_the_new_weights(lib):
    """infer -> Adam step -> infer -> graphed step -> infer: every infer sees the current parameters (the packed-weight
    cache used to be keyed on data_ptr / _version only, which FusedAdam's flat-buffer kernel never changes), and the
    audio equals that of a freshly constructed model holding the same state."""
    import text2speech_b200 as t2s
    from text2speech_b200.training import FusedAdam, GraphedTrainStep
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = t2s.WaveGlow(**train_config())
    m.load_state_dict(_train_state())
    m = m.to(DEV).train()
    opt = FusedAdam(m.parameters(), lr=1e-3)                 # large steps: the audio must move visibly
    crit = t2s.WaveGlowLoss(SIGMA)
    mel, wav = train_inputs()
    mel, wav = mel.to(DEV), wav.to(DEV)
    z = syn.synthetic_z(mel.shape[0], mel.shape[2], seed=5).to(DEV)

    def sample():
        m.eval()
        with torch.no_grad():
            out = m.infer(mel, sigma=SIGMA, z=z).clone()
        m.train()
        return out

    def fresh():
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            f = t2s.WaveGlow(**train_config())
        f.load_state_dict({k: v.detach().clone() for k, v in m.state_dict().items()})
        return f.to(DEV).eval().infer(mel, sigma=SIGMA, z=z)

    a0 = sample()
    opt.zero_grad()
    crit(m((mel, wav))).backward()
    opt.step()
    a1 = sample()
    assert not torch.equal(a1, a0) and util.snr_db(a1.cpu(), a0.cpu()) < 60.0
    assert torch.equal(a1, fresh())
    step = GraphedTrainStep(m, opt, crit, mel.shape[0], mel.shape[1], mel.shape[2], wav.shape[1])
    step(mel, wav)
    a2 = sample()
    assert not torch.equal(a2, a1)
    assert torch.equal(a2, fresh())


@pytest.mark.gpu
def test_logdet_kernel_matches_torch(lib):
    from text2speech_b200.training import _LogDet
    g = torch.Generator().manual_seed(13)
    for c in (8, 6, 4, 2):
        w = torch.randn((c, c, 1), generator=g).to(DEV).requires_grad_(True)
        w_ref = w.detach().clone().requires_grad_(True)
        if float(torch.det(w_ref.detach().squeeze(-1))) < 0:
            with torch.no_grad():
                w[:, 0] = -w[:, 0]
                w_ref[:, 0] = -w_ref[:, 0]
        got = _LogDet.apply(w, 3.0)
        want = 3.0 * torch.logdet(w_ref.squeeze(-1))
        assert abs(float(got) - float(want)) <= 1e-5 * max(1.0, abs(float(want)))
        got.backward()
        want.backward()
        assert util.rel_l2(w.grad.cpu(), w_ref.grad.cpu()) <= 1e-5
        # det W < 0: torch.logdet (what glow.py:100 calls) is NaN, and so is the kernel
        sign = float(torch.linalg.slogdet(w_ref.detach().squeeze(-1))[0])
        flipped = w.detach().clone()
        flipped[:, 0] = -flipped[:, 0]
        for m, s in ((w.detach(), sign), (flipped, -sign)):
            val = float(_LogDet.apply(m, 1.0))
            assert np.isnan(val) == (s < 0), (c, s, val)
            assert np.isnan(float(torch.logdet(m.squeeze(-1)))) == (s < 0)


@pytest.mark.gpu
def test_graphed_train_step_matches_eager(lib):
    """The CUDA-graph replay of a whole step (forward, loss, backward, gradient gather, Adam) follows the eager loop."""
    import text2speech_b200 as t2s
    from text2speech_b200.training import FusedAdam, GraphedTrainStep
    mel, wav = train_inputs()
    mel, wav = mel.to(DEV), wav.to(DEV)
    crit = t2s.WaveGlowLoss(SIGMA)
    runs = []
    for graphed in (False, True):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m = t2s.WaveGlow(**train_config())
        m.load_state_dict(_train_state())
        m = m.to(DEV).train()
        opt = FusedAdam(m.parameters(), lr=1e-6)
        losses = []
        if graphed:
            step = GraphedTrainStep(m, opt, crit, mel.shape[0], mel.shape[1], mel.shape[2], wav.shape[1])
            for _ in range(3):
                losses.append(float(step(mel, wav)))
            assert opt.step_count == 3 and int(opt.step_dev.item()) == 3
        else:
            for _ in range(3):
                opt.zero_grad()
                loss = crit(m((mel, wav)))
                loss.backward()
                opt.step()
                losses.append(float(loss.detach()))
        runs.append(losses)
    eager, graph = runs
    assert eager[-1] < eager[0] and graph[-1] < graph[0], runs
    for a, b in zip(eager, graph):
        assert abs(a - b) <= 0.02 * abs(eager[0]) + 1e-4, runs


@pytest.mark.gpu
def test_gradients_match_oracle_six_flows_ragged(lib):
    """All three coupling widths of config.json (n_half 4, 3, 2), one utterance, audio shorter than 256 * frames
    (trimmed upsample, glow.py:216-218) and a group-step count that is not a multiple of the 64-row K chunk:
    every parameter gradient vs the CPU oracle under autograd (itself pinned to the reference by the 4-flow golden)."""
    import text2speech_b200 as t2s
    from oracle import waveglow_oracle as wo
    cfg = dict(syn.load_config())
    cfg.update(n_flows=6, n_early_every=2, n_early_size=2)
    sd = syn.synthetic_state_dict(cfg, seed=91, end_std=0.05, weight_norm=True)
    frames, n = 5, 5 * 256 - 72                                   # T = 151 group steps
    mel = syn.synthetic_mel(1, frames, seed=31)
    g = torch.Generator().manual_seed(32)
    wav = (0.1 * torch.randn((1, n), generator=g)).clamp(-1, 1)
    ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        z, log_s, log_det = wo.waveglow_forward(ref, mel, wav)
        loss_ref = wo.waveglow_loss(z, log_s, log_det, SIGMA)
        loss_ref.backward()
        m = t2s.WaveGlow(**cfg)
    m.load_state_dict(sd)
    m = m.to(DEV).train()
    loss = t2s.WaveGlowLoss(SIGMA)(m((mel.to(DEV), wav.to(DEV))))
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_ref.detach())) <= 2e-3 * abs(float(loss_ref.detach())) + 1e-4
    worst = {}
    for name, p in m.named_parameters():
        want = ref[name].grad
        err = util.rel_l2(p.grad.cpu(), want)
        kind = ".".join(name.split(".")[2:]) if name.startswith("WN") else name.split(".")[0]
        worst[kind] = max(worst.get(kind, 0.0), err)
        assert err <= 2e-2, (name, err)
    print("six flows, worst relative error per parameter kind:", {k: round(v, 4) for k, v in sorted(worst.items())})


@pytest.mark.gpu
def test_train_cli_checkpoints_and_resume(tmp_path, lib):
    """waveglow/train.py's loop on three synthetic wav files: per-iteration losses, checkpoints in the reference's
    format ({'model': pickled module, 'iteration', 'optimizer' (torch Adam layout), 'learning_rate'}), resume."""
    from scipy.io.wavfile import write
    from text2speech_b200 import train as t2s_train
    from text2speech_b200.inference import load_waveglow
    g = torch.Generator().manual_seed(40)
    names = []
    for i in range(3):
        path = str(tmp_path / f"u{i}.wav")
        write(path, 22050, (0.2 * torch.randn(9000 + 500 * i, generator=g)).clamp(-1, 1).mul(32767).short().numpy())
        names.append(path)
    flist = tmp_path / "train_files.txt"
    flist.write_text("\n".join(names) + "\n")
    data_config = dict(training_files=str(flist), segment_length=4096, sampling_rate=22050, filter_length=1024,
                       hop_length=256, win_length=1024, mel_fmin=0.0, mel_fmax=8000.0)
    out_dir = str(tmp_path / "ckpt")
    common = dict(output_directory=out_dir, learning_rate=1e-5, sigma=1.0, iters_per_checkpoint=1, batch_size=2, seed=1234,
                  waveglow_config=train_config(), data_config=data_config)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        losses = t2s_train.train(1, 0, "", epochs=2, checkpoint_path="", **common)
        assert len(losses) == 2 and all(np.isfinite(losses))
        ck = torch.load(os.path.join(out_dir, "waveglow_1"), map_location="cpu", weights_only=False)
        assert ck["iteration"] == 1 and ck["learning_rate"] == 1e-5
        assert set(ck["optimizer"]) == {"state", "param_groups"} and int(ck["optimizer"]["state"][0]["step"]) == 2
        assert type(ck["model"]).__name__ == "WaveGlow"
        # resume: one more epoch = one more iteration, Adam moments and step restored
        more = t2s_train.train(1, 0, "", epochs=3, checkpoint_path=os.path.join(out_dir, "waveglow_1"), **common)
        assert len(more) == 1 and np.isfinite(more[0])
        ck2 = torch.load(os.path.join(out_dir, "waveglow_2"), map_location="cpu", weights_only=False)
        assert ck2["iteration"] == 2 and int(ck2["optimizer"]["state"][0]["step"]) == 3
        # the checkpoint is a vocoder checkpoint: the inference loader takes it
        m = load_waveglow(os.path.join(out_dir, "waveglow_2")).to(DEV).eval()
    mel = syn.synthetic_mel(1, 4, seed=1).to(DEV)
    audio = m.infer(mel, sigma=0.6)
    assert audio.shape == (1, 1024) and bool(torch.isfinite(audio).all())
