"""Torch-on-CPU stand-ins for the C-ABI entry points the inference engine calls (test helper; see emulate_train.py).

``install(monkeypatch)`` routes ``text2speech_b200._lib.call`` through these, so the REAL ``engine.infer`` /
``engine.forward`` run on the CPU: the padded frame layout with its guard rows, the composed-conditioning operands,
the first-layer fold (x_stack), the skip+end kernel's fused next-flow start / 1x1 mix and the ping-pong of the residual
stream are checked against the oracle without a GPU (tests/test_packing.py).
"""
import contextlib

import torch

from tests.emulate import shift_rows
from tests.emulate_train import flow_mix, flow_to_z, tc_gemm, upsample_im2col, wn_start_padded

PHASES = 32


def _f(t):
    return t.float()


def _gate(u):
    """Packed row order: pass p = tanh rows 128p.. followed by the matching sigmoid rows -> acts [.., 512]."""
    outs = []
    for p in range(u.shape[-1] // 256):
        blk = u[..., p * 256:(p + 1) * 256]
        outs.append(torch.tanh(blk[..., :128]) * torch.sigmoid(blk[..., 128:]))
    return torch.cat(outs, dim=-1)


def flow_from_z(z, x, b, t, sigma, s):
    x.copy_(sigma * z.permute(0, 2, 1))


def x_stack(x, out, b, t, out_batch_rows, n_half, s):
    base = 8 - 2 * n_half
    a0 = torch.zeros(b, t, 4)
    a0[:, :, :n_half] = x[:, :, base: base + n_half]
    row = torch.zeros(b, t, 64)
    ones = torch.ones(b, t, 1)
    for tap in range(3):
        v = shift_rows(a0, tap - 1)
        hi = v.bfloat16().float()
        row[:, :, tap * 4: tap * 4 + 4] = hi
        row[:, :, 12 + tap * 4: 16 + tap * 4] = (v - hi).bfloat16().float()
        row[:, :, 24 + tap * 4: 28 + tap * 4] = hi
        inside = shift_rows(ones, tap - 1)[:, :, 0]
        row[:, :, 36 + tap] = inside
        row[:, :, 39 + tap] = inside
    out[:, :t].copy_(row)


def _mel_term(mel_stack, w_mel, b, t):
    """sum_k V[phase(t)][row][k] mel_stack[b, frame(t), k] for every group step t = 32 frame + phase."""
    frames = -(-t // PHASES)
    ms = _f(mel_stack)[:, :frames]                                            # [B, frames, 320]
    term = torch.einsum("bfk,pnk->bfpn", ms, _f(w_mel))                       # [B, frames, 32, 1024]
    return term.reshape(b, frames * PHASES, -1)[:, :t]


def gate_mel(h, mel_stack, w_packed, w_mel, bias, acts, b, t, frames_pad, dilation, w_comp, skip_acc, skip_first, s):
    assert w_comp is None and skip_acc is None
    hf = _f(h)[:, :t]                                   # rows >= t of the padded layout are guard rows (zero)
    assert float(_f(h)[:, t:].abs().max() if h.shape[1] > t else 0.0) == 0.0, "a kernel wrote into the guard rows"
    w = _f(w_packed)[:, : 3 * 512]
    a = torch.cat([shift_rows(hf, -dilation), hf, shift_rows(hf, dilation)], dim=2)
    acts.copy_(_gate(a @ w.t() + _mel_term(mel_stack, w_mel, b, t) + bias))


def gate_mel0(xs, mel_stack, w0, w_mel, bias, acts, b, t, frames_pad, w_comp, skip_acc, skip_first, s):
    assert w_comp is None and skip_acc is None
    acts.copy_(_gate(_f(xs)[:, :t] @ _f(w0).t() + _mel_term(mel_stack, w_mel, b, t) + bias))


def gate(h, cond, w_packed, bias, acts, b, t, dilation, s):
    hf = _f(h)
    a = torch.cat([shift_rows(hf, -dilation), hf, shift_rows(hf, dilation), _f(cond)], dim=2)
    acts.copy_(_gate(a @ _f(w_packed).t() + bias))


def wn_res(acts, w_res, bias, h_in, h_out, b, t, h_rows, w16, skip_acc, skip_first, s):
    assert w16 is None and skip_acc is None
    h_out[:, :t].copy_(_f(h_in)[:, :t] + _f(acts) @ _f(w_res).t() + bias)


def skip16_end(acts_all, n_layers, w16, b_end, x, w_mix, log_s, b, t, n_half, direction, next_w_start, next_b_start,
               next_n_half, h_next, h_next_batch_rows, skip_acc, next_w_mix, s):
    assert skip_acc is None
    comp = _f(w16)[:8] + _f(w16)[8:]
    a = torch.cat([_f(acts_all[i]) for i in range(n_layers)], dim=2)
    out = a @ comp.t() + b_end
    c, base = 2 * n_half, 8 - 2 * n_half
    b_, s_ = out[:, :, :n_half], out[:, :, n_half:c]
    a0, a1 = x[:, :, base: base + n_half].clone(), x[:, :, base + n_half:].clone()
    if direction == 0:                                                        # glow.py:279-282
        xin = torch.cat([a0, (a1 - b_) * torch.exp(-s_)], dim=2)
        x[:, :, base:] = xin @ w_mix[:c, :c].t()
    else:                                                                     # glow.py:241-246
        x[:, :, base + n_half:] = torch.exp(s_) * a1 + b_
        log_s.copy_(s_.permute(0, 2, 1))
        if next_w_mix is not None:
            cn = 2 * next_n_half
            x[:, :, 8 - cn:] = x[:, :, 8 - cn:] @ next_w_mix[:cn, :cn].t()
    if h_next is not None:
        nb = 8 - 2 * next_n_half
        h_next[:, :t].copy_(x[:, :, nb: nb + next_n_half] @ next_w_start.t() + next_b_start)


TABLE = {
    "wgb_upsample_im2col": upsample_im2col, "wgb_tc_gemm": tc_gemm, "wgb_flow_from_z": flow_from_z,
    "wgb_flow_to_z": flow_to_z, "wgb_flow_mix": flow_mix, "wgb_wn_start_padded": wn_start_padded, "wgb_x_stack": x_stack,
    "wgb_tc2_wn_gate_mel": gate_mel, "wgb_tc2_wn_gate_mel0": gate_mel0, "wgb_tc2_wn_gate": gate, "wgb_tc2_wn_res": wn_res,
    "wgb_tc_wn_skip16_end": skip16_end,
}


def install(monkeypatch):
    from text2speech_b200 import _lib
    calls = []

    def call(name, *args):
        calls.append(name)
        with torch.no_grad():
            TABLE[name](*args)

    monkeypatch.setattr(_lib, "call", call)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: 0)
    monkeypatch.setattr(_lib, "require_b200", lambda device: None)
    monkeypatch.setattr(torch.cuda, "device", lambda device: contextlib.nullcontext())
    return calls
