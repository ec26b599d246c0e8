"""Torch-on-CPU stand-ins for the C-ABI entry points the training direction calls (test helper).

``install(monkeypatch)`` replaces ``text2speech_b200._lib.call`` by a dispatcher that implements each entry point's
CONTRACT (include/waveglow_b200.h) with plain tensor ops, writing into the caller's output tensors (views included).
The REAL host code — ``training._pack_flow / _forward / _backward``, the autograd wiring, the parameter-space algebra —
then runs unchanged on the CPU, so the kernel sequencing, buffer slicing, mirrored-tap / transposed weight layouts and
gradient naming are checked against the reference's gradients without a GPU (tests/test_training.py).  Arithmetic is
fp32 on the operands as stored (bf16 tensors are read back as the bf16 values the kernels would see).
"""
import contextlib

import torch
import torch.nn.functional as F

from tests.emulate import shift_rows


def _f(t):
    return t.float()


def upsample_im2col(mel, a, bf16, b, n_mel, f, taps, ld_tap, s):
    out = torch.zeros(b, f, taps, ld_tap)
    for j in range(taps):
        out[:, j:, j, :n_mel] = mel[:, :, : f - j].permute(0, 2, 1)
    a.copy_(out.reshape(a.shape))


def tc_gemm(a, w, bias, c, out_bf16, batch, t, n, k, s):
    c.copy_((_f(a).reshape(batch, t, k) @ _f(w).t() + (0 if bias is None else bias)).reshape(c.shape))


def flow_mix(x, w, rows, c, s):
    xv = x.view(-1, 8)
    xv[:, 8 - c:] = xv[:, 8 - c:] @ w[:c, :c].t()


def wn_start_padded(x, w_start, b_start, h0, bf16, b, t, h_rows, n_ch, n_half, s):
    base = 8 - 2 * n_half
    h0[:, :t].copy_(x[:, :, base: base + n_half] @ w_start.t() + b_start)


def gate_train(h, cond, w_packed, bias, acts, ts, b, t, dilation, s):
    hf = _f(h)
    a = torch.cat([shift_rows(hf, -dilation), hf, shift_rows(hf, dilation), _f(cond)], dim=2)
    u = a @ _f(w_packed).t() + bias
    n_ch = u.shape[2] // 2
    tt, ss = torch.empty(b, t, n_ch), torch.empty(b, t, n_ch)
    for p in range(u.shape[2] // 256):                      # packed rows: pass p = tanh rows 128p.. | sigmoid rows 128p..
        blk = u[:, :, p * 256:(p + 1) * 256]
        tt[:, :, p * 128:(p + 1) * 128] = torch.tanh(blk[:, :, :128])
        ss[:, :, p * 128:(p + 1) * 128] = torch.sigmoid(blk[:, :, 128:])
    acts.copy_(tt * ss)
    ts.copy_(torch.cat([tt, ss], dim=2))


def wn_res(acts, w_res, bias, h_in, h_out, b, t, h_rows, w16, skip_acc, skip_first, s):
    h_out.copy_(_f(h_in) + _f(acts) @ _f(w_res).t() + bias)


def skip16_end(acts_all, n_layers, w16, b_end, x, w_mix, log_s, b, t, n_half, direction, nws, nbs, nnh, h_next, hbr,
               skip_acc, next_w_mix, s):
    assert direction == 1 and w_mix is None and h_next is None and skip_acc is None and next_w_mix is None
    comp = _f(w16)[:8] + _f(w16)[8:]                       # [8, L*512]
    a = torch.cat([_f(acts_all[i]) for i in range(n_layers)], dim=2)
    out = a @ comp.t() + b_end
    base = 8 - 2 * n_half
    b_, s_ = out[:, :, :n_half], out[:, :, n_half: 2 * n_half]
    x[:, :, base + n_half:] = torch.exp(s_) * x[:, :, base + n_half:] + b_
    log_s.copy_(s_.permute(0, 2, 1))


def flow_to_z(x, z, b, t, s):
    z.copy_(x.permute(0, 2, 1))


def coupling_bwd(g_x, x_mix, log_s, g_ls, w_end_t, g_out, g_skip, stack, b, t, n_ch, n_half, s):
    rows = b * t
    c, base = 2 * n_half, 8 - 2 * n_half
    gx, xm = g_x.view(rows, 8), x_mix.reshape(rows, 8)
    ls = log_s.permute(0, 2, 1).reshape(rows, n_half)
    ga1p = gx[:, base + n_half:].clone()
    out = torch.zeros(rows, 8)
    out[:, :n_half] = ga1p
    out[:, n_half:c] = ga1p * xm[:, base + n_half:] * ls.exp()
    if g_ls is not None:
        out[:, n_half:c] += g_ls.permute(0, 2, 1).reshape(rows, n_half)
    gx[:, base + n_half:] = ga1p * ls.exp()
    g_out.copy_(out)
    g_skip.copy_((out @ w_end_t.t()).reshape(g_skip.shape))
    if stack is not None:
        st = torch.zeros(rows, 64)
        hi = out.bfloat16().float()
        st[:, 0:8], st[:, 8:16] = hi, (out - hi).bfloat16().float()
        xh = xm.bfloat16().float()
        st[:, 16:24], st[:, 24:32] = xh, (xm - xh).bfloat16().float()
        st[:, 32] = 1.0
        stack.copy_(st.reshape(stack.shape))


def colsum8_f32(a, out, rows, accumulate, s):
    v = a.reshape(rows, 8).sum(0)
    out.copy_(out + v if accumulate else v)


def tc_wgrad(g, x, dw, b, t, ca, cb, taps, dilation, accumulate, s):
    gf, xf = _f(g).reshape(b, t, ca), _f(x).reshape(b, t, cb)
    res = torch.stack([torch.einsum("btm,btn->mn", gf, shift_rows(xf, (tap - (taps - 1) // 2) * dilation))
                       for tap in range(taps)])
    dwv = dw.view(taps, ca, cb)
    dwv.copy_(dwv + res if accumulate else res)


def res_seg(a0, a1, n_seg, seg_mask, w, bias, h_in, h_out, b, t, h_rows, c, shift0, dshift, s):
    wf = _f(w)
    out = _f(h_in).clone() + bias
    for seg in range(n_seg):
        src = _f(a1 if (seg_mask >> seg) & 1 else a0).reshape(b, t, c)
        out = out + shift_rows(src, shift0 + seg * dshift) @ wf[:, seg * c:(seg + 1) * c].t()
    h_out.copy_(out)


def res_taps(a, w, bias, h_in, h_out, b, t, h_rows, c, taps, dilation, s):
    res_seg(a, None, taps, 0, w, bias, h_in, h_out, b, t, h_rows, c, -((taps - 1) // 2) * dilation, dilation, s)


def gate_bwd(g_acts, ts, db, rows, n_ch, s):
    g = _f(g_acts).reshape(rows, n_ch)
    tsv = ts.view(rows, 2 * n_ch)
    tt, ss = _f(tsv[:, :n_ch]), _f(tsv[:, n_ch:])
    out = torch.cat([g * ss * (1 - tt * tt), g * tt * ss * (1 - ss)], dim=1)
    tsv.copy_(out)
    if db is not None:
        db.copy_(_f(tsv).sum(0))                          # sums of what was stored (bf16), like the kernel


def gemm_seg(a0, a1, n_seg, seg_mask, w, bias, res, c_out, out_bf16, b, t, n, c, shift0, dshift, act, stacked, s):
    wf = _f(w)
    out = torch.zeros(b, t, n) + (0 if bias is None else bias)
    for seg in range(n_seg):
        if stacked:
            src = _f(a0[seg]).reshape(b, t, c)
        else:
            src = _f(a1 if (seg_mask >> seg) & 1 else a0).reshape(b, t, c)
        out = out + shift_rows(src, shift0 + seg * dshift) @ wf[:, seg * c:(seg + 1) * c].t()
    if act == 1:
        out = torch.tanh(out)
    elif act == 2:
        out = torch.relu(out)
    if res is not None:
        out = out + _f(res).reshape(b, t, n)
    c_out.copy_(out.reshape(c_out.shape))


def start_bwd(g_x, g_h0, w_start, rows, n_ch, n_half, s):
    base = 8 - 2 * n_half
    gx = g_x.view(rows, 8)
    gx[:, base: base + n_half] += _f(g_h0).reshape(rows, n_ch) @ w_start


def mix_bwd(g_x, x_pre, w, dw, rows, c, s):
    base = 8 - c
    gx, xp = g_x.view(rows, 8), x_pre.reshape(rows, 8)
    gy = gx[:, base:].clone()
    gx[:, base:] = gy @ w[:c, :c]
    dw.zero_()
    dw[:c, :c] = gy.t() @ xp[:, base:]


def upsample_wgrad(mel, g_cond, dw, db, b, n_mel, frames, t, ld, ksize, stride, n_group, s):
    w = torch.zeros(n_mel, n_mel, ksize, requires_grad=True)
    bias = torch.zeros(n_mel, requires_grad=True)
    up = F.conv_transpose1d(mel, w, bias, stride=stride)[:, :, : t * n_group]
    cond = up.reshape(b, n_mel, t, n_group).permute(0, 2, 1, 3).reshape(b, t, n_mel * n_group)
    (cond * g_cond[:, :, : n_mel * n_group]).sum().backward()
    dw.copy_(w.grad)
    db.copy_(bias.grad)


def logdet(w, out, inv_t, c, s):
    w2 = w.reshape(c, c).double()
    out[0] = torch.logdet(w2).float()
    inv_t.copy_(torch.linalg.inv(w2).t().float())


TABLE = {
    "wgb_upsample_im2col": upsample_im2col, "wgb_tc_gemm": tc_gemm, "wgb_flow_mix": flow_mix,
    "wgb_wn_start_padded": wn_start_padded, "wgb_tc2_wn_gate_train": gate_train, "wgb_tc2_wn_res": wn_res,
    "wgb_tc_wn_skip16_end": skip16_end, "wgb_flow_to_z": flow_to_z, "wgb_coupling_bwd": coupling_bwd,
    "wgb_colsum8_f32": colsum8_f32, "wgb_tc_wgrad": tc_wgrad, "wgb_tc2_wn_res_seg": res_seg,
    "wgb_tc2_wn_res_taps": res_taps, "wgb_gate_bwd": gate_bwd, "wgb_tc_gemm_seg": gemm_seg, "wgb_start_bwd": start_bwd,
    "wgb_mix_bwd": mix_bwd, "wgb_upsample_wgrad": upsample_wgrad, "wgb_logdet": logdet,
}


def install(monkeypatch):
    """Route text2speech_b200._lib through the table above; returns the list of entry points called (in order)."""
    from text2speech_b200 import _lib
    calls = []

    def call(name, *args):
        calls.append(name)
        with torch.no_grad() if name != "wgb_upsample_wgrad" else torch.enable_grad():
            TABLE[name](*args)

    monkeypatch.setattr(_lib, "call", call)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: 0)
    monkeypatch.setattr(_lib, "require_b200", lambda device: None)
    monkeypatch.setattr(torch.cuda, "device", lambda device: contextlib.nullcontext())     # `with torch.cuda.device(cpu)`
    return calls
