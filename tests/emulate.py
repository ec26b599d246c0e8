"""Torch-on-CPU emulation of what the CUDA kernels compute FROM THE PACKED WEIGHTS (test helper).

It mirrors text2speech_b200/engine.py step by step with the kernels' data layouts (channels-last,
flow state in the last C channels of x[B,T,8], gate rows permuted per pass, one skip GEMM over all
layers, skip biases folded through WN.end) so the host-side packing and the algorithmic
restructuring are validated against the oracle without a GPU.
"""
import torch

from text2speech_b200.packing import PackedWaveGlow


def shift_rows(h, s):
    """out[:, t] = h[:, t + s] with zero fill (the TMA out-of-bounds fill / SGEMM row guard)."""
    out = torch.zeros_like(h)
    t = h.shape[1]
    if s == 0:
        return h.clone()
    if abs(s) >= t:
        return out
    if s > 0:
        out[:, : t - s] = h[:, s:]
    else:
        out[:, -s:] = h[:, : t + s]
    return out


def upsample(pk: PackedWaveGlow, mel):
    b, n_mel, f = mel.shape
    a = torch.zeros(b, f, pk.up_taps, pk.up_ld_tap)
    for j in range(pk.up_taps):
        a[:, j:, j, :n_mel] = mel[:, :, : f - j].permute(0, 2, 1)
    if pk.mode == "bf16":
        a = a.bfloat16().float()
    out = a.reshape(b * f, -1) @ pk.w_up.float().t() + pk.b_up
    tpf = pk.up_stride // pk.n_group
    return out.reshape(b, f * tpf, -1)


def gate_layer(fl, i, h, cond, d, bf16):
    a = torch.cat([shift_rows(h, -d), h, shift_rows(h, d), cond], dim=2)
    w = fl["w_gate"][i].float()
    u = a @ w.t() + fl["b_gate"][i]
    n_pass = u.shape[2] // 256
    acts = []
    for p in range(n_pass):
        blk = u[:, :, p * 256:(p + 1) * 256]
        acts.append(torch.tanh(blk[:, :, :128]) * torch.sigmoid(blk[:, :, 128:]))
    acts = torch.cat(acts, dim=2)
    return acts.bfloat16().float() if bf16 else acts


def wn_bf16(pk, fl, x, cond, direction, round_bf16=True):
    """Returns log_s (forward) or None; updates x in place.  round_bf16 mimics the HBM dtypes."""
    r = (lambda v: v.bfloat16().float()) if round_bf16 else (lambda v: v)
    n_half = fl["n_half"]
    base = 8 - 2 * n_half
    h = r(x[:, :, base: base + n_half] @ fl["w_start"].t() + fl["b_start"])
    acts_all = []
    for i in range(pk.n_layers):
        acts = gate_layer(fl, i, h, cond, 2 ** i, round_bf16)
        acts_all.append(acts)
        if i < pk.n_layers - 1:
            h = r(h + acts @ fl["w_res"][i].float().t() + fl["b_res"][i])
    skip = torch.cat(acts_all, dim=2) @ fl["w_skip"].float().t()
    out = skip @ fl["w_end_t"] + fl["b_end"]
    b_, s_ = out[:, :, :n_half], out[:, :, n_half: 2 * n_half]
    a0 = x[:, :, base: base + n_half]
    a1 = x[:, :, base + n_half:]
    if direction == 0:
        xin = torch.cat([a0, (a1 - b_) * torch.exp(-s_)], dim=2)
        c = 2 * n_half
        x[:, :, base:] = xin @ fl["w_mix_inv"][:c, :c].t()
        return None
    x[:, :, base + n_half:] = torch.exp(s_) * a1 + b_
    return s_.permute(0, 2, 1).contiguous()


def infer(pk: PackedWaveGlow, mel, z, sigma, round_bf16=True):
    cond = upsample(pk, mel)
    if round_bf16:
        cond = cond.bfloat16().float()
    x = sigma * z.permute(0, 2, 1).contiguous()
    for k in reversed(range(pk.n_flows)):
        wn_bf16(pk, pk.flows[k], x, cond, 0, round_bf16)
    return x.reshape(x.shape[0], -1)


def forward(pk: PackedWaveGlow, mel, audio, round_bf16=True):
    b = audio.shape[0]
    t = audio.shape[1] // pk.n_group
    cond = upsample(pk, mel)[:, :t]
    if round_bf16:
        cond = cond.bfloat16().float()
    x = audio[:, : t * pk.n_group].reshape(b, t, pk.n_group).clone()
    log_s, log_det = [], []
    for k in range(pk.n_flows):
        fl = pk.flows[k]
        c = 2 * fl["n_half"]
        x[:, :, 8 - c:] = x[:, :, 8 - c:] @ fl["w_mix"][:c, :c].t()
        log_det.append(b * t * fl["logdet"])
        log_s.append(wn_bf16(pk, fl, x, cond, 1, round_bf16))
    return x.permute(0, 2, 1).contiguous(), log_s, log_det
