"""Kernel sequencing for the WaveGlow flow: which C-ABI entry point runs when, on which buffers.

HBM layout (all channels-last, one allocation per call from torch's caching allocator):
  cond      [B, T, 640]  bf16|fp32   upsampled + regrouped mel, read by all 96 layers
  x         [B, T, 8]    fp32        flow state; flow k touches the last C_k channels; final audio layout
  h0, h1    [B, T, 512]  bf16        residual stream ping-pong (neighbour tiles read h with a halo)
  acts_all  [8, B, T, 512] bf16      gated activations of every layer, K operand of the skip GEMM
No CPU path: every step is a call into libwaveglow_b200.so.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch

from . import _lib
from .packing import PackedWaveGlow

Tensor = torch.Tensor

# Which tcgen05 kernels the BF16 path uses: "pair" = CTA pairs (cta_group::2, UMMA M = 256), "single" = one CTA per tile.
GATE_KERNEL = os.environ.get("WGB_GATE_KERNEL", "pair")
RES_KERNEL = os.environ.get("WGB_RES_KERNEL", "pair")
# skip path (WN.end composed with the skip GEMM in all but "pair" / "single"):
#   "skip16" (default) one N = 16 tcgen05 sweep over the stored activations of all layers (HBM-bound, 8 KB per group step)
#   "res16"  the residual kernel adds each layer's [16 x 512] product into a per-row accumulator while the activations
#            are on chip (third, N = 16 pass); the last layer + coupling run in wgb_tc_wn_skip16_end.  One activation
#            buffer instead of eight (-12.6 GB at batch 64), but measured 2 % SLOWER: every extra accumulator pass pays
#            the TMEM drain and pipeline hand-off of a full pass (+0.14 ms per residual launch vs -1.2 ms per flow).
#   "acc"    the product accumulated in the gate kernel's epilogue (composed-conditioning path only; also slower)
#   "pair" / "single"  the K = 4096 x N = 512 GEMM with WN.end in the epilogue
SKIP_KERNEL = os.environ.get("WGB_SKIP_KERNEL", "skip16")
# composed-conditioning path: fold WN.start into in_layers[0] (first layer reads the stacked flow state, K = 64 + 320)
FOLD_START = os.environ.get("WGB_FOLD_START", "1") == "1"
# infer: run WN.start of flow k-1 inside the skip+end kernel of flow k (one launch and one pass over x less per flow)
FUSE_START = os.environ.get("WGB_FUSE_START", "1") == "1"


def upsample_cond(pk: PackedWaveGlow, mel: Tensor) -> Tensor:
    """mel [B, n_mel, F] fp32 -> cond [B, 32F, n_mel*n_group] (bf16 or fp32): ConvTranspose1d +
    trim + regroup of the reference (glow.py:252-258 / :213-221) as one GEMM."""
    b, n_mel, f = mel.shape
    s = _lib.stream_ptr()
    bf16 = pk.mode == "bf16"
    k = pk.up_taps * pk.up_ld_tap
    n_cols = pk.w_up.shape[0]
    if bf16:            # tcgen05 GEMM: bf16 im2col rows [B, F, 4*128] x packed weight [20480, 512] -> bf16 cond
        a = torch.empty((b, f, k), device=mel.device, dtype=torch.bfloat16)
        _lib.call("wgb_upsample_im2col", mel, a, 1, b, n_mel, f, pk.up_taps, pk.up_ld_tap, s)
        cond = torch.empty((b, f, n_cols), device=mel.device, dtype=torch.bfloat16)
        _lib.call("wgb_tc_gemm", a, pk.w_up, pk.b_up, cond, 1, b, f, n_cols, k, s)
    else:
        a = torch.empty((b, f, k), device=mel.device, dtype=torch.float32)
        _lib.call("wgb_upsample_im2col", mel, a, 0, b, n_mel, f, pk.up_taps, pk.up_ld_tap, s)
        cond = torch.empty((b, f, n_cols), device=mel.device, dtype=torch.float32)
        _lib.call("wgb_sgemm_f32", a, pk.w_up, pk.b_up, cond, 0, 1, b * f, n_cols, k, k, 0, k, n_cols, 0, 0, 0, s)
    t_per_frame = pk.up_stride // pk.n_group
    return cond.view(b, f * t_per_frame, n_cols // t_per_frame)


GUARD_FRAMES = 4          # >= max dilation (128 group steps) / 32 steps per frame


def mel_stack(pk: PackedWaveGlow, mel: Tensor, frames_pad: Optional[int] = None) -> Tensor:
    """mel [B, n_mel, F] fp32 -> bf16 [B, frames_pad, 4*n_mel]: the four frames feeding each frame's group steps (the A
    operand of the upsample GEMM without channel padding), K operand of the composed conditioning in
    wgb_tc2_wn_gate_mel.  Rows >= F (guard frames of the padded layout) are never used."""
    b, n_mel, f = mel.shape
    fp = f if frames_pad is None else frames_pad
    if fp != f:
        padded = torch.zeros((b, n_mel, fp), device=mel.device, dtype=torch.float32)
        padded[:, :, :f] = mel
        mel = padded
    a = torch.empty((b, fp, pk.up_taps * n_mel), device=mel.device, dtype=torch.bfloat16)
    _lib.call("wgb_upsample_im2col", mel, a, 1, b, n_mel, fp, pk.up_taps, n_mel, _lib.stream_ptr())
    return a


def use_mel_path(pk: PackedWaveGlow, batch: int, frames: int, t: int) -> bool:
    """Composed conditioning (K = 1856, rows tiled 128 frames x 1 phase over the padded frame axis of the whole batch)
    vs the cond tensor (K = 2176, rows tiled 128 group steps per utterance): pick the one that executes fewer MMA
    chunks for this shape ('auto'), or as forced by model.cond_path."""
    if pk.mode != "bf16" or not pk.has_mel or pk.cond_path == "cond" or GATE_KERNEL != "pair" or RES_KERNEL != "pair":
        return False
    if t > frames * 32:
        return False
    if pk.cond_path == "mel":
        return True
    used = -(-t // 32)                        # forward() may use fewer (and a partial last) frame than the mel has
    return -(-batch * (used + GUARD_FRAMES) // 128) * 32 * 29 < batch * -(-t // 128) * 34


def _wn_bf16(pk: PackedWaveGlow, fl: dict, x: Tensor, cond, bufs, direction: int, log_s: Optional[Tensor],
             start_done: bool = False, next_fl: Optional[dict] = None) -> bool:
    """One WN + coupling.  start_done: h0 already holds WN.start(x) (written by the previous flow's skip+end kernel);
    next_fl: the flow that runs next in infer order, whose WN.start the pair skip+end kernel fuses behind the
    coupling.  Returns True when that fused start was issued."""
    b, t = x.shape[0], x.shape[1]
    s = _lib.stream_ptr()
    h0, h1, acts_all = bufs["h0"], bufs["h1"], bufs["acts"]
    skip_acc = bufs.get("skip_acc")           # "acc" skip path: [4, B*T, 8] fp32, one acts buffer
    skip_row = bufs.get("skip_row")           # "res16" skip path: [B*T, 8] fp32, one acts buffer
    h_rows = h0.shape[1]                      # row pitch per utterance: t, or 32 * frames_pad in the padded layout
    gate = "wgb_tc2_wn_gate" if GATE_KERNEL == "pair" else "wgb_tc_wn_gate"
    res = "wgb_tc2_wn_res" if RES_KERNEL == "pair" else "wgb_tc_wn_res"
    skip_end = "wgb_tc2_wn_skip_end" if SKIP_KERNEL == "pair" else "wgb_tc_wn_skip_end"
    if not start_done:
        _lib.call("wgb_wn_start_padded", x, fl["w_start"], fl["b_start"], h0, 1, b, t, h_rows, pk.n_ch, fl["n_half"], s)
    cur, nxt = h0, h1
    for i in range(pk.n_layers):
        acts = acts_all[i if (skip_acc is None and skip_row is None) else 0]
        if i == 0 and isinstance(cond, tuple) and FOLD_START and "w_gate0" in fl:
            # WN.start folded into in_layers[0]: the operand is the stacked flow state, not h0 (h0 still feeds the
            # residual GEMM)
            xs = bufs.get("x_stack")
            if xs is None:
                xs = bufs["x_stack"] = torch.empty((b, h_rows, 64), device=x.device, dtype=torch.bfloat16)
            _lib.call("wgb_x_stack", x, xs, b, t, h_rows, fl["n_half"], s)
            _lib.call("wgb_tc2_wn_gate_mel0", xs, cond[1], fl["w_gate0"], fl["w_mel"][0], fl["b_mel"][0], acts, b, t,
                      h_rows // 32, fl["w_comp"][0] if skip_acc is not None else None, skip_acc, 1, s)
        elif isinstance(cond, tuple):         # ("mel", mel_stack): conditioning composed with the upsampler
            _lib.call("wgb_tc2_wn_gate_mel", cur, cond[1], fl["w_gate"][i], fl["w_mel"][i], fl["b_mel"][i], acts,
                      b, t, h_rows // 32, 2 ** i, fl["w_comp"][i] if skip_acc is not None else None, skip_acc,
                      int(i == 0), s)
        else:
            _lib.call(gate, cur, cond, fl["w_gate"][i], fl["b_gate"][i], acts, b, t, 2 ** i, s)
        if i < pk.n_layers - 1:
            if RES_KERNEL == "pair":
                _lib.call(res, acts, fl["w_res"][i], fl["b_res"][i], cur, nxt, b, t, h_rows,
                          fl["w_skip16_layers"][i] if skip_row is not None else None, skip_row, int(i == 0), s)
            else:
                _lib.call(res, acts, fl["w_res"][i], fl["b_res"][i], cur, nxt, b, t, s)
            cur, nxt = nxt, cur
    args = (acts_all, pk.n_layers, fl["w_skip"], fl["w_end_t"], fl["b_end"], x,
            fl["w_mix_inv"] if direction == 0 else None, log_s, b, t, fl["n_half"], direction)
    if SKIP_KERNEL == "single":
        _lib.call(skip_end, *args, s)
        return False
    # infer: the next flow (k-1) starts from the row as updated here; forward: flow k+1 first applies its 1x1 conv
    fuse = next_fl is not None and FUSE_START and (direction == 0 or skip_acc is None)
    if skip_acc is not None:
        _lib.call("wgb_end_from_acc", skip_acc, fl["b_end"], x, fl["w_mix_inv"] if direction == 0 else None, log_s, b, t,
                  fl["n_half"], direction, *((next_fl["w_start"], next_fl["b_start"], next_fl["n_half"], h0, h_rows)
                                             if fuse else (None, None, 0, None, 0)), s)
        return fuse
    tail = ()
    if skip_row is not None:                  # layers 0..L-2 are in skip_row; the last layer's activations are in acts
        skip_end = "wgb_tc_wn_skip16_end"
        args = (acts_all[0], 1, fl["w_skip16_layers"][pk.n_layers - 1], fl["b_end"], x,
                fl["w_mix_inv"] if direction == 0 else None, log_s, b, t, fl["n_half"], direction)
        tail = (skip_row,)
    elif SKIP_KERNEL in ("skip16", "acc", "res16"):
        skip_end = "wgb_tc_wn_skip16_end"
        args = (acts_all, pk.n_layers, fl["w_skip16"], fl["b_end"], x,
                fl["w_mix_inv"] if direction == 0 else None, log_s, b, t, fl["n_half"], direction)
        tail = (None,)
    if SKIP_KERNEL == "pair":                 # K = 4096 x N = 512 variant: fused start for infer only, no fused mix
        fuse = fuse and direction == 0
    elif tail:
        tail = tail + ((next_fl["w_mix"] if (fuse and direction == 1) else None),)
    if fuse:
        _lib.call(skip_end, *args, next_fl["w_start"], next_fl["b_start"], next_fl["n_half"], h0, h_rows, *tail, s)
    else:
        _lib.call(skip_end, *args, None, None, 0, None, 0, *tail, s)
    return fuse


def _wn_fp32(pk: PackedWaveGlow, fl: dict, x: Tensor, cond: Tensor, bufs, direction: int, log_s: Optional[Tensor]):
    """Reference op order, one CUDA-core kernel per reference op (glow.py:154-175)."""
    b, t = x.shape[0], x.shape[1]
    s = _lib.stream_ptr()
    h, u, acts, rs, skip = bufs
    c, n_cond = pk.n_ch, cond.shape[2]
    _lib.call("wgb_wn_start", x, fl["w_start"], fl["b_start"], h, 0, b * t, c, fl["n_half"], s)
    for i in range(pk.n_layers):
        d = 2 ** i
        w_in = fl["w_in"][i]
        taps = w_in.shape[0]
        for j in range(taps):
            shift = (j - (taps - 1) // 2) * d
            _lib.call("wgb_sgemm_f32", h, w_in[j], fl["b_in"][i] if j == 0 else None, u, 0, b, t, 2 * c, c,
                      c, t * c, c, 2 * c, t * 2 * c, shift, int(j > 0), s)
        _lib.call("wgb_sgemm_f32", cond, fl["w_cond"][i], None, u, 0, b, t, 2 * c, n_cond,
                  n_cond, t * n_cond, n_cond, 2 * c, t * 2 * c, 0, 1, s)
        _lib.call("wgb_gate_f32", u, acts, b * t, c, s)
        n_out = fl["w_rs"][i].shape[0]
        _lib.call("wgb_sgemm_f32", acts, fl["w_rs"][i], fl["b_rs"][i], rs, 0, 1, b * t, n_out, c,
                  c, 0, c, n_out, 0, 0, 0, s)
        _lib.call("wgb_res_skip_f32", rs, h, skip, b * t, c, int(n_out == 2 * c), int(i == 0), s)
    _lib.call("wgb_end_coupling_f32", skip, fl["w_end_t"], fl["b_end"], x,
              fl["w_mix_inv"] if direction == 0 else None, log_s, b, t, c, fl["n_half"], direction, s)


def _alloc(pk: PackedWaveGlow, b: int, t: int, device, h_rows: Optional[int] = None):
    if pk.mode == "bf16":
        bf = torch.bfloat16
        acc = h_rows is not None and SKIP_KERNEL == "acc"     # composed-conditioning path: skip accumulated by the gate kernel
        if SKIP_KERNEL == "res16" and RES_KERNEL == "pair" and pk.n_layers > 1:
            hr = t if h_rows is None else h_rows
            mk = torch.zeros if hr != t else torch.empty
            return {"h0": mk((b, hr, pk.n_ch), device=device, dtype=bf), "h1": mk((b, hr, pk.n_ch), device=device, dtype=bf),
                    "acts": torch.empty((1, b, t, pk.n_ch), device=device, dtype=bf),
                    "skip_row": torch.empty((b * t, 8), device=device, dtype=torch.float32)}
        if h_rows is not None and h_rows != t:     # padded layout: guard rows are zero and no kernel ever writes them
            h0 = torch.zeros((b, h_rows, pk.n_ch), device=device, dtype=bf)
            h1 = torch.zeros((b, h_rows, pk.n_ch), device=device, dtype=bf)
        else:
            h0 = torch.empty((b, t, pk.n_ch), device=device, dtype=bf)
            h1 = torch.empty((b, t, pk.n_ch), device=device, dtype=bf)
        if acc:
            return {"h0": h0, "h1": h1, "acts": torch.empty((1, b, t, pk.n_ch), device=device, dtype=bf),
                    "skip_acc": torch.empty((4, b * t, 8), device=device, dtype=torch.float32)}
        return {"h0": h0, "h1": h1, "acts": torch.empty((pk.n_layers, b, t, pk.n_ch), device=device, dtype=bf)}
    f32 = torch.float32
    return (torch.empty((b, t, pk.n_ch), device=device, dtype=f32),
            torch.empty((b, t, 2 * pk.n_ch), device=device, dtype=f32),
            torch.empty((b, t, pk.n_ch), device=device, dtype=f32),
            torch.empty((b, t, 2 * pk.n_ch), device=device, dtype=f32),
            torch.empty((b, t, pk.n_ch), device=device, dtype=f32))


def run_wn(pk: PackedWaveGlow, k: int, x: Tensor, cond: Tensor, bufs, direction: int, log_s: Optional[Tensor],
           start_done: bool = False, next_k: Optional[int] = None) -> bool:
    if pk.mode == "bf16":
        return _wn_bf16(pk, pk.flows[k], x, cond, bufs, direction, log_s, start_done,
                        pk.flows[next_k] if next_k is not None else None)
    _wn_fp32(pk, pk.flows[k], x, cond, bufs, direction, log_s)
    return False


def infer(pk: PackedWaveGlow, mel: Tensor, z: Tensor, sigma: float) -> Tensor:
    """WaveGlow.infer (glow.py:251-292): mel [B,80,F] fp32, z [B,8,32F] fp32 -> audio [B, 256F] fp32."""
    b, _, f = mel.shape
    t = f * pk.up_stride // pk.n_group
    s = _lib.stream_ptr()
    h_rows = None
    if use_mel_path(pk, b, f, t):
        h_rows = 32 * (f + GUARD_FRAMES)
        cond = ("mel", mel_stack(pk, mel, f + GUARD_FRAMES))
    else:
        cond = upsample_cond(pk, mel)
    x = torch.empty((b, t, pk.n_group), device=mel.device, dtype=torch.float32)
    _lib.call("wgb_flow_from_z", z, x, b, t, float(sigma), s)
    bufs = _alloc(pk, b, t, mel.device, h_rows)
    start_done = False
    for k in reversed(range(pk.n_flows)):
        start_done = run_wn(pk, k, x, cond, bufs, 0, None, start_done, k - 1 if k > 0 else None)
    return x.view(b, t * pk.n_group)


def forward(pk: PackedWaveGlow, mel: Tensor, audio: Tensor) -> Tuple[Tensor, List[Tensor], List[Tensor]]:
    """WaveGlow.forward (glow.py:207-249): (mel [B,80,F], audio [B,N]) -> (z, log_s_list, log_det_W_list)."""
    b, _, f = mel.shape
    n = audio.shape[1]
    up_len = (f - 1) * pk.up_stride + pk.up_stride * pk.up_taps
    assert up_len >= n, "upsampled spectrogram shorter than audio"            # glow.py:216
    t = n // pk.n_group
    s = _lib.stream_ptr()
    if t > f * pk.up_stride // pk.n_group:
        raise RuntimeError("audio longer than 256 * frames is not supported by the regrouped upsample GEMM")
    h_rows = None
    if use_mel_path(pk, b, f, t):
        used = -(-t // 32)                    # frames that carry audio; the mel may have more (glow.py:216-218)
        h_rows = 32 * (used + GUARD_FRAMES)
        cond = ("mel", mel_stack(pk, mel[:, :, :used].contiguous(), used + GUARD_FRAMES))
    else:
        cond = upsample_cond(pk, mel)
        if cond.shape[1] > t:
            cond = cond[:, :t].contiguous()                                   # glow.py:217-218
    x = audio[:, : t * pk.n_group].reshape(b, t, pk.n_group).float().contiguous().clone()
    bufs = _alloc(pk, b, t, mel.device, h_rows)
    log_s_list, log_det_list = [], []
    started = False                           # flow k's 1x1 conv + WN.start already done by flow k-1's skip+end kernel
    for k in range(pk.n_flows):
        fl = pk.flows[k]
        if not started:
            _lib.call("wgb_flow_mix", x, fl["w_mix"], b * t, 2 * fl["n_half"], s)
        log_det_list.append(torch.tensor(b * t * fl["logdet"], device=mel.device, dtype=torch.float32))
        log_s = torch.empty((b, fl["n_half"], t), device=mel.device, dtype=torch.float32)
        started = run_wn(pk, k, x, cond, bufs, 1, log_s, started, k + 1 if k + 1 < pk.n_flows else None)
        log_s_list.append(log_s)
    z = torch.empty((b, pk.n_group, t), device=mel.device, dtype=torch.float32)
    _lib.call("wgb_flow_to_z", x, z, b, t, s)
    return z, log_s_list, log_det_list


def wn_standalone(pk: PackedWaveGlow, k: int, audio_0: Tensor, spect: Tensor) -> Tensor:
    """WN.forward((audio_0 [B,n_half,T], spect [B,640,T])) -> [B, 2*n_half, T] (glow.py:154-175).
    Runs the coupling in forward direction on a zero a1, which returns (b, log_s) = WN output."""
    b, n_half, t = audio_0.shape
    s = _lib.stream_ptr()
    x = torch.zeros((b, t, 8), device=audio_0.device, dtype=torch.float32)
    x[:, :, 8 - 2 * n_half: 8 - n_half] = audio_0.permute(0, 2, 1)
    cdt = torch.bfloat16 if pk.mode == "bf16" else torch.float32
    cond = spect.permute(0, 2, 1).contiguous().to(cdt)
    log_s = torch.empty((b, n_half, t), device=audio_0.device, dtype=torch.float32)
    run_wn(pk, k, x, cond, _alloc(pk, b, t, audio_0.device), 1, log_s)
    return torch.cat([x[:, :, 8 - n_half:].permute(0, 2, 1), log_s], dim=1)
