"""Host-side weight packing: reference state_dict layout -> the layouts the kernels consume.

Pure tensor reshuffling (runs on CPU, unit-tested without a GPU).  Source shapes are the
reference's: in_layers [1024,512,3], cond_layers [1024,640,1], res_skip_layers
[1024|512,512,1], start [512,n_half,1], end [2*n_half,512,1], convinv [C,C,1], upsample [80,80,1024].

BF16 tensor-core layouts
  gate GEMM   w [1024, 2176] bf16, K index = tap*512 + c_in for the three dilated taps, then
              1536 + c_cond for the conditioning 1x1.  Row order: pass p in 0..3 holds
              tanh rows 128p..128p+127 followed by sigmoid rows 512+128p..512+128p+127, so one
              256-column accumulator pass contains matching tanh/sigmoid pairs.  bias = b_in + b_cond.
  res GEMM    w [512, 512] bf16 = rows 0..511 of res_skip_layers[i] (i < n_layers-1).
  skip GEMM   w [512, n_layers*512] bf16 = the skip rows of every layer side by side along K
              (rows 512..1023 for i < n_layers-1, rows 0..511 for the last layer); the skip biases are
              folded through WN.end into b_end.
FP32 validation layouts
  in_layers as [3][1024][512] (tap-major), everything else as [out][in].
"""
from __future__ import annotations

from typing import Dict, List

import torch

Tensor = torch.Tensor


def fold_weight_norm(g: Tensor, v: Tensor) -> Tensor:
    """w = g * v / ||v|| per output channel (torch weight_norm dim=0; reference glow.py:294-310)."""
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
    return v * (g / norm)


def folded(state: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """fp32 ``*.weight``-only view of a reference-layout state_dict (either layout accepted)."""
    out = {}
    for key, val in state.items():
        val = val.detach().float().cpu()
        if key.endswith(".weight_g"):
            stem = key[:-2]
            out[stem] = fold_weight_norm(val, state[stem + "_v"].detach().float().cpu())
        elif not key.endswith(".weight_v"):
            out[key] = val
    return out


def gate_row_order(n_ch: int = 512, block: int = 128) -> Tensor:
    """Original in/cond output-row index for each packed row of the gate GEMM."""
    rows = []
    for p in range(n_ch // block):
        rows.append(torch.arange(p * block, (p + 1) * block))
        rows.append(torch.arange(n_ch + p * block, n_ch + (p + 1) * block))
    return torch.cat(rows)


def pack_gate(w_in: Tensor, b_in: Tensor, w_cond: Tensor, b_cond: Tensor):
    """-> (w [2C, taps*C + n_cond] fp32 (cast to bf16 by the caller), bias [2C] fp32)."""
    two_c, c, taps = w_in.shape
    order = gate_row_order(c)
    w = torch.cat([w_in.permute(0, 2, 1).reshape(two_c, taps * c), w_cond[:, :, 0]], dim=1)
    return w[order].contiguous(), (b_in + b_cond)[order].contiguous()


def pack_gate0(w_in0: Tensor, w_start: Tensor, b_start: Tensor) -> Tensor:
    """WN.start folded into in_layers[0] (glow.py:156 + :160; both linear, dilation 1):
        in_layers[0](start(a0))[t] = sum_tap (W_in0[:,:,tap] W_start) a0[t+tap-1] + (W_in0[:,:,tap] b_start) [t+tap-1 in range]
    -> bf16 [2C, 64] in the gate GEMM's packed row order, columns matching the rows written by wgb_x_stack:
    [0..11] hi(W_tap W_start) (meets hi(a0)), [12..23] the same (meets lo(a0)), [24..35] lo(W_tap W_start) (meets hi(a0)),
    [36..38] hi(W_tap b_start), [39..41] lo(W_tap b_start) (meet the in-range indicators), rest zero."""
    two_c, c, taps = w_in0.shape
    assert taps == 3, "the fold is written for kernel_size 3"
    n_half = w_start.shape[1]
    w = torch.zeros(two_c, 64, dtype=torch.float32)
    for tap in range(taps):
        comp = (w_in0[:, :, tap].double() @ w_start[:, :n_half].double()).float()
        wb = (w_in0[:, :, tap].double() @ b_start.double()).float()
        hi = comp.bfloat16().float()
        lo = (comp - hi).bfloat16().float()
        w[:, tap * 4: tap * 4 + n_half] = hi
        w[:, 12 + tap * 4: 12 + tap * 4 + n_half] = hi
        w[:, 24 + tap * 4: 24 + tap * 4 + n_half] = lo
        wbh = wb.bfloat16().float()
        w[:, 36 + tap] = wbh
        w[:, 39 + tap] = (wb - wbh).bfloat16().float()
    return w[gate_row_order(c)].bfloat16().contiguous()


def pack_skip(w_rs: List[Tensor], b_rs: List[Tensor], n_ch: int):
    """-> (w_skip [C, L*C], b_skip_total [C]) from the per-layer res_skip weights."""
    cols, bias = [], torch.zeros(n_ch, dtype=torch.float64)
    last = len(w_rs) - 1
    for i, (w, b) in enumerate(zip(w_rs, b_rs)):
        lo = 0 if i == last else n_ch
        cols.append(w[lo: lo + n_ch, :, 0])
        bias += b[lo: lo + n_ch].double()
    return torch.cat(cols, dim=1).contiguous(), bias


def pack_end(w_end: Tensor, b_end: Tensor, b_skip_total: Tensor):
    """-> (w_end_t [C, 8] fp32 zero padded, b_end_plain [8], b_end_folded [8] = b_end + W_end b_skip)."""
    rows, n_ch = w_end.shape[0], w_end.shape[1]
    w_t = torch.zeros(n_ch, 8, dtype=torch.float32)
    w_t[:, :rows] = w_end[:, :, 0].t()
    plain = torch.zeros(8, dtype=torch.float32)
    plain[:rows] = b_end
    fold = torch.zeros(8, dtype=torch.float64)
    fold[:rows] = b_end.double() + w_end[:, :, 0].double() @ b_skip_total
    return w_t.contiguous(), plain, fold.float()


def pack_skip_end16(w_skip: Tensor, w_end: Tensor):
    """Compose WN.end with the skip GEMM (both linear, glow.py:167-175): W_end [2n_half,512] @ W_skip [512, L*512]
    -> [8, L*512] (zero rows beyond 2*n_half), computed in fp64 and split into bf16 hi / lo parts stacked as
    [16, L*512] (rows 0..7 hi, 8..15 lo): the N = 16 operand of wgb_tc_wn_skip16_end."""
    rows = w_end.shape[0]
    comp = torch.zeros(8, w_skip.shape[1], dtype=torch.float64)
    comp[:rows] = w_end[:, :, 0].double() @ w_skip.double()
    comp = comp.float()
    hi = comp.bfloat16()
    lo = (comp - hi.float()).bfloat16()
    return torch.cat([hi, lo], dim=0).contiguous()


def pack_skip_end_layers(w_skip: Tensor, w_end: Tensor, n_ch: int):
    """Per-layer fp32 [L][n_ch][8] = (W_end W_skip_i)^T (zero columns beyond 2*n_half): the weights the gate kernel's
    epilogue applies to its fp32 activations when it accumulates the skip path (wgb_tc2_wn_gate_mel)."""
    rows = w_end.shape[0]
    comp = torch.zeros(8, w_skip.shape[1], dtype=torch.float64)
    comp[:rows] = w_end[:, :, 0].double() @ w_skip.double()
    layers = w_skip.shape[1] // n_ch
    return comp.float().reshape(8, layers, n_ch).permute(1, 2, 0).contiguous()


# Parameter generation: bumped by every in-place weight update that torch's version counters cannot see (FusedAdam's
# raw kernel over its flat buffer, CUDA-graph replays of a training step).  WaveGlow._signature() includes it, so packed
# weights are rebuilt after training steps instead of silently going stale.
_param_generation = [0]


def bump_param_generation() -> int:
    _param_generation[0] += 1
    return _param_generation[0]


def param_generation() -> int:
    return _param_generation[0]


def pack_mix(w: Tensor):
    """convinv weight [C,C,1] -> (W [8,8] fp32, W^-1 [8,8] fp32 via fp64, log det W python float).
    The reference inverts in fp32 and caches (glow.py:88-95); fp64 here is strictly more accurate.  log det W is
    torch.logdet's (glow.py:100): NaN when det W < 0."""
    c = w.shape[0]
    w2 = w[:, :, 0].double()
    fwd = torch.zeros(8, 8, dtype=torch.float32)
    inv = torch.zeros(8, 8, dtype=torch.float32)
    fwd[:c, :c] = w2.float()
    inv[:c, :c] = torch.linalg.inv(w2).float()
    sign, logabs = torch.linalg.slogdet(w2)
    return fwd, inv, float(logabs) if float(sign) > 0 else float("nan")


def pack_upsample(w: Tensor, b: Tensor, n_group: int, ld_tap: int):
    """ConvTranspose1d weight [c_in, c_out, K] (stride = K/4) -> GEMM weight
    [ (K/4/n_group) * c_out * n_group, taps * ld_tap ] whose output row q, column tt*(c_out*n_group)+m*n_group+g is
    sample 256q + 8tt + g of channel m — i.e. the regrouped cond layout of glow.py:257-258."""
    c_in, c_out, k = w.shape
    taps = 4
    stride = k // taps
    tt = stride // n_group
    w5 = w.reshape(c_in, c_out, taps, tt, n_group)               # k = j*stride + tt*n_group + g
    packed = torch.zeros(tt, c_out, n_group, taps, ld_tap, dtype=torch.float32, device=w.device)
    packed[..., :c_in] = w5.permute(3, 1, 4, 2, 0)
    bias_col = b.reshape(1, c_out, 1).expand(tt, c_out, n_group).reshape(-1).contiguous()
    return packed.reshape(tt * c_out * n_group, taps * ld_tap).contiguous(), bias_col.float()


def upsample_phase_operand(w_up: Tensor, b_up: Tensor, n_group: int):
    """(U^T [32, 4*n_mel, n_mel*n_group] fp32, bias column [n_mel*n_group] fp32): U_phase maps the stacked mel frames
    (mel[f], mel[f-1], mel[f-2], mel[f-3]) to the regrouped upsampled mel at group step 32 f + phase
    (glow.py:183-185,252-258); it is the phase's slice of ``pack_upsample``."""
    c_in = w_up.shape[0]
    u, b_col = pack_upsample(w_up, b_up, n_group, c_in)          # [32 * 640, 4 * 80], [32 * 640]
    n_cond = w_up.shape[1] * n_group
    phases = u.shape[0] // n_cond
    return u.reshape(phases, n_cond, u.shape[1]).transpose(1, 2).contiguous(), b_col[:n_cond].contiguous()


def pack_cond_mel(w_cond_rows: Tensor, w_up: Tensor, b_up: Tensor, n_group: int, device=None, u_t: Tensor = None,
                  b_col: Tensor = None):
    """Compose cond_layers[i] with WaveGlow.upsample (glow.py:183-185,252-258 + :141-143,161).

    For group step t = 32 f + phase, the regrouped upsampled mel is cond[t] = U_phase . stack(mel[f], mel[f-1],
    mel[f-2], mel[f-3]) + b_up, so W_cond cond[t] = (W_cond U_phase) . stack + W_cond b_up.  ``w_cond_rows`` [2C, 640]
    is W_cond in whatever row order the caller wants (the gate GEMM's packed order).  Returns
    (V [32, 2C, 320] fp32, bias_add [2C] fp32).  On a CUDA device the 32 products run through this library's own FP32
    GEMM (wgb_sgemm_f32); on the CPU (unit tests) in torch fp64.  ``u_t`` / ``b_col`` = a cached
    ``upsample_phase_operand`` (on ``device``)."""
    dev = torch.device(device) if device is not None else w_cond_rows.device
    if u_t is None:
        u_t, b_col = upsample_phase_operand(w_up, b_up, n_group)
        u_t, b_col = u_t.to(dev), b_col.to(dev)
    phases, k_mel, n_cond = u_t.shape
    rows = w_cond_rows.shape[0]
    if dev.type != "cuda":
        w64 = w_cond_rows.to(dev, torch.float64)
        v = torch.matmul(w64[None], u_t.to(torch.float64).transpose(1, 2))
        return v.float(), (w64 @ b_col.to(torch.float64)).float()
    from . import _lib
    wc = w_cond_rows.to(dev, torch.float32).contiguous()
    v_t = torch.empty((phases, k_mel, rows), device=dev, dtype=torch.float32)    # V^T: [phase][k][row]
    with torch.cuda.device(dev):
        # C[phase][m = k][n = row] = sum_c U^T[phase][k][c] W_cond[row][c]
        _lib.call("wgb_sgemm_f32", u_t, wc, None, v_t, 0, phases, k_mel, rows, n_cond, n_cond, k_mel * n_cond, n_cond,
                  rows, k_mel * rows, 0, 0, _lib.stream_ptr())
    bias_add = (w_cond_rows.double() @ b_col.cpu().double()).float().to(dev)
    return v_t.transpose(1, 2).contiguous(), bias_add


class PackedWaveGlow:
    """All device-side constants of one WaveGlow, for one numeric mode ('bf16' or 'fp32')."""

    def __init__(self, state: Dict[str, Tensor], n_flows: int, n_layers: int, n_ch: int, n_group: int, mode: str,
                 device: torch.device, compose_cond=None):
        assert mode in ("bf16", "fp32")
        # bf16 mode on a GPU also packs the conditioning path composed with the upsampler (pack_cond_mel):
        # +2 GB of per-phase weights per model, 15 % fewer gate-GEMM FLOPs when the engine picks that path
        self.has_mel = (mode == "bf16" and torch.device(device).type == "cuda") if compose_cond is None else bool(compose_cond)
        self.cond_path = "auto"
        st = folded(state)
        self.mode, self.n_flows, self.n_layers, self.n_ch, self.n_group = mode, n_flows, n_layers, n_ch, n_group
        dev = device
        bf = torch.bfloat16
        self.flows = []
        if self.has_mel:
            u_t, b_col = upsample_phase_operand(st["upsample.weight"], st["upsample.bias"], n_group)
            u_t, b_col = u_t.to(dev), b_col.to(dev)
        for k in range(n_flows):
            p = f"WN.{k}."
            f: Dict[str, object] = {}
            w_rs = [st[p + f"res_skip_layers.{i}.weight"] for i in range(n_layers)]
            b_rs = [st[p + f"res_skip_layers.{i}.bias"] for i in range(n_layers)]
            w_end, b_end = st[p + "end.weight"], st[p + "end.bias"]
            f["n_half"] = w_end.shape[0] // 2
            f["w_start"] = st[p + "start.weight"][:, :, 0].contiguous().to(dev)
            f["b_start"] = st[p + "start.bias"].contiguous().to(dev)
            w_skip, b_skip = pack_skip(w_rs, b_rs, n_ch)
            w_end_t, b_plain, b_fold = pack_end(w_end, b_end, b_skip)
            f["w_end_t"] = w_end_t.to(dev)
            fwd, inv, logdet = pack_mix(st[f"convinv.{k}.conv.weight"])
            f["w_mix"], f["w_mix_inv"], f["logdet"] = fwd.to(dev), inv.to(dev), logdet
            if mode == "bf16":
                f["b_end"] = b_fold.to(dev)
                f["w_skip"] = w_skip.to(dev, bf)
                w16 = pack_skip_end16(w_skip, w_end)
                f["w_skip16"] = w16.to(dev)
                # the same product per layer, [16][n_ch] contiguous: wgb_tc2_wn_res accumulates layers 0..L-2 while the
                # activations are on chip, wgb_tc_wn_skip16_end adds the last layer
                f["w_skip16_layers"] = [w16[:, i * n_ch:(i + 1) * n_ch].contiguous().to(dev) for i in range(n_layers)]
                f["w_comp"] = pack_skip_end_layers(w_skip, w_end, n_ch).to(dev)
                if st[p + "in_layers.0.weight"].shape[2] == 3:
                    f["w_gate0"] = pack_gate0(st[p + "in_layers.0.weight"], st[p + "start.weight"][:, :, 0],
                                              st[p + "start.bias"]).to(dev)
                f["w_gate"], f["b_gate"], f["w_res"], f["b_res"], f["w_mel"], f["b_mel"] = [], [], [], [], [], []
                for i in range(n_layers):
                    wg, bg = pack_gate(st[p + f"in_layers.{i}.weight"], st[p + f"in_layers.{i}.bias"],
                                       st[p + f"cond_layers.{i}.weight"], st[p + f"cond_layers.{i}.bias"])
                    f["w_gate"].append(wg.to(dev, bf))
                    f["b_gate"].append(bg.to(dev))
                    if self.has_mel:
                        taps_k = st[p + f"in_layers.{i}.weight"].shape[1] * st[p + f"in_layers.{i}.weight"].shape[2]
                        v, b_add = pack_cond_mel(wg[:, taps_k:].contiguous(), st["upsample.weight"], st["upsample.bias"],
                                                 n_group, dev, u_t, b_col)
                        f["w_mel"].append(v.to(bf).contiguous())
                        f["b_mel"].append((bg.to(dev) + b_add).contiguous())
                    if i < n_layers - 1:
                        f["w_res"].append(w_rs[i][:n_ch, :, 0].contiguous().to(dev, bf))
                        f["b_res"].append(b_rs[i][:n_ch].contiguous().to(dev))
            else:
                f["b_end"] = b_plain.to(dev)
                f["w_in"] = [st[p + f"in_layers.{i}.weight"].permute(2, 0, 1).contiguous().to(dev) for i in range(n_layers)]
                f["b_in"] = [(st[p + f"in_layers.{i}.bias"] + st[p + f"cond_layers.{i}.bias"]).to(dev) for i in range(n_layers)]
                f["w_cond"] = [st[p + f"cond_layers.{i}.weight"][:, :, 0].contiguous().to(dev) for i in range(n_layers)]
                f["w_rs"] = [w[:, :, 0].contiguous().to(dev) for w in w_rs]
                f["b_rs"] = [b.contiguous().to(dev) for b in b_rs]
            self.flows.append(f)
        up_w, up_b = st["upsample.weight"], st["upsample.bias"]
        self.n_mel = up_w.shape[0]
        self.up_taps = 4
        self.up_stride = up_w.shape[2] // self.up_taps
        # fp32 mode: CUDA-core SGEMM, K = 4 taps x n_mel.  bf16 mode: tcgen05 GEMM, each tap padded to a
        # multiple of 64 channels (one 128 B swizzle row per K chunk) -> K = 4 x 128 for 80 mels.
        self.up_ld_tap = ((self.n_mel + 63) // 64) * 64 if mode == "bf16" else ((self.n_mel + 3) // 4) * 4
        w_up, b_up = pack_upsample(up_w, up_b, n_group, self.up_ld_tap)
        self.w_up = w_up.to(dev, bf) if mode == "bf16" else w_up.to(dev)
        self.b_up = b_up.to(dev)
