"""Training direction of the vocoder (reference waveglow/train.py:55-56,72,116-124).

``model.train(); outputs = model((mel, audio)); loss = criterion(outputs); loss.backward(); optimizer.step()`` works on
the drop-in ``WaveGlow`` exactly as on the reference: ``WaveGlow.forward`` routes here whenever autograd is recording,
and returns tensors wired into one ``torch.autograd.Function`` whose forward and backward are sequences of C-ABI
kernel calls (no torch op touches an activation):

  forward   the inference kernels of engine.py on the cond tensor (tcgen05 gate / residual GEMMs, composed skip + end), the
            gate kernel additionally storing tanh | sigmoid; every layer's h, acts and (tanh | sigmoid) are kept in
            bf16, the flow state before / after each 1x1 conv in fp32  (~2.1 GB per flow at 32 x 16000 samples)
  backward  per flow, last to first: coupling + WN.end (wgb_coupling_bwd), then per layer, last to first:
              g_acts  = [g_h | g_skip] W_rs          wgb_tc2_wn_res_seg (CTA pairs, two operands as K segments, K = 1024)
              g_in    = gate'(g_acts, tanh, sigmoid) wgb_gate_bwd
              g_h     = g_h + conv_in^T(g_in)        wgb_tc2_wn_res_taps (CTA pairs, three shifted taps, K = 3072)
              (per flow) g_cond += [g_in_0 .. g_in_7] W_cond   wgb_tc_gemm_seg (K = 8192, one fp32 accumulate per flow)
              dW_in, dW_cond, dW_res                 wgb_tc_wgrad      (tcgen05, MN-major operands, K = B*T)
              biases, WN.end / skip / start weights  wgb_tc_wgrad against a [rows, 64] hi/lo stack (+ 8 x 512 parameter algebra)
            then WN.start (wgb_start_bwd), the 1x1 conv (wgb_mix_bwd) and finally the upsampler (wgb_upsample_wgrad).

Weight norm (g, v) and log|det W| stay ordinary torch autograd on parameter-sized tensors around the Function, so the
``.grad`` of every reference parameter (weight_g / weight_v / bias / convinv weight / upsample) is filled.
``FusedAdam`` is train.py:79's optimiser as one kernel over a flat buffer; ``allreduce_gradients`` is the flat-bucket
data-parallel reduction of waveglow/distributed.py:90-142.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .packing import bump_param_generation, gate_row_order, pack_upsample

Tensor = torch.Tensor
N_CH, N_COND, N_LAYERS = 512, 640, 8
N_COND_PAD = 768          # g_cond leading dimension: the cond dgrad GEMM writes 256-column passes


# ------------------------------------------------------------------------------------------------ parameters
def _effective(conv: torch.nn.Module) -> Tensor:
    """conv.weight with weight norm applied by autograd-visible torch ops (glow.py:123,138,142,151)."""
    if hasattr(conv, "weight_g"):
        return torch._weight_norm(conv.weight_v, conv.weight_g, 0)
    return conv.weight


def effective_weights(model) -> Tuple[List[str], List[Tensor]]:
    """Flat (names, tensors) of every effective weight / bias of the model in the order ``_Flow`` expects."""
    names, tensors = [], []

    def add(name, t):
        names.append(name)
        tensors.append(t)

    add("upsample.weight", model.upsample.weight)
    add("upsample.bias", model.upsample.bias)
    for k, wn in enumerate(model.WN):
        add(f"convinv.{k}.conv.weight", model.convinv[k].conv.weight)
        add(f"WN.{k}.start.weight", _effective(wn.start))
        add(f"WN.{k}.start.bias", wn.start.bias)
        for i in range(wn.n_layers):
            add(f"WN.{k}.in_layers.{i}.weight", _effective(wn.in_layers[i]))
            add(f"WN.{k}.in_layers.{i}.bias", wn.in_layers[i].bias)
            add(f"WN.{k}.cond_layers.{i}.weight", _effective(wn.cond_layers[i]))
            add(f"WN.{k}.cond_layers.{i}.bias", wn.cond_layers[i].bias)
            add(f"WN.{k}.res_skip_layers.{i}.weight", _effective(wn.res_skip_layers[i]))
            add(f"WN.{k}.res_skip_layers.{i}.bias", wn.res_skip_layers[i].bias)
        add(f"WN.{k}.end.weight", wn.end.weight)
        add(f"WN.{k}.end.bias", wn.end.bias)
    return names, tensors


class _UpsampleOperands:
    """The attributes engine.upsample_cond reads from a PackedWaveGlow, built from live weights on the device."""

    def __init__(self, w_up: Tensor, b_up: Tensor, n_group: int):
        self.mode = "bf16"
        self.n_group = n_group
        self.n_mel = w_up.shape[0]
        self.up_taps = 4
        self.up_stride = w_up.shape[2] // self.up_taps
        self.up_ld_tap = ((self.n_mel + 63) // 64) * 64
        w, b = pack_upsample(w_up, b_up, n_group, self.up_ld_tap)
        self.w_up, self.b_up = w.to(torch.bfloat16), b


_ORDER_CACHE: Dict[str, Tensor] = {}


def _gate_order(dev) -> Tensor:
    """gate_row_order on the device, cached (no host-to-device copy inside a step: CUDA-graph capturable)."""
    key = str(dev)
    if key not in _ORDER_CACHE:
        _ORDER_CACHE[key] = gate_row_order(N_CH).to(dev)
    return _ORDER_CACHE[key]


def _pack_flow(st: Dict[str, Tensor], k: int) -> Dict[str, object]:
    """bf16 / fp32 kernel operands of flow k from the live fp32 weights (all on the device; runs every step)."""
    p = f"WN.{k}."
    dev = st[p + "start.weight"].device
    bf = torch.bfloat16
    order = _gate_order(dev)
    f: Dict[str, object] = {}
    w_end = st[p + "end.weight"][:, :, 0]                                   # [2 n_half, 512]
    n_half = w_end.shape[0] // 2
    f["n_half"] = n_half
    f["w_start"] = st[p + "start.weight"][:, :, 0].contiguous()             # [512, n_half]
    f["b_start"] = st[p + "start.bias"].contiguous()
    mix = torch.zeros(8, 8, device=dev)
    c = 2 * n_half
    mix[:c, :c] = st[f"convinv.{k}.conv.weight"][:, :, 0]
    f["w_mix"] = mix
    # everything below is cast to bf16 FIRST, so the layout shuffles move half the bytes
    w_in = torch.stack([st[p + f"in_layers.{i}.weight"].to(bf) for i in range(N_LAYERS)])             # [8, 1024, 512, 3]
    w_cond = torch.stack([st[p + f"cond_layers.{i}.weight"][:, :, 0].to(bf) for i in range(N_LAYERS)])  # [8, 1024, 640]
    b_gate = torch.stack([st[p + f"in_layers.{i}.bias"] + st[p + f"cond_layers.{i}.bias"] for i in range(N_LAYERS)])
    w_gate = torch.empty((N_LAYERS, 2 * N_CH, 3 * N_CH + N_COND), device=dev, dtype=bf)               # packed row order
    w_gate[:, :, : 3 * N_CH] = w_in[:, order].permute(0, 1, 3, 2).reshape(N_LAYERS, 2 * N_CH, 3 * N_CH)
    w_gate[:, :, 3 * N_CH:] = w_cond[:, order]
    f["w_gate"] = w_gate                                                    # [8, 1024, 2176]
    f["b_gate"] = b_gate[:, order].contiguous()
    # data-gradient operands: W^T with the taps mirrored (tap' = 2 - tap)
    f["wt_in"] = w_in.flip(3).permute(0, 2, 3, 1).reshape(N_LAYERS, N_CH, 3 * 2 * N_CH).contiguous()
    wt_cond = torch.zeros(N_COND_PAD, N_LAYERS, 2 * N_CH, device=dev, dtype=bf)       # [cond ch][layer][gate ch]: K = layer*1024 + o
    wt_cond[:N_COND] = w_cond.permute(2, 0, 1)
    f["wt_cond"] = wt_cond.reshape(N_COND_PAD, N_LAYERS * 2 * N_CH)
    w_rs = [st[p + f"res_skip_layers.{i}.weight"][:, :, 0] for i in range(N_LAYERS)]             # [1024|512, 512]
    b_rs = [st[p + f"res_skip_layers.{i}.bias"] for i in range(N_LAYERS)]
    f["w_res"] = [w_rs[i][:N_CH].to(bf).contiguous() for i in range(N_LAYERS - 1)]
    f["b_res"] = [b_rs[i][:N_CH].contiguous() for i in range(N_LAYERS - 1)]
    f["wt_rs"] = [w_rs[i].t().to(bf).contiguous() for i in range(N_LAYERS)]                      # [512, 1024|512]
    w_skip = [w_rs[i][N_CH:] if i < N_LAYERS - 1 else w_rs[i] for i in range(N_LAYERS)]          # [512, 512] each
    b_skip = sum(b_rs[i][N_CH:] if i < N_LAYERS - 1 else b_rs[i] for i in range(N_LAYERS))
    f["w_skip_f32"] = w_skip
    f["b_skip_total"] = b_skip
    w_skip_cat = torch.cat(w_skip, dim=1)                                   # [512, 8*512]
    w_end_t = torch.zeros(N_CH, 8, device=dev)
    w_end_t[:, :c] = w_end.t()
    f["w_end_t"] = w_end_t
    f["w_end"] = w_end
    b_end = torch.zeros(8, device=dev)
    b_end[:c] = st[p + "end.bias"] + w_end @ b_skip
    f["b_end"] = b_end
    # WN.end composed with the skip GEMM (engine "skip16" path; packing.pack_skip_end16 on the device): [16, 8*512]
    # bf16, rows 0..7 the hi parts of W_end W_skip, rows 8..15 the lo parts
    comp = torch.zeros(8, w_skip_cat.shape[1], device=dev)
    comp[:c] = w_end @ w_skip_cat
    hi = comp.to(bf)
    f["w_skip16"] = torch.cat([hi, (comp - hi.float()).to(bf)], dim=0).contiguous()
    return f


# ------------------------------------------------------------------------------------------------ forward
class _Saved:
    pass


def _forward(st: Dict[str, Tensor], n_flows: int, n_group: int, mel: Tensor, audio: Tensor):
    from . import engine
    b, _, frames = mel.shape
    n = audio.shape[1]
    s = _lib.stream_ptr()
    up = _UpsampleOperands(st["upsample.weight"], st["upsample.bias"], n_group)
    up_len = (frames - 1) * up.up_stride + up.up_stride * up.up_taps
    assert up_len >= n, "upsampled spectrogram shorter than audio"            # glow.py:216
    t = n // n_group
    if t > frames * up.up_stride // n_group:
        raise RuntimeError("audio longer than 256 * frames is not supported by the regrouped upsample GEMM")
    cond = engine.upsample_cond(up, mel)
    if cond.shape[1] > t:
        cond = cond[:, :t].contiguous()                                       # glow.py:217-218
    dev = mel.device
    bf = torch.bfloat16
    x = audio[:, : t * n_group].reshape(b, t, n_group).float().contiguous().clone()
    sv = _Saved()
    sv.b, sv.t, sv.frames, sv.cond, sv.mel, sv.up = b, t, frames, cond, mel, up
    sv.flows, sv.packs = [], []
    log_s_list = []
    for k in range(n_flows):
        f = _pack_flow(st, k)
        nh = f["n_half"]
        fs = _Saved()
        fs.x_pre = x.clone()
        _lib.call("wgb_flow_mix", x, f["w_mix"], b * t, 2 * nh, s)
        fs.x_mix = x.clone()
        fs.h = torch.empty((N_LAYERS, b, t, N_CH), device=dev, dtype=bf)
        fs.acts = torch.empty((N_LAYERS, b, t, N_CH), device=dev, dtype=bf)
        fs.ts = torch.empty((N_LAYERS, b, t, 2 * N_CH), device=dev, dtype=bf)
        _lib.call("wgb_wn_start_padded", x, f["w_start"], f["b_start"], fs.h[0], 1, b, t, t, N_CH, nh, s)
        for i in range(N_LAYERS):
            _lib.call("wgb_tc2_wn_gate_train", fs.h[i], cond, f["w_gate"][i], f["b_gate"][i], fs.acts[i], fs.ts[i], b, t,
                      2 ** i, s)
            if i < N_LAYERS - 1:
                _lib.call("wgb_tc2_wn_res", fs.acts[i], f["w_res"][i], f["b_res"][i], fs.h[i], fs.h[i + 1], b, t, t,
                          None, None, 0, s)
        fs.log_s = torch.empty((b, nh, t), device=dev, dtype=torch.float32)
        _lib.call("wgb_tc_wn_skip16_end", fs.acts, N_LAYERS, f["w_skip16"], f["b_end"], x, None, fs.log_s, b, t, nh, 1,
                  None, None, 0, None, 0, None, None, s)
        log_s_list.append(fs.log_s)
        sv.flows.append(fs)
        sv.packs.append(f)
    z = torch.empty((b, n_group, t), device=dev, dtype=torch.float32)
    _lib.call("wgb_flow_to_z", x, z, b, t, s)
    return z, log_s_list, sv


# ------------------------------------------------------------------------------------------------ backward
class _GradReducer:
    """Data-parallel reduction of the EFFECTIVE-weight gradients while the backward pass is still running
    (waveglow/distributed.py:105-141 averages the parameter gradients once the whole backward is done).

    The map from effective-weight gradients to parameter gradients (weight norm's backward, identity for the rest) is
    linear with coefficients that are identical on every rank, so averaging dL/dw per flow and letting torch's weight-norm
    backward run on the averages leaves exactly the averaged parameter gradients in ``.grad``.  Each flow's gradients are
    packed into one flat bucket the moment the flow's backward is finished and all-reduced asynchronously on NCCL's own
    stream (NVLink / NVSwitch), overlapping the remaining flows' kernels; ``finish`` waits, scales by 1 / world and hands
    back views.  overlap=False issues the same buckets only at the end (the A/B reference)."""

    def __init__(self, group, overlap: bool):
        import torch.distributed as dist
        self.dist, self.group, self.overlap = dist, group, overlap
        self.world = dist.get_world_size(group)
        self.pending = []          # (work, flat, [(name, shape, offset, numel)])
        self.deferred = []

    def add(self, grads: Dict[str, Tensor], names: Sequence[str]) -> None:
        if self.world == 1:
            return
        metas, off = [], 0
        for n in names:
            g = grads[n]
            metas.append((n, g.shape, off, g.numel()))
            off += g.numel()
        flat = torch.cat([grads[n].reshape(-1) for n in names])
        if self.overlap:
            self.pending.append((self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True),
                                 flat, metas))
        else:
            self.deferred.append((flat, metas))

    def finish(self, grads: Dict[str, Tensor]) -> None:
        for flat, metas in self.deferred:
            self.pending.append((self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True),
                                 flat, metas))
        self.deferred = []
        for work, flat, metas in self.pending:
            work.wait()                                   # the current stream waits for NCCL's
            flat.mul_(1.0 / self.world)
            for n, shape, off, numel in metas:
                grads[n] = flat[off: off + numel].view(shape)
        self.pending = []


def _backward(st: Dict[str, Tensor], sv: _Saved, n_flows: int, n_group: int, g_z: Tensor,
              g_log_s: Sequence[Optional[Tensor]], reducer: Optional[_GradReducer] = None) -> Dict[str, Tensor]:
    b, t, cond = sv.b, sv.t, sv.cond
    rows = b * t
    dev = cond.device
    s = _lib.stream_ptr()
    bf, f32 = torch.bfloat16, torch.float32
    grads: Dict[str, Tensor] = {}
    g_x = g_z.permute(0, 2, 1).contiguous().float()                            # [B, T, 8], the layout of x
    g_cond = torch.zeros((b, t, N_COND_PAD), device=dev, dtype=f32)
    g_out = torch.empty((rows, 8), device=dev, dtype=f32)
    g_skip = torch.empty((b, t, N_CH), device=dev, dtype=bf)
    g_h = torch.empty((b, t, N_CH), device=dev, dtype=bf)
    g_acts = torch.empty((b, t, N_CH), device=dev, dtype=bf)
    zero_bias = torch.zeros(N_CH, device=dev, dtype=f32)
    zero_h = torch.zeros((b, t, N_CH), device=dev, dtype=bf)
    stack = torch.empty((b, t, 64), device=dev, dtype=bf)                      # hi/lo(g_out) | hi/lo(x_mix) | 1 (wgb_coupling_bwd)
    L = N_LAYERS
    for k in reversed(range(n_flows)):
        f, fs = sv.packs[k], sv.flows[k]
        p = f"WN.{k}."
        nh = f["n_half"]
        c = 2 * nh
        base = 8 - c
        gls = g_log_s[k]
        gls = None if gls is None else gls.float().contiguous()
        # per-flow gradient buffers: the kernels write straight into slices of these (no per-layer torch ops)
        g_all = torch.empty((L, 64, N_CH), device=dev, dtype=f32)              # stack^T acts_i: rows 0..7 + 8..15 = G_i
        r_all = torch.empty((L, 64, N_CH), device=dev, dtype=f32)              # stack^T g_h: row 32 = column sums of g_h
        d_w_rs = torch.empty((L, 2 * N_CH, N_CH), device=dev, dtype=f32)       # rows 0..511 res, 512..1023 skip
        db_rs = torch.empty((L, 2 * N_CH), device=dev, dtype=f32)
        d_w_in = torch.empty((L, 3, 2 * N_CH, N_CH), device=dev, dtype=f32)    # [layer][tap][out][in]
        db_in = torch.empty((L, 2 * N_CH), device=dev, dtype=f32)
        d_w_cond = torch.empty((L, 2 * N_CH, N_COND), device=dev, dtype=f32)
        _lib.call("wgb_coupling_bwd", g_x, fs.x_mix, fs.log_s, gls, f["w_end_t"], g_out, g_skip, stack, b, t, N_CH, nh, s)
        g_out_sum = torch.empty(8, device=dev, dtype=f32)                      # column sums of g_out = WN.end's bias gradient
        _lib.call("wgb_colsum8_f32", g_out, g_out_sum, rows, 0, s)
        for i in reversed(range(L)):
            d = 2 ** i
            last = i == L - 1
            _lib.call("wgb_tc_wgrad", stack, fs.acts[i], g_all[i], b, t, 64, N_CH, 1, 1, 0, s)
            if last:
                _lib.call("wgb_tc2_wn_res_seg", g_skip, None, 1, 0, f["wt_rs"][i], zero_bias, zero_h, g_acts, b, t, t, N_CH,
                          0, 0, s)
            else:
                _lib.call("wgb_tc2_wn_res_seg", g_h, g_skip, 2, 0b10, f["wt_rs"][i], zero_bias, zero_h, g_acts, b, t, t,
                          N_CH, 0, 0, s)
                _lib.call("wgb_tc_wgrad", g_h, fs.acts[i], d_w_rs[i, :N_CH], b, t, N_CH, N_CH, 1, 1, 0, s)
                _lib.call("wgb_tc_wgrad", stack, g_h, r_all[i], b, t, 64, N_CH, 1, 1, 0, s)       # row 32: bias gradient
            g_in = fs.ts[i]                                                    # (tanh | sigmoid) -> gradient, in place
            _lib.call("wgb_gate_bwd", g_acts, g_in, db_in[i], rows, N_CH, s)
            _lib.call("wgb_tc_wgrad", g_in, fs.h[i], d_w_in[i], b, t, 2 * N_CH, N_CH, 3, d, 0, s)
            _lib.call("wgb_tc_wgrad", g_in, cond, d_w_cond[i], b, t, 2 * N_CH, N_COND, 1, 1, 0, s)
            if last:
                g_h.zero_()
            _lib.call("wgb_tc2_wn_res_taps", g_in, f["wt_in"][i], zero_bias, g_h, g_h, b, t, t, 2 * N_CH, 3, d, s)
        # conditioning gradient of the whole flow: the eight layers' g_in (still in fs.ts) K-concatenated, one fp32
        # read-modify-write of g_cond per flow
        _lib.call("wgb_tc_gemm_seg", fs.ts, None, L, 0, f["wt_cond"], None, g_cond, g_cond, 0, b, t, N_COND_PAD,
                  2 * N_CH, 0, 0, 0, 1, s)
        # WN.start (g_h is now the gradient of h_0)
        _lib.call("wgb_tc_wgrad", stack, g_h, r_all[L - 1], b, t, 64, N_CH, 1, 1, 0, s)       # rows 16..31: x_mix^T g_h0
        _lib.call("wgb_start_bwd", g_x, g_h, f["w_start"], rows, N_CH, nh, s)
        # invertible 1x1 conv
        d_mix = torch.empty((8, 8), device=dev, dtype=f32)
        _lib.call("wgb_mix_bwd", g_x, fs.x_pre, f["w_mix"], d_mix, rows, c, s)
        # parameter-space algebra (8 x 512 matrices): WN.end and the skip rows of res_skip through G_i
        w_end = f["w_end"]                                                     # [c, 512]
        w_skip = torch.stack(f["w_skip_f32"])                                  # [L, 512 (out), 512 (in)]
        g_c = g_all[:, :c] + g_all[:, 8: 8 + c]                                # [L, c, 512] (hi + lo parts of g_out)
        db_rs[: L - 1, :N_CH] = r_all[: L - 1, 32]
        d_start = r_all[L - 1, 16:24] + r_all[L - 1, 24:32]                    # [8, 512] = x_mix^T g_h0
        d_w_end = torch.einsum("ijn,imn->jm", g_c, w_skip) + torch.outer(g_out_sum[:c], f["b_skip_total"])
        d_w_skip = torch.einsum("jm,ijn->imn", w_end, g_c)                     # [L, 512, 512]
        db_skip = w_end.t() @ g_out_sum[:c]
        d_w_rs[: L - 1, N_CH:] = d_w_skip[: L - 1]
        d_w_rs[L - 1, :N_CH] = d_w_skip[L - 1]
        db_rs[: L - 1, N_CH:] = db_skip
        db_rs[L - 1, :N_CH] = db_skip
        d_w_in_p = d_w_in.permute(0, 2, 3, 1).contiguous()                     # [L, out, in, tap]
        grads[p + "end.weight"] = d_w_end.unsqueeze(2)
        grads[p + "end.bias"] = g_out_sum[:c].clone()
        grads[p + "start.weight"] = d_start[base: base + nh].t().contiguous().unsqueeze(2)
        grads[p + "start.bias"] = r_all[L - 1, 32].clone()
        grads[f"convinv.{k}.conv.weight"] = d_mix[:c, :c].contiguous().unsqueeze(2)
        for i in range(L):
            n_rs = 2 * N_CH if i < L - 1 else N_CH
            grads[p + f"res_skip_layers.{i}.weight"] = d_w_rs[i, :n_rs].unsqueeze(2)
            grads[p + f"res_skip_layers.{i}.bias"] = db_rs[i, :n_rs]
            grads[p + f"in_layers.{i}.weight"] = d_w_in_p[i]
            grads[p + f"in_layers.{i}.bias"] = db_in[i]
            grads[p + f"cond_layers.{i}.weight"] = d_w_cond[i].unsqueeze(2)
            grads[p + f"cond_layers.{i}.bias"] = db_in[i].clone()
        fs.h = fs.acts = fs.ts = None                                          # release this flow's activations
        if reducer is not None:                                                # this flow's gradients are final: reduce
            reducer.add(grads, [n for n in grads if n.startswith(p) or n == f"convinv.{k}.conv.weight"])
    up = sv.up
    ksize = up.up_stride * up.up_taps
    d_up = torch.empty((up.n_mel, up.n_mel, ksize), device=dev, dtype=f32)
    db_up = torch.empty(up.n_mel, device=dev, dtype=f32)
    _lib.call("wgb_upsample_wgrad", sv.mel, g_cond, d_up, db_up, b, up.n_mel, sv.frames, t, N_COND_PAD, ksize, up.up_stride,
              n_group, s)
    grads["upsample.weight"] = d_up
    grads["upsample.bias"] = db_up
    if reducer is not None:
        reducer.add(grads, ["upsample.weight", "upsample.bias"])
        reducer.finish(grads)
    return grads


class _Flow(torch.autograd.Function):
    """(mel, audio, *effective weights) -> (z, *log_s): the whole flow as one autograd node."""

    @staticmethod
    def forward(ctx, names, n_flows, n_group, dp, mel, audio, *weights):
        st = {n: w.detach().float() for n, w in zip(names, weights)}
        ctx.dp = dp
        with torch.cuda.device(mel.device):
            z, log_s_list, sv = _forward(st, n_flows, n_group, mel.detach().float().contiguous(),
                                         audio.detach().float().contiguous())
        ctx.names, ctx.n_flows, ctx.n_group, ctx.st, ctx.sv = names, n_flows, n_group, st, sv
        return (z, *log_s_list)

    @staticmethod
    def backward(ctx, g_z, *g_log_s):
        sv = ctx.sv
        if sv is None:
            raise RuntimeError("the flow's saved activations were already consumed (backward twice?)")
        if g_z is None:
            g_z = torch.zeros((sv.b, ctx.n_group, sv.t), device=sv.cond.device)
        reducer = None
        if ctx.dp is not None:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(ctx.dp[0]) > 1:
                reducer = _GradReducer(ctx.dp[0], ctx.dp[1])
        with torch.cuda.device(sv.cond.device):
            grads = _backward(ctx.st, sv, ctx.n_flows, ctx.n_group, g_z, g_log_s, reducer)
        ctx.sv = None
        out = []
        for n, need in zip(ctx.names, ctx.needs_input_grad[6:]):
            out.append(grads[n] if need else None)
        return (None, None, None, None, None, None, *out)


class _LogDet(torch.autograd.Function):
    """scale * log|det W| for a c x c matrix, c <= 8 (glow.py:100) as one capturable kernel; backward = scale W^-T."""

    @staticmethod
    def forward(ctx, w, scale):
        c = w.shape[0]
        w2 = w.detach().float().reshape(c, c).contiguous()
        out = torch.empty(1, device=w.device, dtype=torch.float32)
        inv_t = torch.empty((c, c), device=w.device, dtype=torch.float32)
        with torch.cuda.device(w.device):
            _lib.call("wgb_logdet", w2, out, inv_t, c, _lib.stream_ptr())
        ctx.save_for_backward(inv_t)
        ctx.scale, ctx.shape = scale, w.shape
        return out[0] * scale

    @staticmethod
    def backward(ctx, g):
        (inv_t,) = ctx.saved_tensors
        return (g * ctx.scale * inv_t).reshape(ctx.shape), None


def forward_autograd(model, spect: Tensor, audio: Tensor):
    """WaveGlow.forward under autograd (glow.py:207-249): returns (z, log_s_list, log_det_W_list) whose backward fills
    the ``.grad`` of every model parameter."""
    if model.mode != "bf16":
        raise RuntimeError("the training direction exists for the bf16 tensor-core path only")
    model._check_supported()
    _lib.require_b200(spect.device)
    names, weights = effective_weights(model)
    outs = _Flow.apply(names, model.n_flows, model.n_group, getattr(model, "_dp_allreduce", None), spect, audio, *weights)
    z, log_s_list = outs[0], list(outs[1:])
    bt = z.shape[0] * z.shape[2]
    log_det = [_LogDet.apply(model.convinv[k].conv.weight, float(bt)) for k in range(model.n_flows)]   # glow.py:100
    return z, log_s_list, log_det


# ------------------------------------------------------------------------------------------------ optimiser / DP
class FusedAdam:
    """torch.optim.Adam(model.parameters(), lr) of train.py:79 as ONE kernel per step over a flat fp32 buffer.
    Parameters are re-pointed at views of the flat buffer (their values are preserved); gradients are gathered into a
    flat buffer each step (that copy is the only per-parameter work)."""

    def __init__(self, params, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no parameters to optimise")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam needs parameters on a CUDA device (there is no CPU path)")
        self.lr, self.betas, self.eps, self.step_count = lr, betas, eps, 0
        sizes = [p.numel() for p in self.params]
        pitch = [(sz + 63) // 64 * 64 for sz in sizes]          # every parameter starts 256 B aligned (vector loads)
        self.n = sum(pitch)
        self.flat = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.m = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(self.n, device=dev, dtype=torch.float32)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        off = 0
        self.views = []
        for p, sz, pt in zip(self.params, sizes, pitch):
            self.flat[off: off + sz].copy_(p.data.reshape(-1))
            p.data = self.flat[off: off + sz].view_as(p.data)
            self.views.append(self.grad[off: off + sz].view_as(p.data))
            off += pt

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    # ------------------------------------------------------------------ checkpoints (torch.optim.Adam's format)
    def state_dict(self):
        """The layout ``torch.optim.Adam.state_dict()`` produces (train.py:57-62 stores it in the checkpoint)."""
        state = {}
        if self.step_count > 0:
            for i, (p, gv) in enumerate(zip(self.params, self.views)):
                off = gv.storage_offset()
                sz = p.numel()
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.m[off: off + sz].view_as(p).detach().cpu().clone(),
                            "exp_avg_sq": self.v[off: off + sz].view_as(p).detach().cpu().clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts ``torch.optim.Adam`` state (of this class or of the reference's optimiser over the same parameters)."""
        group = sd["param_groups"][0]
        self.lr, self.betas, self.eps = group["lr"], tuple(group["betas"]), group["eps"]
        if group.get("weight_decay", 0) or group.get("amsgrad", False):
            raise ValueError("FusedAdam has no weight decay / amsgrad")
        self.m.zero_()
        self.v.zero_()
        step = 0
        for i, st in sd["state"].items():
            i = int(i)
            gv = self.views[i]
            off, sz = gv.storage_offset(), self.params[i].numel()
            self.m[off: off + sz].copy_(st["exp_avg"].reshape(-1))
            self.v[off: off + sz].copy_(st["exp_avg_sq"].reshape(-1))
            step = max(step, int(float(st["step"])))
        self.step_count = step
        self.step_dev.fill_(step)

    def gather_grads(self) -> Tensor:
        """Copy every .grad into the flat gradient buffer (missing grads count as zero); returns that buffer."""
        have = [(gv, p.grad) for p, gv in zip(self.params, self.views) if p.grad is not None]
        if len(have) != len(self.params):
            self.grad.zero_()
        if have:
            torch._foreach_copy_([gv for gv, _ in have], [g for _, g in have])     # one batched launch sequence
        return self.grad

    def step(self, grad_scale: float = 1.0, gathered: bool = False):
        """The step counter lives on the device (incremented by the kernel), so a captured step replays correctly."""
        if not gathered:
            self.gather_grads()
        self.step_count += 1
        bump_param_generation()                 # the kernel rewrites the parameters behind torch's version counters
        with torch.cuda.device(self.flat.device):
            _lib.call("wgb_adam_step_dev", self.flat, self.grad, self.m, self.v, self.n, float(self.lr),
                      float(self.betas[0]), float(self.betas[1]), float(self.eps), self.step_dev, float(grad_scale),
                      _lib.stream_ptr())


def apply_gradient_allreduce(module, group=None, overlap: bool = True):
    """waveglow/distributed.py:90-142 for the drop-in WaveGlow (same name, same contract): broadcast rank 0's state to
    every rank (:99-103), then make every ``loss.backward()`` leave gradients AVERAGED over the ranks in ``.grad``
    (:105-141), so the training loop is the reference's: ``model = apply_gradient_allreduce(model)`` ...
    ``loss.backward(); optimizer.step()``.  Instead of flattening all parameter gradients after the backward pass, each
    flow's effective-weight gradients are all-reduced as soon as that flow's backward is done (``_GradReducer``), which
    hides the collective behind the remaining flows.  The log-determinant term's gradient (B T W^-T per rank, identical
    on every rank for equal per-rank batches, which the DistributedSampler + drop_last loader guarantees) needs no
    reduction.  overlap=False: the same buckets, issued after the last flow."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("apply_gradient_allreduce needs an initialised process group (init_distributed)")
    for p in module.state_dict().values():
        if torch.is_tensor(p):
            dist.broadcast(p, 0, group=group)
    bump_param_generation()
    module._dp_allreduce = (group, bool(overlap))
    return module


def allreduce_gradients(optimizer: FusedAdam, group=None, gathered: bool = False) -> float:
    """Data-parallel gradient reduction of waveglow/distributed.py:90-142 (flatten -> all_reduce -> divide by world
    size) on the optimiser's flat buffer: ONE collective per step; returns the scale to hand to ``step``.
    ``gathered``: the flat buffer already holds this step's gradients (GraphedTrainStep gathers inside its graph)."""
    import torch.distributed as dist
    flat = optimizer.grad if gathered else optimizer.gather_grads()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 1.0
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / dist.get_world_size(group)


class GraphedTrainStep:
    """One training step (forward, criterion, backward, gradient gather [, Adam]) captured in a CUDA graph for a fixed
    batch shape: the ~4000 launches of a step (this package's kernels plus the parameter-sized torch ops around them)
    become one graph launch, which matters because the step is otherwise host-bound below ~32 x 16000 samples.

        step = GraphedTrainStep(model, optimizer, criterion, batch, n_mel, frames, samples)
        loss = step(mel, audio)            # copies the batch into the graph's static buffers, replays, returns the loss

    With ``world_size > 1`` build it with ``include_optimizer=False``: the graph then ends after the gradient gather,
    and the caller runs ``allreduce_gradients`` + ``optimizer.step(gathered=True)`` eagerly (two launches) -- the
    fastest data-parallel step measured; the eager loop uses ``apply_gradient_allreduce``."""

    def __init__(self, model, optimizer: FusedAdam, criterion, batch: int, n_mel: int, frames: int, samples: int,
                 include_optimizer: bool = True, warmup: int = 2):
        dev = optimizer.flat.device
        self.model, self.opt, self.include_optimizer = model, optimizer, include_optimizer
        self.mel = torch.zeros((batch, n_mel, frames), device=dev, dtype=torch.float32)
        self.audio = torch.zeros((batch, samples), device=dev, dtype=torch.float32)
        # warm-up on a side stream (allocator, cudaFuncSetAttribute, lazy module init), restoring the training state
        # afterwards so that building the graph does not count as training
        keep = [t.clone() for t in (optimizer.flat, optimizer.m, optimizer.v, optimizer.step_dev)]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._one_step(criterion)
        torch.cuda.current_stream(dev).wait_stream(side)
        optimizer.zero_grad()
        if getattr(model, "_dp_allreduce", None) is not None:
            # measured on 2 x B200 (profiles/r02g_*): capturing the per-flow NCCL all-reduces into the graph is SLOWER than
            # graph + one flat all-reduce (120.2 vs 119.2 ms: the collectives' CTAs compete with the persistent GEMM CTAs
            # for SMs) and the process group then hangs at teardown
            raise RuntimeError("GraphedTrainStep on a model under apply_gradient_allreduce is not supported: build it "
                               "with include_optimizer=False and call allreduce_gradients + optimizer.step after it")
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._one_step(criterion)
        for dst, src in zip((optimizer.flat, optimizer.m, optimizer.v, optimizer.step_dev), keep):
            dst.copy_(src)
        optimizer.step_count = int(keep[3].item())

    def _one_step(self, criterion):
        self.opt.zero_grad()
        loss = criterion(self.model((self.mel, self.audio)))
        loss.backward()
        self.opt.gather_grads()
        if self.include_optimizer:
            self.opt.step(gathered=True)
        return loss.detach()

    def __call__(self, mel: Tensor, audio: Tensor) -> Tensor:
        self.mel.copy_(mel, non_blocking=True)
        self.audio.copy_(audio, non_blocking=True)
        self.graph.replay()
        if self.include_optimizer:
            self.opt.step_count += 1
            bump_param_generation()
        return self.loss
