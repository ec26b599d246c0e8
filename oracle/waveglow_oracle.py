"""CPU oracle: WaveGlow inverse/forward flow as plain torch functions (TEST INFRASTRUCTURE).

A functional restatement (no nn.Module, no hidden state) of the reference's
``waveglow/glow.py``.  It consumes a reference-layout ``state_dict`` (with or without
weight-norm) so the oracle, the reference and the CUDA path all run on the very same
tensors.  Used only by tests/, __graft_entry__.smoke() and bench.py's CPU baseline.

Noise convention (host supplied, replaces the reference's in-place ``.normal_()`` draws at
glow.py:260-267 and :285-288): one tensor ``z[B, n_group, T]`` laid out like the z that
``WaveGlow.forward`` returns (glow.py:248-249).  With 12 flows / early-every 4 / early-size 2
the initial draw is ``z[:, 4:8]``, the k=8 injection ``z[:, 2:4]`` and the k=4 injection
``z[:, 0:2]`` — i.e. the flow at step k always works on the LAST ``C_k`` channels.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def flow_channels(n_flows: int, n_group: int, n_early_every: int, n_early_size: int) -> List[int]:
    """Channels the flow k operates on (glow.py:194-204): 8,8,8,8,6,6,6,6,4,4,4,4 for config.json."""
    out, c = [], n_group
    for k in range(n_flows):
        if k % n_early_every == 0 and k > 0:
            c -= n_early_size
        out.append(c)
    return out


def fold_weight_norm(g: Tensor, v: Tensor) -> Tensor:
    """w[o] = g[o] * v[o] / ||v[o]||_2 with the norm over (in, tap) — torch weight_norm dim=0,
    which is what glow.py:123,138,142,151 apply and glow.py:294-310 later fold."""
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
    return v * (g / norm)


def folded_state(state: Dict[str, Tensor], dtype=torch.float32) -> Dict[str, Tensor]:
    """Return a ``*.weight``-only copy of a reference state_dict (the layout after
    ``WaveGlow.remove_weightnorm``, glow.py:294-302)."""
    out: Dict[str, Tensor] = {}
    for key, val in state.items():
        if key.endswith(".weight_g"):
            stem = key[: -len("_g")]
            out[stem] = fold_weight_norm(val.to(dtype), state[stem + "_v"].to(dtype))
        elif key.endswith(".weight_v"):
            continue
        else:
            out[key] = val.to(dtype)
    return out


def upsample_spect(st: Dict[str, Tensor], mel: Tensor) -> Tensor:
    """ConvTranspose1d(80, 80, 1024, stride 256) of the mel (glow.py:183-185, :213 / :252)."""
    w = st["upsample.weight"]
    return F.conv_transpose1d(mel, w, st["upsample.bias"], stride=w.shape[2] // 4)


def regroup_spect(up: Tensor, n_group: int) -> Tensor:
    """[B, M, n_group*T] -> [B, M*n_group, T] with channel index m*n_group + g holding sample
    n_group*t + g (the unfold/permute/view dance at glow.py:220-221 and :257-258)."""
    b, m, n = up.shape
    t = n // n_group
    return up[:, :, : t * n_group].reshape(b, m, t, n_group).permute(0, 1, 3, 2).reshape(b, m * n_group, t)


def wn_layer(st: Dict[str, Tensor], k: int, i: int, h: Tensor, cond: Tensor) -> Tuple[Tensor, Tensor]:
    """Gated layer i of WN k (glow.py:159-164): returns (acts, res_skip_acts); ``st`` is folded."""
    p = f"WN.{k}."
    w_in = st[p + f"in_layers.{i}.weight"]
    n_ch = w_in.shape[1]
    d = 2 ** i                                                               # glow.py:134-137
    pad = (w_in.shape[2] - 1) * d // 2
    u = F.conv1d(h, w_in, st[p + f"in_layers.{i}.bias"], dilation=d, padding=pad)
    u = u + F.conv1d(cond, st[p + f"cond_layers.{i}.weight"], st[p + f"cond_layers.{i}.bias"])
    acts = torch.tanh(u[:, :n_ch]) * torch.sigmoid(u[:, n_ch:])              # glow.py:33-40
    r = F.conv1d(acts, st[p + f"res_skip_layers.{i}.weight"], st[p + f"res_skip_layers.{i}.bias"])
    return acts, r


def wn_stack(st: Dict[str, Tensor], k: int, a0: Tensor, cond: Tensor, n_layers: int,
             taps: Optional[dict] = None) -> Tensor:
    """One WN module (glow.py:154-175) for flow k; ``st`` must be weight-norm-folded.

    taps: optional dict that receives per-layer intermediates
    (``h{i}`` = layer input, ``acts{i}``, ``skip{i}``) for the per-layer parity tests.
    """
    p = f"WN.{k}."
    h = F.conv1d(a0, st[p + "start.weight"], st[p + "start.bias"])           # glow.py:156
    n_ch = h.shape[1]
    total = None
    for i in range(n_layers):
        acts, r = wn_layer(st, k, i, h, cond)
        if taps is not None:
            taps[f"h{i}"] = h
            taps[f"acts{i}"] = acts
        if i < n_layers - 1:                                                 # glow.py:165-169
            h = h + r[:, :n_ch]
            skip = r[:, n_ch:]
        else:
            skip = r
        if taps is not None:
            taps[f"skip{i}"] = skip
        total = skip if total is None else total + skip                      # glow.py:171-174
    return F.conv1d(total, st[p + "end.weight"], st[p + "end.bias"])         # glow.py:175


def _arch(st: Dict[str, Tensor]) -> Tuple[int, int, List[int]]:
    n_flows = 1 + max(int(k.split(".")[1]) for k in st if k.startswith("convinv."))
    n_layers = 1 + max(int(k.split(".")[3]) for k in st if k.startswith("WN.0.in_layers."))
    chans = [st[f"convinv.{k}.conv.weight"].shape[0] for k in range(n_flows)]
    return n_flows, n_layers, chans


def waveglow_infer(state: Dict[str, Tensor], mel: Tensor, z: Tensor, sigma: float,
                   taps: Optional[dict] = None) -> Tensor:
    """mel [B,80,F], z [B,n_group,32F] -> audio [B, 256F]   (glow.py:251-292)."""
    st = folded_state(state, mel.dtype)
    n_flows, n_layers, chans = _arch(st)
    n_group = chans[0]
    up = upsample_spect(st, mel)
    kernel, stride = st["upsample.weight"].shape[2], st["upsample.weight"].shape[2] // 4
    up = up[:, :, : up.shape[2] - (kernel - stride)]                         # glow.py:254-255
    cond = regroup_spect(up, n_group)
    x = sigma * z[:, n_group - chans[-1]:, :].to(mel.dtype)                  # glow.py:260-269
    for k in reversed(range(n_flows)):
        c = x.shape[1]
        n_half = c // 2
        a0, a1 = x[:, :n_half], x[:, n_half:]
        sub = {} if taps is not None else None
        out = wn_stack(st, k, a0, cond, n_layers, sub)
        b, s = out[:, :n_half], out[:, n_half:]                              # glow.py:277-278
        a1 = (a1 - b) / torch.exp(s)                                         # glow.py:279
        x = torch.cat([a0, a1], 1)
        w = st[f"convinv.{k}.conv.weight"].squeeze(-1)
        w_inv = torch.linalg.inv(w.double()).to(x.dtype)                     # glow.py:88-95
        x = torch.einsum("oc,bct->bot", w_inv, x)                            # glow.py:96
        if taps is not None:
            sub["wn_out"] = out
            sub["x_after"] = x
            taps[k] = sub
        if k > 0 and chans[k - 1] > c:                                       # glow.py:284-289
            grow = chans[k - 1] - c
            lo = n_group - chans[k - 1]
            x = torch.cat([sigma * z[:, lo: lo + grow].to(x.dtype), x], 1)
    return x.permute(0, 2, 1).reshape(x.shape[0], -1)                        # glow.py:291


def waveglow_forward(state: Dict[str, Tensor], mel: Tensor, audio: Tensor):
    """(mel [B,80,F], audio [B,N]) -> (z [B,n_group,N/n_group], [log_s]*n_flows, [log_det_W]*n_flows)
    following glow.py:207-249."""
    st = folded_state(state, mel.dtype)
    n_flows, n_layers, chans = _arch(st)
    n_group = chans[0]
    up = upsample_spect(st, mel)
    assert up.shape[2] >= audio.shape[1]                                     # glow.py:216
    up = up[:, :, : audio.shape[1]]
    cond = regroup_spect(up, n_group)
    bsz = audio.shape[0]
    t = audio.shape[1] // n_group
    x = audio[:, : t * n_group].reshape(bsz, t, n_group).permute(0, 2, 1)    # glow.py:223
    early, log_s_list, log_det_list = [], [], []
    for k in range(n_flows):
        if x.shape[1] > chans[k]:                                            # glow.py:229-231
            drop = x.shape[1] - chans[k]
            early.append(x[:, :drop])
            x = x[:, drop:]
        w = st[f"convinv.{k}.conv.weight"].squeeze(-1)
        log_det_list.append(bsz * t * torch.logdet(w))                       # glow.py:100
        x = torch.einsum("oc,bct->bot", w, x)                                # glow.py:101
        n_half = x.shape[1] // 2
        a0, a1 = x[:, :n_half], x[:, n_half:]
        out = wn_stack(st, k, a0, cond, n_layers)
        b, log_s = out[:, :n_half], out[:, n_half:]                          # glow.py:241-242
        a1 = torch.exp(log_s) * a1 + b                                       # glow.py:243
        log_s_list.append(log_s)
        x = torch.cat([a0, a1], 1)
    early.append(x)
    return torch.cat(early, 1), log_s_list, log_det_list                     # glow.py:248-249


def waveglow_loss(z: Tensor, log_s_list, log_det_list, sigma: float = 1.0) -> Tensor:
    """WaveGlowLoss (glow.py:43-59)."""
    log_s_total = sum(torch.sum(ls) for ls in log_s_list)
    log_det_total = sum(log_det_list)
    loss = torch.sum(z * z) / (2 * sigma * sigma) - log_s_total - log_det_total
    return loss / z.numel()


def algorithmic_flop_per_group_step(st: Dict[str, Tensor]) -> int:
    """2*MACs of every conv on the infer path per group step (522 302 368 for
    config.json, of which 522 190 848 are the in/cond/res_skip GEMMs)."""
    st = folded_state(st)
    total = 0
    for key, w in st.items():
        if key.endswith(".weight") and key.startswith(("WN.", "convinv.")):
            total += 2 * math.prod(w.shape)
    return total
