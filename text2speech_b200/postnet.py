"""Drop-in Tacotron-2 Postnet (reference: tacotron/modules.py:94-137, ConvNorm :177-197) on the B200 kernels.

Five 1-d convolutions (kernel 5, 512 channels) with BatchNorm and tanh between them: the non-autoregressive block that
turns the decoder's mel into ``mel_outputs_postnet`` — the tensor the reference hands to ``waveglow.infer``
(inference.py:87-94).  Same constructor (``Postnet(hparams)``), sub-module names and ``state_dict`` keys as the
reference (``convolutions.{i}.0.conv.{weight,bias}``, ``convolutions.{i}.1.{weight,bias,running_mean,running_var,
num_batches_tracked}``).  Inference only: BatchNorm uses its running statistics and is folded into the conv weights,
dropout is the identity (the reference passes ``self.training``); calling it in training mode raises.

``mode`` 'bf16': every conv is an implicit GEMM on tcgen05 (wgb_tc_conv1d: taps = time-shifted TMA boxes, bias + tanh
in the epilogue, bf16 activations between layers).  'fp32': CUDA-core validation path (wgb_sgemm_f32 per tap).
The autoregressive decoder, attention and encoder LSTM stay in reference PyTorch (out of scope).
"""
from __future__ import annotations

import torch

from . import _lib


class ConvNorm(torch.nn.Module):
    """tacotron/modules.py:177-197: Conv1d with xavier init, parameters only (the kernels do the work)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=None, dilation=1, bias=True,
                 w_init_gain="linear"):
        super().__init__()
        if padding is None:
            assert kernel_size % 2 == 1
            padding = int(dilation * (kernel_size - 1) / 2)
        self.conv = torch.nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                                    dilation=dilation, bias=bias)
        torch.nn.init.xavier_uniform_(self.conv.weight, gain=torch.nn.init.calculate_gain(w_init_gain))

    def forward(self, signal):
        raise RuntimeError("ConvNorm is a parameter container here; run the owning Postnet")


def _round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


class _ConvBnStack(torch.nn.Module):
    """``convolutions`` = ModuleList of Sequential(ConvNorm, BatchNorm1d) run as folded implicit GEMMs; ``acts[i]`` is
    the activation after layer i (0 none, 1 tanh, 2 relu), applied in the GEMM epilogue."""

    def __init__(self, chans, kernel_size, acts, gains):
        super().__init__()
        pad = int((kernel_size - 1) / 2)
        self.convolutions = torch.nn.ModuleList()
        for i in range(len(chans) - 1):
            self.convolutions.append(torch.nn.Sequential(
                ConvNorm(chans[i], chans[i + 1], kernel_size=kernel_size, stride=1, padding=pad, dilation=1,
                         w_init_gain=gains[i]),
                torch.nn.BatchNorm1d(chans[i + 1])))
        self.acts = list(acts)
        self.mode = "bf16"
        self._pack = None

    # ------------------------------------------------------------------ packing
    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _packed(self, device):
        key = (self.mode, str(device), self._signature())
        if self._pack is None or self._pack[0] != key:
            layers = []
            for seq in self.convolutions:
                conv, bn = seq[0].conv, seq[1]
                w, b = conv.weight.detach().double().cpu(), conv.bias.detach().double().cpu()
                scale = bn.weight.detach().double().cpu() / torch.sqrt(bn.running_var.detach().double().cpu() + bn.eps)
                w = w * scale[:, None, None]                                           # BatchNorm (eval) folded in
                b = (b - bn.running_mean.detach().double().cpu()) * scale + bn.bias.detach().double().cpu()
                c_out, c_in, taps = w.shape
                if self.mode == "bf16":
                    cp, npad = _round_up(c_in, 64), _round_up(c_out, 256)
                    wp = torch.zeros(npad, taps, cp, dtype=torch.float64)
                    wp[:c_out, :, :c_in] = w.permute(0, 2, 1)                          # K index = tap * cp + channel
                    bp = torch.zeros(npad, dtype=torch.float64)
                    bp[:c_out] = b
                    layers.append(dict(w=wp.reshape(npad, taps * cp).to(device, torch.bfloat16), b=bp.float().to(device),
                                       c_in=c_in, cp=cp, c_out=c_out, npad=npad, taps=taps))
                else:
                    cp = _round_up(c_in, 4)
                    wp = torch.zeros(taps, c_out, cp, dtype=torch.float64)
                    wp[:, :, :c_in] = w.permute(2, 0, 1)
                    layers.append(dict(w=wp.float().to(device).contiguous(), b=b.float().to(device), c_in=c_in, cp=cp,
                                       c_out=c_out, npad=c_out, taps=taps))
            self._pack = (key, layers)
        return self._pack[1]

    # ------------------------------------------------------------------ reference API
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, C_in, F] (CUDA) -> [B, C_out, F]   (eval mode: BatchNorm running statistics, dropout = identity)."""
        name = type(self).__name__
        if self.training:
            raise RuntimeError(f"{name} on the B200 kernels is inference-only (BatchNorm folded, no dropout): call .eval()")
        if not x.is_cuda:
            raise RuntimeError(f"{name} needs a CUDA tensor on a B200; there is no CPU fallback")
        _lib.require_b200(x.device)
        with torch.cuda.device(x.device):
            return self._forward(x)

    def _forward(self, x: torch.Tensor) -> torch.Tensor:
        layers = self._packed(x.device)
        b, c, f = x.shape
        s = _lib.stream_ptr()
        n = len(layers)
        if self.mode == "bf16":
            cur = torch.zeros((b, f, layers[0]["cp"]), device=x.device, dtype=torch.bfloat16)
            cur[:, :, :c] = x.permute(0, 2, 1)
            for i, ly in enumerate(layers):
                last = i == n - 1
                out = torch.empty((b, f, ly["npad"]), device=x.device, dtype=torch.float32 if last else torch.bfloat16)
                _lib.call("wgb_tc_conv1d", cur, ly["w"], ly["b"], out, 0 if last else 1, b, f, ly["npad"], ly["cp"],
                          ly["taps"], 1, self.acts[i], s)
                cur = out
            return cur[:, :, : layers[-1]["c_out"]].permute(0, 2, 1).contiguous().to(x.dtype)
        cur = torch.zeros((b, f, layers[0]["cp"]), device=x.device, dtype=torch.float32)
        cur[:, :, :c] = x.float().permute(0, 2, 1)
        for i, ly in enumerate(layers):
            out = torch.empty((b, f, ly["c_out"]), device=x.device, dtype=torch.float32)
            half = (ly["taps"] - 1) // 2
            for tap in range(ly["taps"]):
                _lib.call("wgb_sgemm_f32", cur, ly["w"][tap], ly["b"] if tap == 0 else None, out, 0, b, f, ly["c_out"],
                          ly["cp"], ly["cp"], f * ly["cp"], ly["cp"], ly["c_out"], f * ly["c_out"], tap - half,
                          int(tap > 0), s)
            if self.acts[i]:
                _lib.call("wgb_act_f32", out, out.numel(), self.acts[i], s)
            cur = out
        return cur.permute(0, 2, 1).contiguous().to(x.dtype)


class Postnet(_ConvBnStack):
    """tacotron/modules.py:94-137: five k = 5 convs, tanh after all but the last."""

    def __init__(self, hparams):
        n_mel, dim = hparams["n_mel_channels"], hparams["postnet_embedding_dim"]
        k, n_conv = hparams["postnet_kernel_size"], hparams["postnet_n_convolutions"]
        super().__init__([n_mel] + [dim] * (n_conv - 1) + [n_mel], k, [1] * (n_conv - 1) + [0],
                         ["tanh"] * (n_conv - 1) + ["linear"])


class EncoderConvs(_ConvBnStack):
    """The three-conv bank of the Tacotron-2 Encoder (tacotron/tacotron.py:175-186 built, :197-198 / :211-212 run:
    ``x = F.dropout(F.relu(conv(x)), 0.5, self.training)`` per layer), eval mode.  Same ``convolutions.{i}.{0.conv,1}.*``
    state_dict keys as the reference Encoder, so ``load_state_dict(encoder.state_dict(), strict=False)`` picks the conv
    bank out of a reference Encoder; the bidirectional LSTM that consumes the result (``x.transpose(1, 2)`` first,
    tacotron.py:214-217) stays in reference PyTorch.  Input = embedded text [B, enc_conv_channels, T]."""

    def __init__(self, hparams):
        c, k, n = hparams["enc_conv_channels"], hparams["enc_conv_kernel_size"], hparams["enc_conv_num_layers"]
        super().__init__([c] * (n + 1), k, [2] * n, ["relu"] * n)
