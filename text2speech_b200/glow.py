"""Drop-in WaveGlow / WN / Invertible1x1Conv / WaveGlowLoss (reference: waveglow/glow.py).

Same constructor signatures, attribute names and ``state_dict`` keys as the reference, so
``WaveGlow(**config["waveglow_config"])``, ``load_state_dict`` (weight-norm or folded layout),
``WaveGlow.remove_weightnorm`` and pickled-module checkpoints behave as before.  The torch
sub-modules are parameter containers only: ``infer`` / ``forward`` never call them, they pack the
weights once (text2speech_b200.packing) and run hand-written sm_100a kernels through the C ABI
(include/waveglow_b200.h).  There is no CPU / eager fallback: inputs must be CUDA tensors on a
B200, otherwise a RuntimeError is raised.

Numeric modes (``model.mode``):
  'bf16'  tcgen05 BF16 GEMMs, fp32 accumulation / flow state  (per-layer rel-L2 <= 2e-3)
  'fp32'  CUDA-core FP32 validation path                       (<= 1e-5)

Conscious deviations from reference quirks:
  * noise: ``infer(spect, sigma, z=None)`` takes optional host-supplied noise z [B, 8, 32F] in the
    layout ``forward`` returns; the reference draws it in place (glow.py:260-267, :285-288).  When
    z is None it is drawn with torch.randn on the mel's device.
  * W^-1 is recomputed whenever the weights are repacked, not cached forever (glow.py:89-95).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, engine
from .packing import PackedWaveGlow, param_generation


def fused_add_tanh_sigmoid_multiply(input_a: torch.Tensor, input_b: torch.Tensor, n_channels) -> torch.Tensor:
    """Reference glow.py:33-40 as a standalone op: tanh((a+b)[:, :C]) * sigmoid((a+b)[:, C:]) for [B, 2C, T] CUDA
    tensors (``n_channels`` may be an int or the reference's one-element IntTensor).  Inside the drop-in WaveGlow this
    is the epilogue of the gate GEMM; this entry exists for callers that use the function on its own."""
    if not input_a.is_cuda:
        raise RuntimeError("fused_add_tanh_sigmoid_multiply needs CUDA tensors on a B200; there is no CPU fallback")
    _lib.require_b200(input_a.device)
    c = int(n_channels[0]) if isinstance(n_channels, torch.Tensor) else int(n_channels)
    b, two_c, t = input_a.shape
    if two_c != 2 * c or input_b.shape != input_a.shape:
        raise ValueError(f"expected two [B, {2 * c}, T] tensors, got {tuple(input_a.shape)} and {tuple(input_b.shape)}")
    out = torch.empty((b, c, t), device=input_a.device, dtype=torch.float32)
    with torch.cuda.device(input_a.device):
        _lib.call("wgb_fused_add_tanh_sigmoid_multiply", input_a.float().contiguous(), input_b.float().contiguous(), out,
                  b, c, t, _lib.stream_ptr())
    return out.to(input_a.dtype)


class WaveGlowLoss(torch.nn.Module):
    """Reference glow.py:43-59 (training-side scalar; plain tensor ops on whatever device z lives)."""

    def __init__(self, sigma: float = 1.0):
        super().__init__()
        self.sigma = sigma

    def forward(self, model_output):
        z, log_s_list, log_det_W_list = model_output
        log_s_total = sum(torch.sum(ls) for ls in log_s_list)
        log_det_total = sum(log_det_W_list)
        loss = torch.sum(z * z) / (2 * self.sigma * self.sigma) - log_s_total - log_det_total
        return loss / (z.size(0) * z.size(1) * z.size(2))


class Invertible1x1Conv(torch.nn.Module):
    """Reference glow.py:62-102: c x c channel mix, orthonormal init with det +1."""

    def __init__(self, c: int):
        super().__init__()
        self.conv = torch.nn.Conv1d(c, c, kernel_size=1, stride=1, padding=0, bias=False)
        w = torch.linalg.qr(torch.randn(c, c))[0]
        if torch.det(w) < 0:
            w[:, 0] = -w[:, 0]
        self.conv.weight.data = w.view(c, c, 1).contiguous()

    def forward(self, z: torch.Tensor, reverse: bool = False):
        """z [B, c, T] on a CUDA device.  reverse=False returns (W z, B*T*log|det W|)."""
        from .packing import pack_mix
        _lib.require_b200(z.device)
        b, c, t = z.shape
        fwd, inv, logdet = pack_mix(self.conv.weight.detach().float().cpu())
        x = torch.zeros((b, t, 8), device=z.device, dtype=torch.float32)
        x[:, :, 8 - c:] = z.float().permute(0, 2, 1)
        with torch.cuda.device(z.device):
            _lib.call("wgb_flow_mix", x, (inv if reverse else fwd).to(z.device), b * t, c, _lib.stream_ptr())
        out = x[:, :, 8 - c:].permute(0, 2, 1).contiguous().to(z.dtype)
        if reverse:
            return out
        return out, torch.tensor(b * t * logdet, device=z.device, dtype=torch.float32)


class WN(torch.nn.Module):
    """Reference glow.py:105-175.  Holds the parameters; computation is driven by the owning WaveGlow
    (``WN.forward`` works standalone only once attached to one, because the kernels are packed per model)."""

    def __init__(self, n_in_channels, n_mel_channels, n_layers, n_channels, kernel_size):
        super().__init__()
        assert kernel_size % 2 == 1
        assert n_channels % 2 == 0
        self.n_layers = n_layers
        self.n_channels = n_channels
        self.in_layers = torch.nn.ModuleList()
        self.res_skip_layers = torch.nn.ModuleList()
        self.cond_layers = torch.nn.ModuleList()
        wn = torch.nn.utils.weight_norm
        self.start = wn(torch.nn.Conv1d(n_in_channels, n_channels, 1), name="weight")
        end = torch.nn.Conv1d(n_channels, 2 * n_in_channels, 1)
        end.weight.data.zero_()                     # glow.py:128-131: couplings start as identity
        end.bias.data.zero_()
        self.end = end
        for i in range(n_layers):
            dilation = 2 ** i
            padding = (kernel_size * dilation - dilation) // 2
            self.in_layers.append(wn(torch.nn.Conv1d(n_channels, 2 * n_channels, kernel_size, dilation=dilation,
                                                     padding=padding), name="weight"))
            self.cond_layers.append(wn(torch.nn.Conv1d(n_mel_channels, 2 * n_channels, 1), name="weight"))
            rs = 2 * n_channels if i < n_layers - 1 else n_channels
            self.res_skip_layers.append(wn(torch.nn.Conv1d(n_channels, rs, 1), name="weight"))
        self._owner = None          # (WaveGlow, flow index), set by WaveGlow.__init__ / _attach

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_owner"] = None      # back-reference to the owning model: re-attached by WaveGlow.__setstate__
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self.__dict__.setdefault("_owner", None)

    def forward(self, forward_input):
        audio, spect = forward_input
        if self._owner is None:
            raise RuntimeError("WN.forward needs its owning WaveGlow (kernels are packed per model)")
        model, k = self._owner
        with torch.cuda.device(audio.device), torch.no_grad():
            pk = model._packed(audio.device)
            return engine.wn_standalone(pk, k, audio.float().contiguous(), spect.float().contiguous())


class WaveGlow(torch.nn.Module):
    """Reference glow.py:178-302."""

    def __init__(self, n_mel_channels, n_flows, n_group, n_early_every, n_early_size, WN_config):
        super().__init__()
        self.upsample = torch.nn.ConvTranspose1d(n_mel_channels, n_mel_channels, 1024, stride=256)
        assert n_group % 2 == 0
        self.n_flows = n_flows
        self.n_group = n_group
        self.n_early_every = n_early_every
        self.n_early_size = n_early_size
        self.WN = torch.nn.ModuleList()
        self.convinv = torch.nn.ModuleList()
        n_half = n_group // 2
        n_remaining_channels = n_group
        for k in range(n_flows):
            if k % self.n_early_every == 0 and k > 0:
                n_half = n_half - self.n_early_size // 2
                n_remaining_channels = n_remaining_channels - self.n_early_size
            self.convinv.append(Invertible1x1Conv(n_remaining_channels))
            self.WN.append(WN(n_half, n_mel_channels * n_group, **WN_config))
        self.n_remaining_channels = n_remaining_channels
        self.mode = "bf16"
        # 'auto' | 'mel' | 'cond': how the bf16 gate GEMM gets its conditioning (engine.use_mel_path)
        self.cond_path = "auto"
        self._pack_cache = {}
        self._attach()

    # ------------------------------------------------------------------ plumbing
    def _attach(self):
        for k, wn in enumerate(self.WN):
            object.__setattr__(wn, "_owner", (self, k))

    def __getstate__(self):
        # torch.save(model) / copy.deepcopy(model) carry parameters only: the packed-weight cache (2+ GB of derived
        # device tensors) is rebuilt on first use; WN._owner is restored by _attach.  weight_norm's cached non-leaf
        # ``weight`` attributes are detached first (current torch refuses to deepcopy them; the kernels never read them)
        from .convert_model import detach_weight_norm_cache
        detach_weight_norm_cache(self)
        state = self.__dict__.copy()
        state["_pack_cache"] = {}
        return state

    def __setstate__(self, state):          # pickled-module checkpoints (waveglow/inference.py:37)
        super().__setstate__(state)
        self.__dict__.setdefault("mode", "bf16")
        self.__dict__.setdefault("cond_path", "auto")
        self.__dict__["_pack_cache"] = {}
        self._attach()

    def _check_supported(self):
        wn0 = self.WN[0]
        k = self.upsample.kernel_size[0]
        if self.n_group != 8 or self.upsample.stride[0] * 4 != k:
            raise RuntimeError("kernels are specialised for n_group = 8 and upsample kernel = 4 * stride")
        if self.mode == "bf16" and (wn0.n_channels != 512 or wn0.n_layers != 8 or
                                    wn0.cond_layers[0].in_channels != 640 or
                                    wn0.in_layers[0].kernel_size[0] != 3):
            raise RuntimeError("bf16 tensor-core path is specialised for WN 8 layers x 512 ch, k=3, 80 mels x 8 "
                               "(config.json); use mode='fp32' for other shapes")

    def _signature(self):
        # data_ptr / _version catch torch-side updates; the generation counter catches raw-kernel updates (FusedAdam
        # re-points p.data at a flat buffer and steps it with wgb_adam_step_dev, which changes neither)
        return (param_generation(),) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _packed(self, device) -> PackedWaveGlow:
        _lib.require_b200(device)
        self._check_supported()
        key = (self.mode, str(device))
        sig = self._signature()
        hit = self._pack_cache.get(key)
        if hit is None or hit[0] != sig:
            wn0 = self.WN[0]
            pk = PackedWaveGlow(self.state_dict(), self.n_flows, wn0.n_layers, wn0.n_channels, self.n_group,
                                self.mode, device)
            self._pack_cache = {key: (sig, pk)}
            hit = self._pack_cache[key]
        hit[1].cond_path = self.cond_path
        return hit[1]

    def repack(self):
        """Drop packed weights (call after mutating parameters in ways version counters miss)."""
        self._pack_cache = {}

    # ------------------------------------------------------------------ reference API
    def forward(self, forward_input):
        """forward_input = (mel [B, n_mel, frames], audio [B, time]) -> (z, log_s_list, log_det_W_list)."""
        spect, audio = forward_input
        if not spect.is_cuda:
            raise RuntimeError("WaveGlow.forward needs CUDA tensors on a B200; there is no CPU fallback")
        if self.training and torch.is_grad_enabled():
            # training direction (waveglow/train.py:116-124): the same outputs, wired into autograd so that
            # criterion(outputs).backward() fills every parameter's .grad through this package's backward kernels
            from . import training
            return training.forward_autograd(self, spect, audio)
        with torch.cuda.device(spect.device), torch.no_grad():      # kernels launch on the tensors' device / its current stream
            pk = self._packed(spect.device)
            return engine.forward(pk, spect.float().contiguous(), audio.float().contiguous())

    def infer(self, spect: torch.Tensor, sigma: float = 1.0, z: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not spect.is_cuda:
            raise RuntimeError("WaveGlow.infer needs CUDA tensors on a B200; there is no CPU fallback")
        b, _, f = spect.shape
        t = f * self.upsample.stride[0] // self.n_group
        if z is None:
            z = torch.randn((b, self.n_group, t), device=spect.device, dtype=torch.float32)
        if tuple(z.shape) != (b, self.n_group, t):
            raise ValueError(f"z must be [{b}, {self.n_group}, {t}], got {tuple(z.shape)}")
        with torch.cuda.device(spect.device), torch.no_grad():      # kernels launch on the tensors' device / its current stream
            pk = self._packed(spect.device)
            audio = engine.infer(pk, spect.float().contiguous(), z.to(spect.device).float().contiguous(), sigma)
        return audio.to(spect.dtype) if spect.dtype in (torch.float16, torch.bfloat16) else audio

    def graphed_infer(self, batch: int, frames: int, sigma: float = 1.0, before_capture=None):
        """CUDA-graph replay of ``infer`` for a fixed shape (the ~210 kernel launches of one call become one graph
        launch: worth ~7 % at batch 1 and a few % at batch 8, where launch gaps show).  Returns ``run(spect, z) ->
        audio``; spect / z (device or pinned host tensors) are copied into the graph's static buffers, the returned
        tensor is the graph's static output (clone it to keep it).  ``run.replay()`` replays on whatever the static
        buffers hold.  ``before_capture`` (optional callable) runs after the eager warm-up, right before the capture
        (bench.py arms its launch counter / event-record nodes there).  The captured graph holds the weights as packed
        at capture time: re-create it after changing parameters."""
        device = self.upsample.weight.device
        pk = self._packed(device)
        t = frames * self.upsample.stride[0] // self.n_group
        spect_buf = torch.zeros((batch, self.upsample.in_channels, frames), device=device, dtype=torch.float32)
        z_buf = torch.zeros((batch, self.n_group, t), device=device, dtype=torch.float32)
        with torch.cuda.device(device):
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side), torch.no_grad():
                engine.infer(pk, spect_buf, z_buf, sigma)             # warm-up outside the capture (allocator, func attrs)
            torch.cuda.current_stream(device).wait_stream(side)
            if before_capture is not None:
                before_capture()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph), torch.no_grad():
                out = engine.infer(pk, spect_buf, z_buf, sigma)

        def run(spect: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
            spect_buf.copy_(spect, non_blocking=True)
            z_buf.copy_(z, non_blocking=True)
            graph.replay()
            return out

        run.graph = graph
        run.replay = graph.replay
        run.output = out
        return run

    @staticmethod
    def remove_weightnorm(model):
        waveglow = model
        for wn in waveglow.WN:
            wn.start = torch.nn.utils.remove_weight_norm(wn.start)
            wn.in_layers = remove(wn.in_layers)
            wn.cond_layers = remove(wn.cond_layers)
            wn.res_skip_layers = remove(wn.res_skip_layers)
        waveglow.repack()
        return waveglow


def remove(conv_list):
    new_conv_list = torch.nn.ModuleList()
    for old_conv in conv_list:
        new_conv_list.append(torch.nn.utils.remove_weight_norm(old_conv))
    return new_conv_list
