"""Import the UNMODIFIED reference modules read-only from /root/reference (build container only).

Used by tests/golden/make_golden.py (fixture generation) and by the optional
``tests/test_reference_live.py`` checks that run only where /root/reference exists.  Nothing
here is copied from the reference: it only arranges for ``glow``, ``denoiser``, ``utils.stft``
and ``utils.layers`` to import and run on CPU:

  1. sys.path gets <ref>/waveglow (module ``glow``) and <ref> (package ``utils``);
  2. a tiny ``librosa`` stand-in (librosa is not installed; the reference pins 0.6.0):
     ``util.pad_center``, ``util.tiny``, ``util.normalize(norm=None)`` and ``filters.mel``,
     the latter backed by torchaudio's independent Slaney filterbank;
  3. ``torch.cuda.FloatTensor(...).normal_()`` (glow.py:260-267, :285-288) is replaced by a
     factory that hands out host-supplied noise slices in call order;
  4. ``Tensor.cuda`` / ``Module.cuda`` become no-ops (stft.py:85-89, denoiser.py:15,36).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("T2S_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "waveglow", "glow.py"))


def _install_librosa_shim():
    if "librosa" in sys.modules and not getattr(sys.modules["librosa"], "_t2s_shim", False):
        return
    import torchaudio

    lib = types.ModuleType("librosa")
    lib._t2s_shim = True
    util = types.ModuleType("librosa.util")
    filters = types.ModuleType("librosa.filters")

    def pad_center(data, size, axis=-1, **kw):
        n = data.shape[axis]
        lpad = (size - n) // 2
        widths = [(0, 0)] * data.ndim
        widths[axis] = (lpad, size - n - lpad)
        return np.pad(data, widths, **kw)

    def tiny(x):
        x = np.asarray(x)
        dt = x.dtype if np.issubdtype(x.dtype, np.floating) else np.float32
        return np.finfo(dt).tiny

    def normalize(s, norm=None, **kw):
        assert norm is None
        return s

    def mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
        fmax = sr / 2.0 if fmax is None else fmax
        fb = torchaudio.functional.melscale_fbanks(
            n_freqs=n_fft // 2 + 1, f_min=float(fmin), f_max=float(fmax), n_mels=n_mels,
            sample_rate=int(sr), norm="slaney", mel_scale="slaney")
        return fb.T.contiguous().numpy()

    util.pad_center, util.tiny, util.normalize = pad_center, tiny, normalize
    filters.mel = mel
    lib.util, lib.filters = util, filters
    sys.modules["librosa"] = lib
    sys.modules["librosa.util"] = util
    sys.modules["librosa.filters"] = filters


class NoiseFeed:
    """Stand-in for ``torch.cuda.FloatTensor``: ``feed(*shape).normal_()`` pops the next slice."""

    def __init__(self):
        self.queue = []

    def load(self, z: torch.Tensor, chans):
        """Queue the draws ``WaveGlow.infer`` will make for host noise z [B, n_group, T]."""
        n_group = z.shape[1]
        self.queue = [z[:, n_group - chans[-1]:].clone()]
        for k in reversed(range(1, len(chans))):
            if chans[k - 1] > chans[k]:
                lo = n_group - chans[k - 1]
                self.queue.append(z[:, lo: lo + chans[k - 1] - chans[k]].clone())

    def __call__(self, *shape):
        feed = self

        class _Draw:
            def normal_(self_inner):
                if not feed.queue:      # nothing loaded: Denoiser.__init__ draws with sigma=0
                    return torch.zeros(*shape)
                nxt = feed.queue.pop(0)
                assert tuple(nxt.shape) == tuple(shape), (nxt.shape, shape)
                return nxt

        return _Draw()


_FEED = NoiseFeed()


def load():
    """Returns (glow module, denoiser module, utils.stft module, utils.layers module, noise feed)."""
    assert available(), f"reference not found under {REF_ROOT}"
    sys.dont_write_bytecode = True
    for p in (os.path.join(REF_ROOT, "waveglow"), REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    _install_librosa_shim()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import glow as ref_glow                      # noqa
        import utils.stft as ref_stft                # noqa
        import utils.layers as ref_layers            # noqa
        import denoiser as ref_denoiser              # noqa
    torch.cuda.FloatTensor = _FEED
    return ref_glow, ref_denoiser, ref_stft, ref_layers, _FEED


@contextlib.contextmanager
def cpu_cuda_noop():
    """Make ``.cuda()`` a no-op so stft.py:85-89 / denoiser.py:15,36 run without a GPU."""
    t_cuda, m_cuda = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = t_cuda, m_cuda


def build_reference_waveglow(ref_glow, config, state_dict, weight_norm: bool):
    """Construct the reference WaveGlow and load ``state_dict`` (folded or weight-norm layout)."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = ref_glow.WaveGlow(**config)
        if not weight_norm:
            model = ref_glow.WaveGlow.remove_weightnorm(model)
    missing = model.load_state_dict(state_dict, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model.eval()
