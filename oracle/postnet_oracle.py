"""CPU oracle: Tacotron-2 Postnet in eval mode (TEST INFRASTRUCTURE).

Restates ``tacotron/modules.py:94-137`` of the reference as a plain function on a ``state_dict``:
five Conv1d (kernel 5, "same" zero padding) each followed by BatchNorm1d with running statistics
(``torch.nn.BatchNorm1d`` eval semantics, eps 1e-5), tanh after all but the last; dropout is the
identity in eval mode (``F.dropout(x, 0.5, self.training)``, modules.py:134-135).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def postnet(state: Dict[str, Tensor], x: Tensor, eps: float = 1e-5) -> Tensor:
    """x [B, n_mel, F] -> [B, n_mel, F]."""
    n = 1 + max(int(k.split(".")[1]) for k in state if k.startswith("convolutions."))
    for i in range(n):
        p = f"convolutions.{i}."
        w, b = state[p + "0.conv.weight"], state[p + "0.conv.bias"]
        x = F.conv1d(x, w, b, padding=(w.shape[2] - 1) // 2)                                  # modules.py:105-109
        x = F.batch_norm(x, state[p + "1.running_mean"], state[p + "1.running_var"], state[p + "1.weight"],
                         state[p + "1.bias"], training=False, eps=eps)                         # modules.py:110
        if i < n - 1:
            x = torch.tanh(x)                                                                  # modules.py:134
    return x


def encoder_convs(state: Dict[str, Tensor], x: Tensor, eps: float = 1e-5) -> Tensor:
    """Conv bank of the Tacotron-2 Encoder in eval mode (tacotron/tacotron.py:175-186 built, :211-212 run):
    x [B, C, T] -> relu(batch_norm(conv(x))) per layer -> [B, C, T] (what is transposed and fed to the LSTM at :214-217)."""
    n = 1 + max(int(k.split(".")[1]) for k in state if k.startswith("convolutions."))
    for i in range(n):
        p = f"convolutions.{i}."
        w, b = state[p + "0.conv.weight"], state[p + "0.conv.bias"]
        x = F.conv1d(x, w, b, padding=(w.shape[2] - 1) // 2)
        x = F.batch_norm(x, state[p + "1.running_mean"], state[p + "1.running_var"], state[p + "1.weight"],
                         state[p + "1.bias"], training=False, eps=eps)
        x = torch.relu(x)                                                                      # tacotron.py:212
    return x
