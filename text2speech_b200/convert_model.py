"""Checkpoint migration for old WaveGlow pickles (reference: waveglow/convert_model.py:11-38).

Early WaveGlow checkpoints keep the residual and skip 1x1 convs of every WN layer as two
ModuleLists (``res_layers`` with n_layers-1 entries, ``skip_layers`` with n_layers); the current
layout stacks them into one ``res_skip_layers[i]`` whose rows are [res ; skip] (last layer: skip
only).  ``update_model`` performs that fusion and re-applies weight norm, so the result has the
state_dict keys the packing step expects.

    python -m text2speech_b200.convert_model old_checkpoint.pt new_checkpoint.pt
"""
from __future__ import annotations

import copy
import sys

import torch


def _check_model_old_version(model) -> bool:
    return hasattr(model.WN[0], "res_layers")


def _plain(conv):
    """weight / bias of a 1x1 conv whether or not it still carries weight norm."""
    if hasattr(conv, "weight_g"):
        conv = torch.nn.utils.remove_weight_norm(conv)
    return conv.weight.detach(), conv.bias.detach()


def fuse_res_skip(wavenet) -> torch.nn.ModuleList:
    """New-layout ``res_skip_layers`` (weight-normed Conv1d, rows [res ; skip]) of one old-layout WN."""
    n_channels, n_layers = wavenet.n_channels, wavenet.n_layers
    fused = torch.nn.ModuleList()
    for i in range(n_layers):
        w_skip, b_skip = _plain(wavenet.skip_layers[i])
        if i < n_layers - 1:
            w_res, b_res = _plain(wavenet.res_layers[i])
            weight, bias = torch.cat([w_res, w_skip]), torch.cat([b_res, b_skip])
        else:
            weight, bias = w_skip, b_skip
        layer = torch.nn.Conv1d(n_channels, weight.shape[0], 1)
        layer.weight = torch.nn.Parameter(weight.clone())
        layer.bias = torch.nn.Parameter(bias.clone())
        fused.append(torch.nn.utils.weight_norm(layer, name="weight"))
    return fused


def detach_weight_norm_cache(model) -> None:
    """weight_norm keeps the last computed ``weight`` as a plain (non-leaf) attribute, which current torch
    refuses to deepcopy; replace it by a detached tensor (the pre-forward hook recomputes it anyway)."""
    for m in model.modules():
        for hook in m._forward_pre_hooks.values():
            name = getattr(hook, "name", None)
            if name is not None and isinstance(m.__dict__.get(name), torch.Tensor):
                m.__dict__[name] = m.__dict__[name].detach()


def update_model(old_model):
    """Old res/skip-split model -> current layout (deep copy); current-layout models pass through."""
    if not _check_model_old_version(old_model):
        return old_model
    detach_weight_norm_cache(old_model)
    new_model = copy.deepcopy(old_model)
    for wavenet in new_model.WN:
        wavenet.res_skip_layers = fuse_res_skip(wavenet)
        del wavenet.res_layers
        del wavenet.skip_layers
    if hasattr(new_model, "repack"):
        new_model.repack()
    return new_model


if __name__ == "__main__":
    from .inference import install_glow_alias
    install_glow_alias()
    old_model_path, new_model_path = sys.argv[1], sys.argv[2]
    model = torch.load(old_model_path, map_location="cpu", weights_only=False)
    model["model"] = update_model(model["model"])
    torch.save(model, new_model_path)
