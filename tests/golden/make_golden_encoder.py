"""Golden output of the conv bank of the reference Tacotron-2 Encoder (tacotron/tacotron.py:167-217), build container only.

    python tests/golden/make_golden_encoder.py

The UNMODIFIED ``Encoder.inference`` runs end to end; a forward pre-hook on its LSTM captures the tensor the conv bank
hands over (tacotron.py:214-217).  ``tacotron/__init__.py`` pulls in the text front end, so the three files are loaded
by path under a scratch package name, with ``utils.data_utils`` (librosa / data loaders) stubbed: only ``to_gpu`` is
imported from it and the Encoder never calls it.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from text2speech_b200 import synthetic as syn       # noqa: E402

REF = os.environ.get("T2S_REFERENCE_ROOT", "/root/reference")


def load_reference_tacotron():
    sys.dont_write_bytecode = True
    pkg = types.ModuleType("ref_tacotron")
    pkg.__path__ = [os.path.join(REF, "tacotron")]
    sys.modules["ref_tacotron"] = pkg
    stub_utils = types.ModuleType("utils")
    stub_utils.__path__ = []
    stub_du = types.ModuleType("utils.data_utils")
    stub_du.to_gpu = lambda x: x
    sys.modules.setdefault("utils", stub_utils)
    sys.modules["utils.data_utils"] = stub_du
    if REF not in sys.path:
        sys.path.insert(0, REF)                      # hparams.py
    mods = {}
    for name in ("modules", "attention", "tacotron"):
        spec = importlib.util.spec_from_file_location(f"ref_tacotron.{name}", os.path.join(REF, "tacotron", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"ref_tacotron.{name}"] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["tacotron"]


def encoder_input(bsz=2, steps=41, channels=512, seed=21):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((bsz, channels, steps), generator=g)


def main():
    ref = load_reference_tacotron()
    enc = ref.Encoder()
    sd = syn.synthetic_encoder_convs_state_dict(seed=78)
    res = enc.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all(k.startswith("lstm.") for k in res.missing_keys), res
    enc.eval()
    captured = {}
    enc.lstm.register_forward_pre_hook(lambda mod, inp: captured.__setitem__("x", inp[0].detach().clone()))
    with torch.no_grad():
        enc.inference(encoder_input())
    out = captured["x"]                               # [B, T, 512]
    path = os.path.join(HERE, "encoder_golden.npz")
    np.savez_compressed(path, conv_bank_out=out.numpy())
    print(path, os.path.getsize(path) // 1024, "KiB", tuple(out.shape), "std", float(out.std()))


if __name__ == "__main__":
    main()
