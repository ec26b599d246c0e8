"""Golden output of the reference Postnet (tacotron/modules.py:94-137), imported read-only by file path.

    python tests/golden/make_golden_postnet.py        (build container only)

`tacotron/__init__.py` pulls in the whole Tacotron-2 (hparams, text front end ...), so modules.py is loaded on its own:
it only needs torch / numpy.  Weights and input come from seeds (text2speech_b200.synthetic); only the output is stored.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from text2speech_b200 import synthetic as syn       # noqa: E402

REF = os.environ.get("T2S_REFERENCE_ROOT", "/root/reference")


def main():
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_tacotron_modules", os.path.join(REF, "tacotron", "modules.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    hp = syn.DEFAULT_POSTNET_HPARAMS
    model = mod.Postnet(hp)
    missing = model.load_state_dict(syn.synthetic_postnet_state_dict(hp, seed=77), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.eval()
    mel = syn.synthetic_mel(2, 37, seed=3)
    with torch.no_grad():
        out = model(mel)
    path = os.path.join(HERE, "postnet_golden.npz")
    np.savez_compressed(path, postnet_out=out.numpy())
    print(path, os.path.getsize(path) // 1024, "KiB", tuple(out.shape), "absmax", float(out.abs().max()), "std", float(out.std()))


if __name__ == "__main__":
    main()
