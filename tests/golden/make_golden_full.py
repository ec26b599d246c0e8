"""Golden output of the UNMODIFIED reference for one full-size utterance (BASELINE configs[1]: 80x860 mel, 10 s).

    python tests/golden/make_golden_full.py       (build container only; ~1 min of CPU)

Stores the reference's WaveGlow.infer audio (fp32, 220 160 samples) for the 'bench' weight recipe, mel seed 0 and
noise seed 2024 (the inputs bench.py uses for utterance 0), so the GPU path is checked end to end at full length
against the reference itself rather than against the oracle port.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import ref_harness as rh          # noqa: E402
from tests.golden.make_golden import flow_channels, RECIPES, SIGMA   # noqa: E402
from text2speech_b200 import synthetic as syn       # noqa: E402


def main():
    warnings.simplefilter("ignore")
    torch.set_num_threads(os.cpu_count() or 1)
    ref_glow, _, _, _, feed = rh.load()
    cfg = syn.load_config()
    sd = syn.synthetic_state_dict(cfg, **RECIPES["bench"])
    model = rh.build_reference_waveglow(ref_glow, cfg, sd, weight_norm=False)
    mel = syn.synthetic_mel(1, 860, seed=0)
    z = syn.synthetic_z(1, 860, seed=2024)
    feed.load(z, flow_channels(cfg))
    with torch.no_grad():
        audio = model.infer(mel, sigma=SIGMA)
    assert audio.shape == (1, 220160) and torch.isfinite(audio).all()
    path = os.path.join(HERE, "full_utterance_golden.npz")
    np.savez_compressed(path, bench_full_infer_audio=audio.numpy())
    print(path, os.path.getsize(path) // 1024, "KiB", "absmax", float(audio.abs().max()), "std", float(audio.std()))


if __name__ == "__main__":
    main()
