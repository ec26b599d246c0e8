"""2+ GPU check of MultiGpuVocoder: sharded result == single-GPU result, bit for bit (run with gpurun --gpus 2)."""
import sys
import time

import torch

sys.path.insert(0, "/root/repo")
import text2speech_b200 as t2s                                    # noqa: E402
from text2speech_b200 import synthetic as syn                     # noqa: E402
from text2speech_b200.sharding import MultiGpuVocoder             # noqa: E402

cfg = syn.load_config()
m = t2s.WaveGlow.remove_weightnorm(t2s.WaveGlow(**cfg))
m.load_state_dict(syn.synthetic_state_dict(cfg, seed=1234, end_std=0.01))
B, F = 7, 200                                                     # ragged split over the devices
mel, z = syn.synthetic_mel(B, F, seed=1), syn.synthetic_z(B, F, seed=2)
single = m.to("cuda:0").eval().infer(mel.to("cuda:0"), sigma=0.666, z=z.to("cuda:0")).cpu()
voc = MultiGpuVocoder(m)
t0 = time.perf_counter()
multi = voc.infer(mel, sigma=0.666, z=z)
print(f"{len(voc.devices)} devices, {time.perf_counter() - t0:.3f} s")
assert multi.shape == single.shape and torch.equal(multi, single), "sharded result differs"
print("MultiGpuVocoder OK")
