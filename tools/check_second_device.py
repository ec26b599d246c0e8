# runs on a 2-GPU box: the model lives on cuda:1 while cuda:0 is the current device
import sys, torch
sys.path.insert(0, '/root/repo')
import text2speech_b200 as t2s
from text2speech_b200 import synthetic as syn
assert torch.cuda.current_device() == 0
cfg = syn.load_config()
m = t2s.WaveGlow.remove_weightnorm(t2s.WaveGlow(**cfg))
m.load_state_dict(syn.synthetic_state_dict(cfg, seed=1234, end_std=0.01))
mel, z = syn.synthetic_mel(2, 40, seed=1), syn.synthetic_z(2, 40, seed=2)
outs = []
for dev in ("cuda:0", "cuda:1"):
    mm = m.to(dev).eval()
    outs.append(mm.infer(mel.to(dev), sigma=0.666, z=z.to(dev)).cpu())
    taco = t2s.TacotronSTFT(1024, 256, 1024, 80, 22050, 0.0, 8000.0).to(dev)
    y = syn.synthetic_waveforms(2, 8192).to(dev)
    outs.append(taco.mel_spectrogram(y).cpu())
    outs.append(t2s.Denoiser(mm)(y, 0.1).cpu())
assert all(torch.equal(a, b) for a, b in zip(outs[:3], outs[3:])), "cuda:1 result differs from cuda:0"
print("non-current device OK")
