"""text2speech_b200 — B200-native (sm_100a) vocoding path of DonggeunYu/Text2Speech.

Drop-in module API (same names / signatures / state_dict layout as the reference):
    WaveGlow, WN, Invertible1x1Conv, WaveGlowLoss   (waveglow/glow.py)
    Denoiser                                        (waveglow/denoiser.py)
    STFT                                            (utils/stft.py)
    TacotronSTFT                                    (utils/layers.py)
    Postnet, ConvNorm                               (tacotron/modules.py:94-137,177-197; inference only)
    EncoderConvs                                    (conv bank of tacotron/tacotron.py:175-186,211-212; inference only)
All compute runs in hand-written CUDA behind the C ABI in include/waveglow_b200.h; there is no CPU
fallback.  Importing the package does not need a GPU (the library is loaded on first use).
"""
from .glow import (WaveGlow, WN, Invertible1x1Conv, WaveGlowLoss, remove,      # noqa: F401
                   fused_add_tanh_sigmoid_multiply)
from .denoiser import Denoiser                                                # noqa: F401
from .stft import STFT                                                        # noqa: F401
from .layers import TacotronSTFT                                              # noqa: F401
from .postnet import Postnet, EncoderConvs, ConvNorm                          # noqa: F401

MAX_WAV_VALUE = 32768.0        # waveglow/mel2samp.py:40

__all__ = ["WaveGlow", "WN", "Invertible1x1Conv", "WaveGlowLoss", "Denoiser", "STFT", "TacotronSTFT", "Postnet", "EncoderConvs", "ConvNorm"]
