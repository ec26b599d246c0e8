"""CPU oracle for the WaveGlow vocoding path — TEST INFRASTRUCTURE ONLY.

This package is a plain fp32/fp64 PyTorch-on-CPU restatement of the reference's
algorithm for the hot path (WaveGlow.infer / WaveGlow.forward, Denoiser, the
conv-basis STFT and TacotronSTFT.mel_spectrogram).  Every function cites the
reference file:line it follows (paths relative to the reference checkout).

Rules (enforced by tests/test_layout.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
    ``--impl reference`` legs may import it, and only as the checker / the timed
    CPU baseline — never as part of the product path;
  * ``text2speech_b200`` never imports ``oracle``; the product path raises if the
    CUDA library is missing instead of falling back to anything in here.

Parity pin: the reference is pure Python and runs in the build container, so the
oracle is pinned against outputs of the UNMODIFIED reference modules
(``tests/golden/make_golden.py`` imports them read-only from /root/reference and
writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks the oracle
against those files).  The one exception is the mel filterbank: the reference
gets it from ``librosa==0.6.0`` (waveglow/requirements.txt:5), which is not
vendored in the reference tree nor installed here -> ``mel_filterbank`` is a
restatement of librosa 0.6's published Slaney-scale / Slaney-norm algorithm,
cross-checked against torchaudio's independent implementation:
"parity unpinned" for ``mel_basis`` only.
"""
from .waveglow_oracle import (  # noqa: F401
    fold_weight_norm, folded_state, wn_layer, wn_stack, waveglow_infer, waveglow_forward,
    regroup_spect, upsample_spect, flow_channels,
)
from .postnet_oracle import postnet, encoder_convs  # noqa: F401
from .stft_oracle import (  # noqa: F401
    stft_bases, stft_transform, stft_inverse, window_sumsquare, mel_filterbank,
    mel_spectrogram, denoiser_bias_spec, denoise, griffin_lim, griffin_lim_initial_angles, pcm16,
)
