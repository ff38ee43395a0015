"""vispeech_b200 - B200-native (sm_100a) implementation of vispeech's `SynthesizerTrn.infer` hot path.

Host side: `SynthesizerTrn` (reference call surface) -> ctypes -> libvispeech_b200.so (include/vispeech_b200.h).
"""
from .config import HParams, build_from_hparams, get_hparams_from_file  # noqa: F401
from .synthesizer import SynthesizerTrn, load_checkpoint  # noqa: F401

from .mel import MelSpectrogram, mel_spectrogram_torch, spectrogram_torch  # noqa: F401

__all__ = ["SynthesizerTrn", "load_checkpoint", "get_hparams_from_file", "HParams", "build_from_hparams",
           "MelSpectrogram", "mel_spectrogram_torch", "spectrogram_torch"]
