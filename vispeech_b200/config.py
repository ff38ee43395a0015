"""JSON config -> attribute dict, mirroring reference utils.get_hparams_from_file / HParams (utils.py:218-224, 281-310)."""
from __future__ import annotations

import json
import os

DEFAULT_CONFIG_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "config.json")


class HParams:
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            if isinstance(v, dict):
                v = HParams(**v)
            self[k] = v

    def keys(self):
        return self.__dict__.keys()

    def items(self):
        return self.__dict__.items()

    def values(self):
        return self.__dict__.values()

    def __len__(self):
        return len(self.__dict__)

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        return setattr(self, key, value)

    def __contains__(self, key):
        return key in self.__dict__

    def __repr__(self):
        return self.__dict__.__repr__()


def get_hparams_from_file(config_path: str = DEFAULT_CONFIG_PATH) -> HParams:
    with open(config_path, "r") as f:
        return HParams(**json.loads(f.read()))


N_SYMBOLS = 519   # len(text/symbols.py:22): "_" + 401 zh + 42 ja + 69 en + 6 punctuation


def build_from_hparams(hps: HParams, device=None):
    """What inference.py:26-34 does: SynthesizerTrn(len(symbols), filter_length//2+1, hop, sr, segment//hop, ...)."""
    from .synthesizer import SynthesizerTrn
    return SynthesizerTrn(N_SYMBOLS, hps.data.filter_length // 2 + 1, hps.data.hop_length, hps.data.sampling_rate,
                          hps.train.segment_size // hps.data.hop_length, n_speakers=hps.data.n_speakers,
                          device=device, **hps.model).eval()
