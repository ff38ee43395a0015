"""JSON config -> attribute dict, mirroring reference utils.get_hparams_from_file / HParams (utils.py:218-224, 281-310)."""
from __future__ import annotations

import json
import os

DEFAULT_CONFIG_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "config.json")


class HParams(dict):
    """Nested config with both `hps.data.hop_length` and `hps["data"]["hop_length"]` access - what the reference scripts
    use from utils.HParams (utils.py:281-310).  A dict underneath (keys / items / values / len / in come with it); nested
    dicts become HParams on the way in."""

    def __init__(self, **kwargs):
        super().__init__()
        for k, v in kwargs.items():
            self[k] = v

    def __setitem__(self, key, value):
        super().__setitem__(key, HParams(**value) if isinstance(value, dict) and not isinstance(value, HParams) else value)

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key) from None

    __setattr__ = __setitem__


def get_hparams_from_file(config_path: str = DEFAULT_CONFIG_PATH) -> HParams:
    with open(config_path, "r") as f:
        return HParams(**json.loads(f.read()))


from .text import symbols as _symbols

N_SYMBOLS = len(_symbols)   # 519 = len(text/symbols.py:22): "_" + 401 zh + 42 ja + 69 en + 6 punctuation


def build_from_hparams(hps: HParams, device=None):
    """What inference.py:26-34 does: SynthesizerTrn(len(symbols), filter_length//2+1, hop, sr, segment//hop, ...)."""
    from .synthesizer import SynthesizerTrn
    return SynthesizerTrn(N_SYMBOLS, hps.data.filter_length // 2 + 1, hps.data.hop_length, hps.data.sampling_rate,
                          hps.train.segment_size // hps.data.hop_length, n_speakers=hps.data.n_speakers,
                          device=device, **hps.model).eval()
