"""Seeded synthetic inputs for BASELINE.json's configs.  Test/bench infrastructure.

Every generator returns a list of per-utterance dicts:
  ids int64 [Tp], sid int, duration (int64 or float32 [Tp]) or None, f0 (Hz, [Tp]) or None,
  energy ([Tp]) or None.
Noise is drawn separately with `draw_noise` once the frame counts are known.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

ZH_LO, ZH_HI = 1, 401          # zh block of text/symbols.py:16-22
N_SPK_USED = 67                # configs/config.json:37 lists 67 speakers
HOP, SR = 512, 44100


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _randint(g, lo, hi, n):
    return torch.randint(lo, hi + 1, (n,), generator=g)


def c1(seed: int = 0) -> List[Dict]:
    """C1: B=1, 40 zh phonemes, MFA-style integer durations U{3..18}, sid 0."""
    g = _gen(seed)
    return [dict(ids=_randint(g, ZH_LO, ZH_HI, 40), sid=0, duration=_randint(g, 3, 18, 40), f0=None, energy=None)]


def c2(batch: int = 64, seed: int = 1, tp: int = 40, target_frames: int = 431) -> List[Dict]:
    """C2: `batch` utterances of ~5 s: durations U{3..18} rescaled so sum = 431 +- 20 frames."""
    g = _gen(seed)
    out = []
    for _ in range(batch):
        ids = _randint(g, ZH_LO, ZH_HI, tp)
        d = _randint(g, 3, 18, tp).double()
        tgt = target_frames + int(_randint(g, -20, 20, 1))
        d = torch.clamp(torch.round(d * tgt / d.sum()), min=1).long()
        d[-1] = max(1, int(d[-1]) + tgt - int(d.sum()))
        out.append(dict(ids=ids, sid=int(_randint(g, 0, N_SPK_USED - 1, 1)), duration=d, f0=None, energy=None))
    return out


def _fixture_like_durations(g, tp: int, total: int) -> torch.Tensor:
    """Durations with the shape of filelists/train.list (median 7, long tail, 1.3 % zeros), summing to `total`."""
    base = torch.exp(torch.randn(tp, generator=g) * 0.8 + 1.95)
    tail = torch.rand(tp, generator=g) < 0.03
    base = torch.where(tail, base * 8, base)
    zero = torch.rand(tp, generator=g) < 0.013
    base = torch.where(zero, torch.zeros(()), base)
    if float(base.sum()) <= 0:
        base = torch.ones(tp)
    d = torch.floor(base * total / base.sum()).long()
    d = torch.where(zero, torch.zeros((), dtype=torch.long), d)
    rest = total - int(d.sum())
    nz = (~zero).nonzero().reshape(-1)
    if nz.numel() == 0:
        nz = torch.arange(tp)
    d[nz[0]] += rest
    return d


def c3(batch: int = 512, seed: int = 2) -> List[Dict]:
    """C3: mixed lengths 1-15 s (Tf 86..1292), Tp = round(Tf/10.5), fixture-like duration histogram."""
    g = _gen(seed)
    out = []
    for _ in range(batch):
        secs = 1.0 + 14.0 * float(torch.rand(1, generator=g))
        tf = int(round(secs * SR / HOP))
        tp = max(2, int(round(tf / 10.5)))
        out.append(dict(ids=_randint(g, ZH_LO, ZH_HI, tp), sid=int(_randint(g, 0, N_SPK_USED - 1, 1)),
                        duration=_fixture_like_durations(g, tp, tf), f0=None, energy=None))
    return out


def c4(seed: int = 3, tf: int = 5168, tp: int = 492) -> List[Dict]:
    """C4: one 60 s utterance."""
    g = _gen(seed)
    return [dict(ids=_randint(g, ZH_LO, ZH_HI, tp), sid=int(_randint(g, 0, N_SPK_USED - 1, 1)),
                 duration=_fixture_like_durations(g, tp, tf), f0=None, energy=None)]


def c5(batch: int = 256, seed: int = 4) -> List[Dict]:
    """C5: manual-edit path; jittered durations in mixed dtypes (ints, fractional floats, zeros,
    negatives, a few very long), F0 in Hz with unvoiced zeros, raw energy."""
    g = _gen(seed)
    out = []
    for b in range(batch):
        tp = int(_randint(g, 30, 60, 1))
        base = _randint(g, 3, 18, tp).float() * (0.5 + torch.rand(tp, generator=g))
        r = torch.rand(tp, generator=g)
        base = torch.where(r < 0.02, torch.zeros(()), base)
        base = torch.where((r >= 0.02) & (r < 0.025), -base, base)
        base = torch.where(r > 0.995, 150 + 71 * torch.rand(tp, generator=g), base)
        if b % 2 == 0:
            dur = torch.trunc(base).long()              # integer dtype
        else:
            dur = base.float()                          # fractional: exercises int() truncation
        if int(torch.trunc(dur.double()).clamp_min(0).sum()) == 0:
            dur[0] = 5
        f0 = 80 + 670 * torch.rand(tp, generator=g)
        f0 = torch.where(torch.rand(tp, generator=g) < 0.05, torch.zeros(()), f0)
        en = 150 * torch.rand(tp, generator=g)
        out.append(dict(ids=_randint(g, ZH_LO, ZH_HI, tp), sid=int(_randint(g, 0, N_SPK_USED - 1, 1)),
                        duration=dur, f0=f0, energy=en))
    return out


def frame_counts(utts: List[Dict]) -> List[int]:
    return [int(torch.trunc(u["duration"].double()).clamp_min(0).sum()) for u in utts]


def audio_seconds(utts: List[Dict]) -> float:
    """Valid audio only - padding never counts."""
    return sum(frame_counts(utts)) * HOP / SR


def draw_noise(frames: List[int], seed: int, channels: int = 192) -> List[torch.Tensor]:
    """The eps of models.py:718, one [192, Tf] tensor per utterance."""
    g = _gen(seed)
    return [torch.randn(channels, tf, generator=g) for tf in frames]
