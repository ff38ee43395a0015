"""Utterance-list front end: shard across ranks, bucket by length, run `SynthesizerTrn.infer` per bucket.

This is the multi-utterance entry for configs C3 (mixed 1-15 s, 512 utterances, 1-8 GPUs) and C5; every reference
call site synthesises one utterance at a time (inference.py:40-44), so there is no reference counterpart beyond the
per-utterance result, which is what each entry of the returned list equals.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .sharding import bucket_batches, frames_from_durations, plan_shards


def _pad(seqs: Sequence[torch.Tensor], dtype) -> torch.Tensor:
    n = max(int(s.numel()) for s in seqs)
    out = torch.zeros(len(seqs), n, dtype=dtype)
    for i, s in enumerate(seqs):
        out[i, : s.numel()] = s.to(dtype)
    return out


@torch.no_grad()
def synthesize(net, utts: Sequence[Dict], noise_scale: float = 0.667, rank: int = 0, world_size: int = 1,
               max_frames_per_batch: int = 65536, noises: Optional[Sequence[torch.Tensor]] = None,
               keep_on_device: bool = False) -> Dict[int, torch.Tensor]:
    """utts: dicts with `ids` [Tp], `sid`, `duration` [Tp] (required here: it fixes the frame counts used for the plan),
    optional `f0` (Hz) and `energy` per phoneme.  Returns {utterance index: waveform [samples]} for THIS rank."""
    frames = frames_from_durations([u["duration"] for u in utts])
    plan = plan_shards(frames, world_size)
    mine = plan.indices(rank)
    out: Dict[int, torch.Tensor] = {}
    # one `infer` call takes pitch / energy controls for all of its utterances or for none: utterances are grouped by
    # which controls they carry BEFORE bucketing, so a supplied control is never dropped (and the audio of an utterance
    # does not depend on what it happens to share a bucket with)
    groups: Dict[tuple, List[int]] = {}
    for i in mine:
        groups.setdefault((utts[i].get("f0") is not None, utts[i].get("energy") is not None), []).append(int(i))
    batches = [b for key in sorted(groups) for b in bucket_batches(groups[key], frames, max_frames_per_batch)]
    for batch in batches:
        sel = [utts[i] for i in batch]
        ids = _pad([u["ids"] for u in sel], torch.long)
        float_dur = any(u["duration"].is_floating_point() for u in sel)
        dur = _pad([u["duration"] for u in sel], torch.float64 if float_dur else torch.long)
        kw = {}
        if sel[0].get("f0") is not None:
            kw["pitch_control"] = _pad([u["f0"] for u in sel], torch.float32)
        if sel[0].get("energy") is not None:
            kw["energy_control"] = _pad([u["energy"] for u in sel], torch.float32)
        if noises is not None:
            kw["noise"] = [noises[i] for i in batch]
        o, x_mask, *_ = net.infer(ids, torch.LongTensor([u["ids"].numel() for u in sel]),
                                  sid=torch.LongTensor([int(u["sid"]) for u in sel]), noise_scale=noise_scale,
                                  duration_control=dur, outputs="audio", **kw)
        for b, i in enumerate(batch):
            w = o[b, 0, : int(frames[i]) * net.hop_length]
            out[int(i)] = w.clone() if keep_on_device else w.cpu()
    return out
