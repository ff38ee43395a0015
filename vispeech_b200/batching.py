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
               keep_on_device: bool = False, pcm16: bool = False, host_pool: Optional[Dict] = None,
               stats: Optional[Dict] = None) -> Dict[int, torch.Tensor]:
    """utts: dicts with `ids` [Tp], `sid`, `duration` [Tp] (required here: it fixes the frame counts used for the plan),
    optional `f0` (Hz) and `energy` per phoneme.  Returns {utterance index: waveform [samples]} for THIS rank.

    Results travel to the host batch by batch: each `infer` call's output is copied to pinned memory on a copy stream while
    the next call computes, and the host synchronises ONCE at the end (`keep_on_device=True` skips the copies).
    `pcm16=True` converts to signed 16-bit on the device first (vs_wave_pcm16; what a WAV writer wants, half the D2H bytes).
    `host_pool`: a dict the caller keeps between calls to recycle the pinned buffers (their contents are then only valid
    until the next call).  `stats` (optional dict) receives the plan's imbalance, the number of `infer` calls and the
    frames / bytes this rank handled."""
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
    dev = net.device
    copy_stream = None if keep_on_device else _copy_stream(net)
    pending = []                                   # (batch, host buffer) in flight
    d2h = 0
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
        if pcm16:
            from .postprocess import to_pcm16
            o = to_pcm16(o, [int(frames[i]) * net.hop_length for i in batch], net.sampling_rate, net.sampling_rate)
        else:
            o = o[:, 0]
        if keep_on_device:
            for b, i in enumerate(batch):
                out[int(i)] = o[b, : int(frames[i]) * net.hop_length].clone()
            continue
        host = _pinned(host_pool, tuple(o.shape), o.dtype, len(pending))
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(dev))
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            host.copy_(o, non_blocking=True)
            o.record_stream(copy_stream)
        d2h += o.numel() * o.element_size()
        pending.append((batch, host))
    if not keep_on_device:
        copy_stream.synchronize()
        for batch, host in pending:
            for b, i in enumerate(batch):
                out[int(i)] = host[b, : int(frames[i]) * net.hop_length]
    if stats is not None:
        stats.update(imbalance=plan.imbalance, infer_calls=len(batches), utterances=int(len(mine)),
                     frames=int(frames[mine].sum()) if len(mine) else 0, d2h_bytes=int(d2h),
                     planned_load_share=float(plan.load[rank] / max(plan.load.sum(), 1e-30)))
    return out


def _copy_stream(net) -> torch.cuda.Stream:
    if getattr(net, "_copy_stream", None) is None:
        net._copy_stream = torch.cuda.Stream(device=net.device)
    return net._copy_stream


def _pinned(pool: Optional[Dict], shape: tuple, dtype, slot: int) -> torch.Tensor:
    """A pinned host buffer of at least `shape`; recycled through `pool` (keyed by call position) when the caller gave one."""
    n = 1
    for d in shape:
        n *= int(d)
    if pool is not None:
        buf = pool.get((slot, dtype))
        if buf is None or buf.numel() < n:
            buf = torch.empty(n, dtype=dtype, pin_memory=True)
            pool[(slot, dtype)] = buf
        return buf[:n].view(shape)
    return torch.empty(shape, dtype=dtype, pin_memory=True)
