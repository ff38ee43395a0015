"""Utterance sharding for multi-GPU inference: independent utterances, no exchange step.

One process per GPU.  Every rank computes the SAME deterministic plan from the batch's frame counts and then
synthesises only its own utterances; nothing crosses NVLink on the data path.  The plan
  1. estimates each utterance's cost (decoder work is linear in frames, attention quadratic),
  2. assigns utterances to ranks by longest-processing-time-first (greedy on the least loaded rank),
  3. within a rank groups utterances into length buckets (boundaries in frames, cf. the reference's
     DistributedBucketSampler, data_utils.py:219-318 / train.py:71) so that one `infer` call handles similar lengths
     and its activation workspace stays bounded.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

DEFAULT_BOUNDARIES = (0, 128, 256, 384, 512, 768, 1024, 1536, 2048, 4096, 1 << 30)   # frames
DECODER_FLOP_PER_FRAME = 815_300_608 + 14_155_776 + 8_400_000                        # decoder + flow + frame prior
ATTN_FLOP_PER_FRAME2 = 4 * 2 * (2 * 96 + 2 * 96)                                    # 4 layers x 2 heads x (QK^T + PV)


def utterance_cost(frames: np.ndarray) -> np.ndarray:
    f = frames.astype(np.float64)
    return DECODER_FLOP_PER_FRAME * f + ATTN_FLOP_PER_FRAME2 * f * f


@dataclass
class ShardPlan:
    world_size: int
    assignment: np.ndarray            # [n_utt] -> rank
    load: np.ndarray                  # [world_size] estimated cost

    def indices(self, rank: int) -> np.ndarray:
        return np.nonzero(self.assignment == rank)[0]

    @property
    def imbalance(self) -> float:
        """max load / mean load (1.0 = perfect)."""
        return float(self.load.max() / max(self.load.mean(), 1e-30))


def plan_shards(frames: Sequence[int], world_size: int) -> ShardPlan:
    frames = np.asarray(frames, dtype=np.int64)
    cost = utterance_cost(frames)
    order = np.lexsort((np.arange(frames.size), -cost))      # cost descending, index ascending: deterministic
    load = np.zeros(world_size, dtype=np.float64)
    assignment = np.zeros(frames.size, dtype=np.int32)
    for i in order:
        r = int(np.argmin(load))                              # first minimum: deterministic tie-break
        assignment[i] = r
        load[r] += cost[i]
    return ShardPlan(world_size, assignment, load)


def bucket_batches(indices: Sequence[int], frames: Sequence[int], max_frames_per_batch: int = 65536,
                   boundaries: Sequence[int] = DEFAULT_BOUNDARIES) -> List[List[int]]:
    """Split one rank's utterances into `infer` calls: same length bucket, at most `max_frames_per_batch` frames each."""
    frames = np.asarray(frames, dtype=np.int64)
    buckets: dict = {}
    for i in indices:
        b = int(np.searchsorted(np.asarray(boundaries), frames[i], side="right")) - 1
        buckets.setdefault(b, []).append(int(i))
    batches: List[List[int]] = []
    for b in sorted(buckets):
        cur, tot = [], 0
        for i in sorted(buckets[b], key=lambda j: (-int(frames[j]), j)):
            if cur and tot + int(frames[i]) > max_frames_per_batch:
                batches.append(cur)
                cur, tot = [], 0
            cur.append(i)
            tot += int(frames[i])
        if cur:
            batches.append(cur)
    return batches


def frames_from_durations(durations: Sequence) -> np.ndarray:
    """n = sum(max(int(d), 0)) per utterance - the length regulator's rule (models.py:421-423)."""
    out = []
    for d in durations:
        a = np.asarray(d, dtype=np.float64)
        out.append(int(np.clip(np.trunc(a), 0, None).sum()))
    return np.asarray(out, dtype=np.int64)
