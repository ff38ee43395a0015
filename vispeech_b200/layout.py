"""Ragged-row layouts (host side of `VsRows`, include/vispeech_b200.h).

A batch is one long sequence of rows; utterance b owns [start[b], start[b]+len[b]) and at least `gap`
zero rows separate utterances, so that every conv sees the zero padding a batch-1 reference call sees.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch

from ._lib import VsRows

PHONEME_GAP = 4      # >= 1 (k=3 convs at phoneme rate)
FRAME_GAP = 4        # >= 3 frames (conv_pre k=7); x8 at decoder stage 0 = 32 rows >= 25 (k=11, dilation 5)
ROW_ALIGN = 16       # frame rows x 8 (first upsample) = multiple of the 128-row MMA tile


@dataclass
class RaggedRows:
    lengths: np.ndarray          # [B] int32
    starts: np.ndarray           # [B] int32
    n_rows: int
    row_utt_host: np.ndarray     # [n_rows] int32
    row_utt: torch.Tensor        # device int32
    utt_start: torch.Tensor
    utt_len: torch.Tensor
    sid: torch.Tensor
    struct: VsRows

    @property
    def n_utt(self) -> int:
        return int(self.lengths.shape[0])

    @property
    def max_len(self) -> int:
        return int(self.lengths.max()) if self.lengths.size else 0

    def scatter(self, values: Sequence, dtype, fill=0, width: int = 0) -> np.ndarray:
        """Per-utterance arrays ([len_b] or [len_b, width]) -> ragged host array ([n_rows] or [n_rows, width])."""
        shape = (self.n_rows, width) if width else (self.n_rows,)
        out = np.full(shape, fill, dtype=dtype)
        for b in range(self.n_utt):
            n = int(self.lengths[b])
            if n:
                out[self.starts[b]:self.starts[b] + n] = np.asarray(values[b])[:n]
        return out


def plan_starts(lengths: Sequence[int], gap: int, align: int = ROW_ALIGN):
    lengths = np.asarray(lengths, dtype=np.int64)
    starts = np.zeros_like(lengths)
    pos = 0
    for b, n in enumerate(lengths):
        starts[b] = pos
        pos += int(n) + gap
    n_rows = int(max(align, -(-pos // align) * align))
    return starts.astype(np.int32), n_rows


def make_rows(lengths: Sequence[int], sids: Sequence[int], gap: int, device, n_rows: int = 0,
              out: "RaggedRows" = None) -> RaggedRows:
    """`n_rows` > 0: lay the batch out in a BUCKET of that many rows (>= what the batch needs): the tail is gap rows and
    `max_len` (a launch-geometry bound only; kernels read the true lengths from device memory) becomes the bucket size, so
    the layout's shape no longer depends on the utterance - what a captured CUDA graph needs.  `out`: reuse that layout's
    device arrays (same bucket, same utterance count) instead of allocating new ones."""
    lengths = np.asarray(lengths, dtype=np.int32)
    if lengths.ndim != 1 or lengths.size == 0:
        raise ValueError("need at least one utterance")
    if (lengths < 0).any():
        raise ValueError("negative length")
    starts, need = plan_starts(lengths, gap)
    bucket = n_rows > 0
    if bucket and n_rows < need:
        raise ValueError("bucket of %d rows is too small for %d" % (n_rows, need))
    n_rows = n_rows if bucket else need
    if n_rows >= 2 ** 31 // 512:
        raise ValueError("batch too long for 32-bit row indices at sample rate")
    row_utt = np.full(n_rows, -1, dtype=np.int32)
    for b in range(lengths.size):
        row_utt[starts[b]:starts[b] + lengths[b]] = b
    meta = np.concatenate([row_utt, starts, lengths, np.asarray(sids, dtype=np.int32)])
    # pinned + asynchronous: a blocking copy would make the host wait for everything already queued on this stream (in the
    # engine's throughput mode that is the previous call's latent stage, which shares the SMs with a decoder - the host then
    # falls behind and the GPU idles; seen as sporadic 20-40 % slower end-to-end runs)
    meta_host = torch.from_numpy(meta).pin_memory() if torch.cuda.is_available() else torch.from_numpy(meta)
    if out is not None:
        if out.meta_dev.numel() != meta_host.numel() or out.n_rows != n_rows:
            raise ValueError("layout does not match the bucket it is written into")
        meta_dev = out.meta_dev
        meta_dev.copy_(meta_host, non_blocking=True)
    else:
        meta_dev = meta_host.to(device, non_blocking=True)
    B = lengths.size
    d_row_utt = meta_dev[:n_rows]
    d_start = meta_dev[n_rows:n_rows + B]
    d_len = meta_dev[n_rows + B:n_rows + 2 * B]
    d_sid = meta_dev[n_rows + 2 * B:]
    st = VsRows(n_utt=B, n_rows=n_rows, max_len=n_rows if bucket else max(1, int(lengths.max())), reserved=0,
                row_utt=d_row_utt.data_ptr(), utt_start=d_start.data_ptr(), utt_len=d_len.data_ptr(),
                sid=d_sid.data_ptr())
    rows = RaggedRows(lengths, starts, n_rows, row_utt, d_row_utt, d_start, d_len, d_sid, st)
    rows.meta_host = meta_host          # keeps the staging buffer referenced with the layout
    rows.meta_dev = meta_dev
    return rows
