"""Multi-process (gloo, world_size 2, CPU) checks of the utterance-sharding host logic used for N>1 GPUs."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vispeech_b200.sharding import bucket_batches, frames_from_durations, plan_shards


def test_plan_is_partition_and_balanced():
    rng = np.random.default_rng(0)
    frames = rng.integers(86, 1293, size=512)                 # C3: 1-15 s
    for ws in (1, 2, 4, 8):
        plan = plan_shards(frames, ws)
        assert sorted(np.concatenate([plan.indices(r) for r in range(ws)]).tolist()) == list(range(512))
        assert plan.imbalance < 1.02                          # LPT on 512 items: within 2 % of perfect
    assert plan_shards([100], 4).assignment.tolist() == [0]


def test_buckets_respect_boundaries_and_cap():
    rng = np.random.default_rng(1)
    frames = rng.integers(50, 3000, size=200)
    batches = bucket_batches(range(200), frames, max_frames_per_batch=8000)
    seen = sorted(i for b in batches for i in b)
    assert seen == list(range(200))
    bounds = np.asarray([0, 128, 256, 384, 512, 768, 1024, 1536, 2048, 4096, 1 << 30])
    for b in batches:
        ks = {int(np.searchsorted(bounds, frames[i], side="right")) for i in b}
        assert len(ks) == 1
        assert sum(int(frames[i]) for i in b) <= 8000 or len(b) == 1


def test_frames_from_durations_rule():
    d = [torch.tensor([2.9, 0.0, -1.5, 3.2]), torch.tensor([1, 0, 7])]
    assert frames_from_durations(d).tolist() == [5, 8]


def _worker(rank, world, port, frames, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = plan_shards(frames, world)
    mine = plan.indices(rank)
    # every rank derives the same plan without communicating: check by all-gathering a digest and the shard sizes
    digest = torch.tensor([int(np.dot(plan.assignment.astype(np.int64), np.arange(1, frames.size + 1)))])
    got = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(got, digest)
    assert all(int(g) == int(digest) for g in got)
    n = torch.tensor([mine.size])
    dist.all_reduce(n)
    assert int(n) == frames.size
    work = torch.tensor([float(frames[mine].sum())])
    tot = work.clone()
    dist.all_reduce(tot)
    mx = work.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), mine)
    assert float(mx) <= 0.52 * float(tot)
    dist.destroy_process_group()


def test_two_ranks_agree_and_cover(tmp_path):
    frames = np.random.default_rng(2).integers(86, 1293, size=64)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, frames, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert sorted(np.concatenate([a, b]).tolist()) == list(range(64)) and not set(a) & set(b)
