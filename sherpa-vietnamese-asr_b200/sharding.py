"""Multi-GPU: utterance sharding, one process per GPU, no collective on the data path.

The reference's only parallelism is an even/odd chunk split over two worker threads with results stitched on the
host (/root/reference core/asr_engine.py:2384-2397, :2488-2494). Here segments are partitioned across ranks by
duration (longest-processing-time first, so every rank gets the same amount of audio), each rank decodes its
shard on its own GPU, and only the transcripts (a few bytes per token) are gathered on rank 0 through
`torch.distributed.gather_object` — host-side plumbing, not a data-path collective.
"""
from __future__ import annotations

from typing import List, Sequence


def partition_by_duration(n_samples: Sequence[int], world_size: int) -> List[List[int]]:
    """LPT assignment: indices per rank, balanced by total samples; deterministic."""
    order = sorted(range(len(n_samples)), key=lambda i: (-int(n_samples[i]), i))
    loads = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(n_samples[i])
    return shards


def batches_by_length(indices: Sequence[int], n_samples: Sequence[int], max_batch_seconds: float = 3000.0,
                      max_batch_size: int = 512) -> List[List[int]]:
    """Length-sorted batches bounded by total audio, so one decode_streams call stays inside its workspace."""
    order = sorted(indices, key=lambda i: (-int(n_samples[i]), i))
    out, cur, tot = [], [], 0.0
    for i in order:
        d = n_samples[i] / 16000.0
        if cur and (tot + d > max_batch_seconds or len(cur) >= max_batch_size):
            out.append(cur)
            cur, tot = [], 0.0
        cur.append(i)
        tot += d
    if cur:
        out.append(cur)
    return out


def transcribe_sharded(decode_fn, audios, rank: int, world_size: int, gather=None):
    """Each rank runs `decode_fn(list_of_audio) -> list_of_results` on its shard; rank 0 returns results in the
    original order (other ranks return None). `gather(obj)` defaults to torch.distributed.gather_object."""
    n_samples = [len(a) for a in audios]
    shards = partition_by_duration(n_samples, world_size)
    mine = shards[rank]
    local = {}
    for batch in batches_by_length(mine, n_samples):
        for i, r in zip(batch, decode_fn([audios[i] for i in batch])):
            local[i] = r
    if world_size == 1:
        return [local[i] for i in range(len(audios))]
    if gather is None:
        import torch.distributed as dist

        def gather(obj):
            out = [None] * world_size if rank == 0 else None
            dist.gather_object(obj, out, dst=0)
            return out
    parts = gather(local)
    if rank != 0:
        return None
    merged = {}
    for p in parts:
        merged.update(p)
    assert len(merged) == len(audios)
    return [merged[i] for i in range(len(audios))]
