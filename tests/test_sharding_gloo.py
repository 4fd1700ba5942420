"""N>1 path on CPU: world_size-2 gloo processes shard the utterances, decode their shard with a stand-in
decoder, and rank 0 gathers every transcript in the original order."""
import os
import socket

import numpy as np

from sherpa_vietnamese_asr_b200 import sharding


def test_partition_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    n = [int(x) for x in rng.integers(16000, 480000, 257)]
    for world in (1, 2, 4, 8):
        shards = sharding.partition_by_duration(n, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(n)))
        loads = [sum(n[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(n)
    batches = sharding.batches_by_length(range(len(n)), n, max_batch_seconds=600.0)
    assert sorted(i for b in batches for i in b) == list(range(len(n)))
    assert all(sum(n[i] for i in b) / 16000.0 <= 600.0 + 30.0 for b in batches)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    audios = [np.full(int(n), i, dtype=np.float32) for i, n in enumerate(rng.integers(100, 5000, 23))]
    seen = []

    def decode(batch):
        seen.extend(int(a[0]) for a in batch)
        return [(int(a[0]), len(a), rank) for a in batch]
    out = sharding.transcribe_sharded(decode, audios, rank, world)
    # the corpus driver over the same process group: recordings dealt to ranks, transcripts gathered on rank 0
    from oracle import chunk_cases as cc
    from sherpa_vietnamese_asr_b200 import pipeline
    recs = [cc.silence_audio(60 + i, sec) for i, sec in enumerate([40.0, 65.0, 8.0, 33.0])]

    def fake_decode(rec, chunks, offsets):
        return [[{"text": f"r{rank}", "start": o, "end": o + 0.2, "local_start": 0.0, "local_end": 0.2, "prob": 0.9}] for o in offsets]
    dist.barrier()
    stats = {}
    corpus = pipeline.transcribe_corpus(None, recs, rank=rank, world_size=world, decode_chunks=fake_decode, stats=stats)
    assert stats["recordings"] + 0 >= 0 and stats["idle_s"] >= 0.0
    q.put((rank, out, seen, None if corpus is None else [(c["text"], len(c["chunk_plan"])) for c in corpus]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, corpora = {}, {}
    for _ in range(2):
        rank, out, seen, corpus = q.get(timeout=120)
        got[rank] = (out, seen)
        corpora[rank] = corpus
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out0, seen0 = got[0]
    out1, seen1 = got[1]
    assert out1 is None
    assert [o[0] for o in out0] == list(range(23))
    assert set(seen0).isdisjoint(seen1) and len(seen0) + len(seen1) == 23
    assert {o[2] for o in out0} == {0, 1}
    # transcribe_corpus: rank 1 returns nothing, rank 0 has every recording in input order, decoded on both ranks
    assert corpora[1] is None and len(corpora[0]) == 4
    assert [n for _, n in corpora[0]] == [2, 3, 1, 2]
    # the ranks pull from one shared queue (an atomic counter in the job's store): who decodes what is a race by design,
    # every recording is decoded exactly once by one of them
    assert {t.split()[0] for t, _ in corpora[0]} <= {"R0", "R1"}


def test_store_work_queue_hands_every_item_out_once():
    """StoreWorkQueue over a stand-in store: two consumers interleaving their pulls get disjoint items, longest first, and
    None once the queue is drained (also for late askers)."""
    from sherpa_vietnamese_asr_b200 import pipeline

    class Store:
        def __init__(self):
            self.v = {}

        def add(self, key, n):
            self.v[key] = self.v.get(key, 0) + n
            return self.v[key]

    store = Store()
    order = [3, 0, 2, 1, 4]
    a, b = pipeline.StoreWorkQueue(store, order), pipeline.StoreWorkQueue(store, order)
    got = [a.next(), b.next(), b.next(), a.next(), b.next(), a.next(), b.next()]
    assert got == [3, 0, 2, 1, 4, None, None]
    lq = pipeline.LocalWorkQueue([5, 6])
    assert [lq.next(), lq.next(), lq.next()] == [5, 6, None]
