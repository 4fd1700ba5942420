"""The transcription phase of the reference's file pipeline as one call over the GPU engine:

    VAD segments -> [RMS normalise] -> peak limit -> 5 s gap merge -> speech-only concat -> silence-aligned 30 s chunks ->
    ONE ragged GPU batch (both models in ROVER mode) -> word times back to the recording -> overlap stitch ->
    suspect flags -> filler removal

following /root/reference core/asr_engine.py `_run_pipeline` :2076-2161 (VAD, preprocessing, chunk plan), :2326-2496 (decode,
ROVER, stitch) and :2556-2580 (suspect detection, fillers, text). Diarization, punctuation and the UI events of that method are
outside the path. Every step is one of the parity-tested functions of vad.py / staging.py / chunking.py / postprocess.py.
"""
from __future__ import annotations

import queue as _queue
import threading
import time
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import chunking, postprocess, staging, vad




def prepare_recording(audio: np.ndarray, vad_prob_fn: Optional[vad.ProbFn] = None,
                      vad_segments: Optional[Sequence[Tuple[int, int]]] = None, skip_preprocessing: bool = False,
                      rms_normalize: bool = False, device_id: Optional[int] = None):
    """The steps `_run_pipeline` takes before the chunk plan (core/asr_engine.py:2076-2128): VAD segments (computed from
    `vad_prob_fn` or given), preprocess_audio on the recording (peak limit, optional per-segment RMS normalisation; an error
    there is skipped, :2099-2113), the 5 s gap merge. With neither VAD input the whole recording is speech (:2085-2086); any
    VAD failure takes the same path (:2171-2204). One helper for every entry point, so a recording gets the same speech
    concatenation and chunk plan however it is transcribed. `device_id` not None: the preprocessing passes run on that GPU
    (csrc/staging.cu). Returns (audio, vad_segments or None, vad_probs or None, vad_error)."""
    audio = np.ascontiguousarray(audio, dtype=np.float32)
    probs, vad_error = None, None
    try:
        if vad_segments is None and vad_prob_fn is not None:
            vad_segments, probs = vad.get_vad_segments(audio, vad_prob_fn)
        if vad_segments is not None:
            vad_segments = list(vad_segments)
            if not skip_preprocessing:
                try:
                    if device_id is None:
                        audio = staging.preprocess_audio(audio, vad_segments, enable_rms_normalize=rms_normalize)
                    else:
                        audio = staging.preprocess_audio_gpu(audio, vad_segments, enable_rms_normalize=rms_normalize, device_id=device_id)
                except Exception as e:  # noqa: BLE001
                    vad_error = f"preprocess: {e!r}"
            vad_segments = vad.merge_close_segments(vad_segments, vad.MAX_VAD_GAP, inclusive=True)
    except Exception as e:  # noqa: BLE001
        vad_error, vad_segments, probs = f"vad: {e!r}", None, None
    return audio, vad_segments, probs, vad_error


def transcribe_recording(recognizer, audio: np.ndarray, vad_prob_fn: Optional[vad.ProbFn] = None,
                         vad_segments: Optional[Sequence[Tuple[int, int]]] = None, rover_recognizer=None,
                         hotword_phrases: Sequence[str] = (), skip_preprocessing: bool = False, rms_normalize: bool = False,
                         decode_chunks=None) -> Dict[str, object]:
    """`vad_prob_fn` (windows[n, 576] -> probs[n]) or precomputed `vad_segments`; with neither the whole recording is speech.
    Returns {"words", "text", "vad_segments", "chunk_plan", "chunk_results", "timing", "vad_error"}; `timing` holds the wall-clock
    spans the reference keeps in timing_details (:1969-1977)."""
    timing: Dict[str, float] = {}
    t0 = time.perf_counter()
    device_id = None if decode_chunks is not None else int(recognizer.engine.device_id)   # real engine: staging on its GPU too
    audio, vad_segments, probs, vad_error = prepare_recording(audio, vad_prob_fn, vad_segments, skip_preprocessing, rms_normalize, device_id)
    timing["vad_preprocess"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    res = chunking.transcribe_long(recognizer, audio, vad_segments or (), decode_chunks=decode_chunks,
                                   rover_recognizer=rover_recognizer, hotword_phrases=hotword_phrases)
    timing["transcription"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    words, text = postprocess.finish_transcript(res["words"], audio, is_rover=rover_recognizer is not None, vad_probs=probs)
    timing["postprocess"] = time.perf_counter() - t0
    return {"words": words, "text": text, "vad_segments": vad_segments, "chunk_plan": res["chunk_plan"],
            "chunk_results": res["chunk_results"], "timing": timing, "vad_error": vad_error}


# ----------------------------------------------------------------------------- corpus: a pull queue over recordings
class LocalWorkQueue:
    """Work items 0..n-1 handed out once each, in the given order, to whoever asks (threads of one process)."""

    def __init__(self, order: Sequence[int]):
        self._order, self._at, self._mu = list(order), 0, threading.Lock()

    def next(self) -> Optional[int]:
        with self._mu:
            if self._at >= len(self._order):
                return None
            self._at += 1
            return self._order[self._at - 1]


class StoreWorkQueue:
    """The same queue shared by all ranks of a torch.distributed job: an atomic counter in the job's key-value store
    (`store.add`), so a GPU that finishes early pulls the next recording instead of waiting for a static share. Host
    plumbing only - nothing on the data path is a collective."""

    def __init__(self, store, order: Sequence[int], key: str = "b200asr/next_recording"):
        self._store, self._order, self._key = store, list(order), key

    def next(self) -> Optional[int]:
        k = int(self._store.add(self._key, 1)) - 1
        return self._order[k] if k < len(self._order) else None


def transcribe_corpus(recognizer, recordings: Sequence[np.ndarray], vad_segments: Optional[Sequence[Optional[Sequence[Tuple[int, int]]]]] = None,
                      rover_recognizer=None, hotword_phrases: Sequence[str] = (), rank: int = 0, world_size: int = 1,
                      max_batch_seconds: float = 3000.0, decode_chunks=None, gather=None, vad_prob_fn: Optional[vad.ProbFn] = None,
                      skip_preprocessing: bool = False, rms_normalize: bool = False, work_queue=None, prefetch: int = 8,
                      stats: Optional[dict] = None, vad_prob_fns: Optional[Sequence[Optional[vad.ProbFn]]] = None, planners: int = 2):
    """Many recordings at once (BASELINE config C5: a 10 h corpus of 15-minute files), every rank pulling from one queue.

    * Queue: recordings, longest first. `work_queue.next()` hands each one out exactly once - LocalWorkQueue inside a
      process, StoreWorkQueue across the ranks of a torchrun job (default when world_size > 1 and torch.distributed is
      initialised; without it the recordings are dealt statically by duration, sharding.partition_by_duration).
    * `planners` producer threads (host): pull -> prepare_recording (VAD, preprocessing, 5 s merge) -> speech concatenation and
      chunk plan (energy scan on the GPU), together up to `prefetch` recordings ahead of the decoder.
    * Consumer (this thread): takes a round of up to `prefetch` planned recordings (while it decodes, the producer plans the
      next round), pools their chunks ACROSS recordings into
      length-sorted batches of up to `max_batch_seconds` (the search's frame loop costs the same whether a batch holds 30
      chunks or 256), one ragged GPU pass per batch (two in ROVER mode, sharing the fbank), then word times back to each
      recording, overlap stitch, suspect flags, filler removal - the same functions as transcribe_recording, so
      transcribe_corpus([a]) equals transcribe_recording(a).
    Only finished transcripts travel to rank 0 (`gather`, default torch.distributed.gather_object). `stats`, if given,
    receives this rank's counters: recordings, batches, chunks, idle_s (decoder waiting for the producer), decode_s.
    Returns, on rank 0, one result dict per recording in input order (None on other ranks)."""
    from . import sharding
    if decode_chunks is None:
        from .asr_engine import decode_chunks as _dc
        device_id = int(recognizer.engine.device_id)
        shared_fbank = rover_recognizer is not None
    else:
        _dc, device_id, shared_fbank = decode_chunks, None, False
    n = len(recordings)
    segs = list(vad_segments) if vad_segments is not None else [None] * n
    order = sorted(range(n), key=lambda i: (-len(recordings[i]), i))
    if work_queue is None:
        store = None
        if world_size > 1:
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized() and dist.get_world_size() == world_size:
                    store = dist.distributed_c10d._get_default_store()
            except Exception:  # noqa: BLE001
                store = None
        if store is not None:
            work_queue = StoreWorkQueue(store, order)
        elif world_size > 1:
            mine = set(sharding.partition_by_duration([len(a) for a in recordings], world_size)[rank])
            work_queue = LocalWorkQueue([i for i in order if i in mine])
        else:
            work_queue = LocalWorkQueue(order)

    # a rank must not hoard the shared queue: it runs at most half its fair share ahead of its decoder
    prefetch = max(1, prefetch if world_size <= 1 else min(prefetch, -(-n // (2 * world_size))))
    ready: "_queue.Queue" = _queue.Queue(maxsize=prefetch)
    done = object()

    stop = threading.Event()          # set when the consumer leaves early (an error): planners stop pulling and never block

    def hand_over(item):
        while not stop.is_set():
            try:
                ready.put(item, timeout=0.2)
                return
            except _queue.Full:
                continue

    def producer():
        try:
            while not stop.is_set():
                i = work_queue.next()
                if i is None:
                    break
                pf = vad_prob_fns[i] if vad_prob_fns is not None else vad_prob_fn      # per-recording VAD, or one for all
                audio, vs, probs, err = prepare_recording(recordings[i], pf, segs[i], skip_preprocessing, rms_normalize, device_id)
                hand_over((i, audio, probs, err, vs, chunking.plan_recording(audio, vs or (), device_id)))
        except BaseException as e:  # noqa: BLE001   surfaces in the consumer
            hand_over(e)
        hand_over(done)

    planners = max(1, int(planners))
    ths = [threading.Thread(target=producer, daemon=True) for _ in range(planners)]
    for th in ths:
        th.start()
    local: Dict[int, dict] = {}
    st = {"recordings": 0, "batches": 0, "chunks": 0, "idle_s": 0.0, "decode_s": 0.0}
    is_rover = rover_recognizer is not None

    # Finisher thread: word times back to each recording, overlap stitch, suspect flags, filler removal of a decoded round,
    # while this thread already decodes the next round (the GPU would otherwise idle through the host-only tail of every round).
    to_finish: "_queue.Queue" = _queue.Queue(maxsize=2)
    finish_error: List[BaseException] = []

    def finisher():
        while True:
            job = to_finish.get()
            if job is done:
                return
            if finish_error:
                continue                      # drain: the consumer must never block on a dead finisher
            try:
                items, decoded = job
                for i, audio, probs, err, vs, rp in items:
                    k = len(rp.plan)
                    per_chunk = [decoded["a"][(i, c)] for c in range(k)]
                    per_rover = [decoded["b"][(i, c)] for c in range(k)] if is_rover else None
                    res = chunking.finish_recording(rp, per_chunk, per_rover, hotword_phrases)
                    words, text = postprocess.finish_transcript(res["words"], audio, is_rover=is_rover, vad_probs=probs)
                    local[i] = {"words": words, "text": text, "chunk_plan": res["chunk_plan"], "vad_segments": vs, "vad_error": err}
            except BaseException as e:  # noqa: BLE001   surfaces in the consumer
                finish_error.append(e)

    fin = threading.Thread(target=finisher, daemon=True)
    fin.start()
    finished, live = False, planners
    try:
        while not finished:
            items = []
            while len(items) < max(1, prefetch) and not finished:     # a round = up to `prefetch` planned recordings
                t0 = time.perf_counter()
                it = ready.get()
                st["idle_s"] += time.perf_counter() - t0
                if it is done:
                    live -= 1
                    finished = live == 0
                elif isinstance(it, BaseException):
                    raise it
                else:
                    items.append(it)
            if finish_error:
                raise finish_error[0]
            if not items:
                continue
            plans = {i: rp for i, _, _, _, _, rp in items}
            pool = [(i, k) for i in plans for k in range(len(plans[i].plan))]                 # (recording, chunk)
            lengths = [plans[i].plan[k][1] - plans[i].plan[k][0] for i, k in pool]
            decoded: Dict[str, dict] = {"a": {}, "b": {}}
            t0 = time.perf_counter()
            for batch in sharding.batches_by_length(range(len(pool)), lengths, max_batch_seconds):
                audio_b = [plans[pool[j][0]].chunks[pool[j][1]] for j in batch]
                offs_b = [plans[pool[j][0]].offsets[pool[j][1]] for j in batch]
                feats = recognizer.engine.fbank_batch(audio_b) if shared_fbank else None
                kw = {"precomputed_features": feats} if feats is not None else {}
                for j, words in zip(batch, _dc(recognizer, audio_b, offs_b, **kw)):
                    decoded["a"][pool[j]] = words
                if rover_recognizer is not None:
                    for j, words in zip(batch, _dc(rover_recognizer, audio_b, offs_b, **kw)):
                        decoded["b"][pool[j]] = words
                st["batches"] += 1
                st["chunks"] += len(batch)
            st["decode_s"] += time.perf_counter() - t0
            st["recordings"] += len(items)
            to_finish.put((items, decoded))
    finally:
        stop.set()
        to_finish.put(done)
        fin.join()
    if finish_error:
        raise finish_error[0]
    for th in ths:
        th.join()
    if stats is not None:
        stats.update(st)
    if world_size == 1:
        return [local[i] for i in range(n)]
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
    for part in parts:
        merged.update(part)
    return [merged[i] for i in range(n)]
