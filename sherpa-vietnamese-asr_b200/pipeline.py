"""The transcription phase of the reference's file pipeline as one call over the GPU engine:

    VAD segments -> [RMS normalise] -> peak limit -> 5 s gap merge -> speech-only concat -> silence-aligned 30 s chunks ->
    ONE ragged GPU batch (both models in ROVER mode) -> word times back to the recording -> overlap stitch ->
    suspect flags -> filler removal

following /root/reference core/asr_engine.py `_run_pipeline` :2076-2161 (VAD, preprocessing, chunk plan), :2326-2496 (decode,
ROVER, stitch) and :2556-2580 (suspect detection, fillers, text). Diarization, punctuation and the UI events of that method are
outside the path. Every step is one of the parity-tested functions of vad.py / staging.py / chunking.py / postprocess.py.
"""
from __future__ import annotations

import time
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import chunking, postprocess, staging, vad


def transcribe_recording(recognizer, audio: np.ndarray, vad_prob_fn: Optional[vad.ProbFn] = None,
                         vad_segments: Optional[Sequence[Tuple[int, int]]] = None, rover_recognizer=None,
                         hotword_phrases: Sequence[str] = (), skip_preprocessing: bool = False, rms_normalize: bool = False,
                         decode_chunks=None) -> Dict[str, object]:
    """`vad_prob_fn` (windows[n, 576] -> probs[n]) or precomputed `vad_segments`; with neither the whole recording is speech
    (the reference's bypass_vad path, :2085-2086); a VAD failure takes the same path, as in the reference (:2171-2204).
    Returns {"words", "text", "vad_segments", "chunk_plan", "chunk_results", "timing", "vad_error"}; `timing` holds the wall-clock
    spans the reference keeps in timing_details (:1969-1977)."""
    audio = np.ascontiguousarray(audio, dtype=np.float32)
    timing: Dict[str, float] = {}
    t0 = time.perf_counter()
    probs = None
    vad_error = None
    try:
        if vad_segments is None and vad_prob_fn is not None:
            vad_segments, probs = vad.get_vad_segments(audio, vad_prob_fn)
        if vad_segments is not None:
            vad_segments = list(vad_segments)
            if not skip_preprocessing:
                try:                                  # a preprocessing error is skipped, not fatal (:2099-2113)
                    audio = staging.preprocess_audio(audio, vad_segments, enable_rms_normalize=rms_normalize)
                except Exception as e:  # noqa: BLE001
                    vad_error = f"preprocess: {e!r}"
            vad_segments = vad.merge_close_segments(vad_segments, vad.MAX_VAD_GAP, inclusive=True)
    except Exception as e:  # noqa: BLE001   any VAD failure falls back to silence chunking of the whole recording (:2171-2204)
        vad_error, vad_segments, probs = f"vad: {e!r}", None, None
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


def transcribe_corpus(recognizer, recordings: Sequence[np.ndarray], vad_segments: Optional[Sequence[Optional[Sequence[Tuple[int, int]]]]] = None,
                      rover_recognizer=None, hotword_phrases: Sequence[str] = (), rank: int = 0, world_size: int = 1,
                      max_batch_seconds: float = 3000.0, decode_chunks=None, gather=None):
    """Many recordings at once (BASELINE config C5: a 10 h corpus of 15-minute files). The frame loop of the search costs
    the same ~27 ms whether a batch holds 30 chunks or 256, so chunks are pooled ACROSS recordings: every recording is
    planned, the chunks of this rank's recordings are sorted by length into batches of up to `max_batch_seconds` of audio,
    each batch is one ragged GPU pass, and the words go back to their recording for time mapping, stitching and the
    post-ASR steps. Recordings are dealt to ranks by duration (sharding.partition_by_duration); only the finished
    transcripts travel to rank 0 (`gather`, default torch.distributed.gather_object) - no data-path collective.
    Returns, on rank 0, one result dict per recording in input order (None on other ranks)."""
    from . import sharding
    if decode_chunks is None:
        from .asr_engine import decode_chunks
        device_id = int(recognizer.engine.device_id)
    else:
        device_id = None
    segs = list(vad_segments) if vad_segments is not None else [None] * len(recordings)
    mine = sharding.partition_by_duration([len(a) for a in recordings], world_size)[rank]
    plans = {i: chunking.plan_recording(recordings[i], segs[i] or (), device_id) for i in mine}
    pool = [(i, k) for i in mine for k in range(len(plans[i].plan))]                 # (recording, chunk)
    lengths = [plans[i].plan[k][1] - plans[i].plan[k][0] for i, k in pool]
    chunk_audio = {i: plans[i].chunks for i in mine}
    chunk_offsets = {i: plans[i].offsets for i in mine}
    decoded = {"a": {}, "b": {}}
    for batch in sharding.batches_by_length(range(len(pool)), lengths, max_batch_seconds):
        audio_b = [chunk_audio[pool[j][0]][pool[j][1]] for j in batch]
        offs_b = [chunk_offsets[pool[j][0]][pool[j][1]] for j in batch]
        for j, words in zip(batch, decode_chunks(recognizer, audio_b, offs_b)):
            decoded["a"][pool[j]] = words
        if rover_recognizer is not None:
            for j, words in zip(batch, decode_chunks(rover_recognizer, audio_b, offs_b)):
                decoded["b"][pool[j]] = words
    local = {}
    for i in mine:
        n = len(plans[i].plan)
        per_chunk = [decoded["a"][(i, k)] for k in range(n)]
        per_rover = [decoded["b"][(i, k)] for k in range(n)] if rover_recognizer is not None else None
        res = chunking.finish_recording(plans[i], per_chunk, per_rover, hotword_phrases)
        words, text = postprocess.finish_transcript(res["words"], recordings[i], is_rover=rover_recognizer is not None)
        local[i] = {"words": words, "text": text, "chunk_plan": res["chunk_plan"]}
    if world_size == 1:
        return [local[i] for i in range(len(recordings))]
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
    return [merged[i] for i in range(len(recordings))]
