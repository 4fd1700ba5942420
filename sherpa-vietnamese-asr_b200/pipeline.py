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
