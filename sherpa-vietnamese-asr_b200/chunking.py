"""Chunk planner and overlap stitcher: the immediate callers of the recognizer in the reference's pipeline
(SURVEY.md section 8f, rank 1), so a long recording goes through the GPU engine as ONE ragged batch.

Host-side restatement (own code, same behaviour) of /root/reference core/asr_engine.py:
  find_silent_regions      :521-553   10 ms RMS frames under a threshold, runs of at least 0.3 s
  find_best_split_point    :556-573   midpoint of the silent run closest to the target, within +-2 s
  plan_chunks              :2141-2161 30 s logical segments cut at silences, 3 s overlap prepended from the 2nd chunk on
  chunk_long_segment       :583-614   one long VAD segment into ceil(d/30) equal chunks, neighbours overlapping by 3 s
  concat_vad_speech        :617-643   speech-only audio + (concat start, original start, length) map
  map_concat_time_to_original :646-676 word times back to the original recording
  words_match              :52-67     equal | containment (both > 2 chars) | difflib ratio >= 0.8
  find_overlap_alignment   :70-179    sliding-offset fuzzy alignment of the previous tail and the next head; on divergence
                                      the side with the lower mean word probability is dropped
  merge_chunks_with_overlap:182-237   stitches per-chunk word lists with it
`transcribe_long` is the batch replacement of the reference's per-chunk worker loop (:2326-2467): plan, one
`decode_chunks` call over all chunks (ragged GPU batch), stitch.
Parity: tests/test_chunking.py runs these against the reference's own functions (live where /root/reference exists, and
against golden vectors produced by them).
"""
from __future__ import annotations

import bisect
import functools
import math
import re
import unicodedata
from difflib import SequenceMatcher
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

OVERLAP_SEC = 3.0                      # core/asr_engine.py:33
OVERLAP_SAMPLES = int(OVERLAP_SEC * 16000)
MAX_OVERLAP_WORDS = 100                # :35
FUZZY_MATCH_THRESHOLD = 0.8            # :36
MIN_MATCH_RATIO = 0.5                  # :37
SEGMENT_SAMPLES = 30 * 16000           # :2141
MIN_SEGMENT_SAMPLES = 20 * 16000       # :2147 (a split closer than 20 s to the previous one falls back to the hard target);
                                       # two thirds of the segment when a caller asks for another segment length


def _quiet_runs(quiet: np.ndarray, frame: int, n_samples: int, min_silence_duration: float) -> List[Tuple[int, int]]:
    """Run-length encoding of the per-frame quiet flags: runs of at least min_silence_duration, in samples."""
    edges = np.flatnonzero(np.diff(np.concatenate(([0], quiet.astype(np.int8), [0]))))
    starts, ends = edges[0::2], edges[1::2]
    min_frames = int(min_silence_duration / 0.01)
    return [(s * frame, min(e * frame, n_samples)) for s, e in zip(starts.tolist(), ends.tolist()) if e - s >= min_frames]


def find_silent_regions_gpu(audio: np.ndarray, sample_rate: int = 16000, threshold: float = 0.01,
                            min_silence_duration: float = 0.3, device_id: int = 0) -> List[Tuple[int, int]]:
    """`find_silent_regions` with the energy scan on the GPU (csrc/energy.cu through B200AsrSilentFrames); the kernel
    reproduces NumPy's float32 arithmetic, so the regions are the reference's regions. float32 PCM at 16 kHz only."""
    from . import _capi
    if sample_rate != 16000:
        raise ValueError("the GPU energy scan handles 16 kHz audio only")
    x = np.ascontiguousarray(audio, dtype=np.float32)
    n = len(x) // 160
    if n == 0:
        return []
    quiet = np.empty(n, dtype=np.uint8)
    rc = _capi.lib().B200AsrSilentFrames(_capi.fptr(x), len(x), sample_rate, float(np.float32(threshold)),
                                         quiet.ctypes.data_as(_capi.C.POINTER(_capi.C.c_uint8)), device_id)
    if rc != n:
        raise RuntimeError("B200AsrSilentFrames failed: " + _capi.last_error())
    return _quiet_runs(quiet, 160, len(x), min_silence_duration)


def find_silent_regions(audio: np.ndarray, sample_rate: int = 16000, threshold: float = 0.01,
                        min_silence_duration: float = 0.3) -> List[Tuple[int, int]]:
    """Host NumPy version (any sample rate / dtype)."""
    frame = int(sample_rate * 0.01)
    n = len(audio) // frame
    if n == 0:
        return []
    x = np.asarray(audio[: n * frame]).reshape(n, frame)
    quiet = np.sqrt(np.mean(x ** 2, axis=1)) < threshold
    return _quiet_runs(quiet, frame, len(audio), min_silence_duration)


def find_best_split_point(target: int, total: int, silent_regions: Sequence[Tuple[int, int]], search_window: int = 2 * 16000) -> int:
    lo, hi = max(0, target - search_window), min(total, target + search_window)
    best, best_d = target, None
    for s, e in silent_regions:
        if e >= lo and s <= hi:
            mid = (s + e) // 2
            d = abs(mid - target)
            if best_d is None or d < best_d:
                best, best_d = mid, d
    return best


def plan_chunks(total_samples: int, silent_regions: Sequence[Tuple[int, int]], segment_samples: int = SEGMENT_SAMPLES,
                overlap_samples: int = OVERLAP_SAMPLES) -> List[Tuple[int, int, int]]:
    """[(start, end, overlap_at_start)] in samples."""
    min_segment = segment_samples * 2 // 3
    bounds = [0]
    pos = 0
    while pos + segment_samples < total_samples:
        target = pos + segment_samples
        cut = find_best_split_point(target, total_samples, silent_regions)
        if cut <= pos + min_segment:
            cut = target
        bounds.append(cut)
        pos = cut
    bounds.append(total_samples)
    plan = []
    for i in range(len(bounds) - 1):
        a, b = bounds[i], bounds[i + 1]
        if i == 0:
            plan.append((a, b, 0))
        else:
            s = max(0, a - overlap_samples)
            plan.append((s, b, a - s))
    return plan


def chunk_long_segment(seg_start: int, seg_end: int, max_sec: float = 30, overlap_sec: float = 3.0,
                       sample_rate: int = 16000) -> List[Tuple[int, int, int]]:
    duration = (seg_end - seg_start) / sample_rate
    if duration <= max_sec:
        return [(seg_start, seg_end, 0)]
    n = math.ceil(duration / max_sec)
    ov = int(overlap_sec * sample_rate)
    length = int(((duration + (n - 1) * overlap_sec) / n) * sample_rate)
    step = length - ov
    out = []
    for i in range(n):
        s = seg_start + i * step
        e = seg_end if i == n - 1 else min(s + length, seg_end)
        out.append((s, e, ov if i else 0))
    return out


def concat_vad_speech(audio: np.ndarray, vad_segments: Sequence[Tuple[int, int]]):
    if not len(vad_segments):
        return audio.copy(), [(0, 0, len(audio))]
    offset_map, pos = [], 0
    for s, e in vad_segments:
        offset_map.append((pos, s, e - s))
        pos += e - s
    return np.concatenate([audio[s:e] for s, e in vad_segments]), offset_map


class ConcatTimeMap:
    """offset_map with its concat-start column kept for bisection (a recording has thousands of words to map back)."""

    def __init__(self, offset_map: Sequence[Tuple[int, int, int]], sample_rate: int = 16000):
        self.map = list(offset_map)
        self.starts = [m[0] for m in self.map]
        self.sr = sample_rate

    def __call__(self, concat_time: float) -> float:
        if not self.map:
            return concat_time
        x = int(concat_time * self.sr)
        k = bisect.bisect_right(self.starts, x) - 1
        if k < 0:
            return self.map[0][1] / self.sr
        c0, o0, n = self.map[k]
        if x < c0 + n:
            return (o0 + x - c0) / self.sr
        _, o_last, n_last = self.map[-1]
        return (o_last + n_last) / self.sr


def map_concat_time_to_original(concat_time: float, offset_map: Sequence[Tuple[int, int, int]], sample_rate: int = 16000) -> float:
    return ConcatTimeMap(offset_map, sample_rate)(concat_time)


def normalize_word_for_overlap(word: str) -> str:
    word = unicodedata.normalize("NFC", word.lower().strip())
    return re.sub(r"[^\w]", "", word, flags=re.UNICODE)


@functools.lru_cache(maxsize=1 << 16)
def _fuzzy_match(w1: str, w2: str, threshold: float) -> bool:
    m = SequenceMatcher(None, w1, w2)
    # difflib's two upper bounds on ratio() first: most word pairs of an overlap window are unrelated
    return m.real_quick_ratio() >= threshold and m.quick_ratio() >= threshold and m.ratio() >= threshold


def words_match(w1: str, w2: str, threshold: float = FUZZY_MATCH_THRESHOLD) -> bool:
    if w1 == w2:
        return True
    if not w1 or not w2:
        return False
    if len(w1) > 2 and len(w2) > 2 and (w1 in w2 or w2 in w1):
        return True
    return _fuzzy_match(w1, w2, threshold)      # speech repeats its words: the same pairs come back in every overlap window


def _mean_prob(words: Sequence[dict]) -> float:
    return sum(w.get("prob", 1.0) for w in words) / max(1, len(words))


def find_overlap_alignment(tail_words: Sequence[dict], head_words: Sequence[dict]) -> Tuple[int, str, int]:
    """-> (first head index past the overlap, action in {cut_head, drop_tail, drop_head, none}, words to pop from the tail)."""
    if not tail_words or not head_words:
        return 0, "none", 0
    tail = [normalize_word_for_overlap(w["text"]) for w in tail_words[-MAX_OVERLAP_WORDS:]]
    head = [normalize_word_for_overlap(w["text"]) for w in head_words[:MAX_OVERLAP_WORDS]]
    nt, nh = len(tail), len(head)
    # pairwise fuzzy matches once; every offset then reads a diagonal
    match = [[words_match(t, h) for h in head] for t in tail]
    best_score, best_cut, best_pop = 0, 0, 0
    for offset in range(-nt + 1, nh):
        i_lo, i_hi = max(0, -offset), min(nt, nh - offset)       # tail indices whose partner i + offset is inside the head
        hits = [i for i in range(i_lo, i_hi) if match[i][i + offset]]
        window = min(nh, nt + offset) - max(0, offset)
        if len(hits) > best_score and len(hits) / max(1, window) >= MIN_MATCH_RATIO:
            best_score = len(hits)
            best_cut = hits[-1] + offset + 1
            best_pop = nt - 1 - hits[-1]
    diverged = best_score < min(nt, nh) and best_pop > 0
    if best_score == 0 or diverged:
        if best_score == 0:
            div_tail, div_head = list(tail_words), list(head_words)
        else:
            div_tail = list(tail_words[-best_pop:]) if best_pop > 0 else []
            div_head = list(head_words[best_cut:]) if best_cut < len(head_words) else []
        if _mean_prob(div_tail) > _mean_prob(div_head):
            return len(head_words), "drop_head", 0
        return 0, "drop_tail", len(tail_words)
    return best_cut, "cut_head", best_pop


def merge_chunks_with_overlap(chunk_results: Sequence[dict], overlap_duration_sec: float = OVERLAP_SEC) -> Tuple[List[dict], str]:
    """chunk_results: [{"words", "audio_start_abs", "audio_end_abs", ...}] -> (merged words, text)."""
    merged: List[dict] = []
    for idx, chunk in enumerate(chunk_results):
        words = chunk["words"]
        if idx == 0:
            merged.extend(words)
            continue
        prev = chunk_results[idx - 1]
        tail_from = max(0, (prev["audio_end_abs"] - prev["audio_start_abs"]) - overlap_duration_sec)
        tail = [w for w in prev["words"] if w.get("local_start", 0) >= tail_from]
        head = [w for w in words if w.get("local_start", 0) < overlap_duration_sec]
        cut, _action, pop = find_overlap_alignment(tail, head)
        if pop > 0:
            del merged[-pop:]
        merged.extend(words[cut:] if cut < len(words) else [])
    return merged, " ".join(w["text"] for w in merged)


class RecordingPlan:
    """What has to be decoded for one recording, and what is needed to put the answers back together."""

    def __init__(self, speech: np.ndarray, offset_map, plan: List[Tuple[int, int, int]], overlap_samples: int):
        self.speech, self.offset_map, self.plan, self.overlap_samples = speech, offset_map, plan, overlap_samples

    @property
    def chunks(self) -> List[np.ndarray]:
        return [self.speech[s:e] for s, e, _ in self.plan]

    @property
    def offsets(self) -> List[float]:
        return [s / 16000.0 for s, _, _ in self.plan]


def plan_recording(audio: np.ndarray, vad_segments: Sequence[Tuple[int, int]] = (), device_id: Optional[int] = None,
                   segment_samples: int = SEGMENT_SAMPLES, overlap_samples: int = OVERLAP_SAMPLES) -> RecordingPlan:
    """Speech-only concatenation of the VAD segments and the silence-aligned chunk plan over it
    (core/asr_engine.py:2130-2161). `device_id` not None: the energy scan runs on that GPU (csrc/energy.cu, same flags as
    NumPy); None keeps it on the host."""
    audio = np.ascontiguousarray(audio, dtype=np.float32)
    speech, offset_map = concat_vad_speech(audio, list(vad_segments))
    regions = find_silent_regions(speech) if device_id is None else find_silent_regions_gpu(speech, device_id=device_id)
    return RecordingPlan(speech, offset_map, plan_chunks(len(speech), regions, segment_samples, overlap_samples), overlap_samples)


def finish_recording(rp: RecordingPlan, per_chunk: List[List[dict]], per_chunk_rover: Optional[List[List[dict]]] = None,
                     hotword_phrases: Sequence[str] = ()) -> Dict[str, object]:
    """Word times back to the original recording, ROVER per chunk when a second model's words are given (combined in
    recording time, :2352-2357, :2469-2486), per-chunk results, overlap stitch (:2488-2494)."""
    to_original = ConcatTimeMap(rp.offset_map)
    for decoded in (per_chunk, per_chunk_rover or []):
        for words in decoded:
            for w in words:
                w["start"], w["end"] = to_original(w["start"]), to_original(w["end"])
    if per_chunk_rover is not None:
        from .asr_engine import rover_merge_words
        per_chunk = [rover_merge_words(a, b, hotword_phrases)[0] for a, b in zip(per_chunk, per_chunk_rover)]
    results = [{"text": " ".join(w["text"] for w in words), "words": words, "audio_start_abs": s / 16000.0,
                "audio_end_abs": e / 16000.0, "overlap_sec": o / 16000.0, "vad_group": 0}
               for words, (s, e, o) in zip(per_chunk, rp.plan)]
    if len(results) == 1:
        words = list(results[0]["words"])
    else:
        words, _ = merge_chunks_with_overlap(results, rp.overlap_samples / 16000.0)
    return {"words": words, "text": " ".join(w["text"] for w in words), "chunk_plan": rp.plan, "chunk_results": results}


def transcribe_long(recognizer, audio: np.ndarray, vad_segments: Sequence[Tuple[int, int]] = (), decode_chunks=None,
                    rover_recognizer=None, hotword_phrases: Sequence[str] = (), segment_samples: int = SEGMENT_SAMPLES,
                    overlap_samples: int = OVERLAP_SAMPLES) -> Dict[str, object]:
    """A whole recording, the way the reference's transcription phase runs it (core/asr_engine.py:2130-2161, :2326-2496):
    speech-only concatenation of the VAD segments, silence-aligned 30 s chunks with 3 s overlap, word times mapped back to
    the original recording, per-chunk word lists stitched. The per-chunk worker loop becomes ONE ragged GPU batch.
    Returns {"words", "text", "chunk_plan", "chunk_results"}."""
    # real engine -> the energy scan runs on the GPU as well; an injected decoder (tests on CPU) keeps the host scan
    device_id = None if decode_chunks is not None else int(recognizer.engine.device_id)
    rp = plan_recording(audio, vad_segments, device_id, segment_samples, overlap_samples)
    if decode_chunks is None:
        from .asr_engine import decode_chunks
    per_chunk = decode_chunks(recognizer, rp.chunks, rp.offsets)
    per_chunk_rover = decode_chunks(rover_recognizer, rp.chunks, rp.offsets) if rover_recognizer is not None else None
    return finish_recording(rp, per_chunk, per_chunk_rover, hotword_phrases)
