"""Host logic around the voice-activity model: everything the reference does before and after
the Silero network itself, with the network as a pluggable batch function. The reference runs the ONNX model one 512-sample
window at a time (31 sequential session calls per audio-second); the layout here is what a batched GPU recurrence consumes:
all windows of a recording as one [n_windows, 64 + 512] matrix (context + window), probabilities back as one vector.

Host-side restatement (own code, same behaviour) of /root/reference:
  window layout + state carry   core/vad_utils.py:80-106    `window_matrix`
  probabilities -> segments     :121-151                    `segments_from_probs`
  get_vad_segments              :158-263                    low-level boost to -23 dBFS, retry at 0.3, fallback, 1 s padding,
                                                            250 ms merge
  5 s gap merge                 core/asr_engine.py:2115-2128  `merge_close_segments`
`prob_fn(windows[n, 576]) -> probs[n]` is the seam; `GpuVad` (csrc/vad.cu) is the network behind it on the GPU. The Silero
weights are not available offline, so it runs seeded random tensors of that architecture (weights.init_vad_weights).
Parity: tests/test_vad_logic.py drives the reference's functions and these with the same stand-in network.
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

WINDOW = 512
CONTEXT = 64
VAD_BOOST_TARGET = 0.071       # -23 dBFS (core/vad_utils.py:203)
MAX_VAD_GAP = 5 * 16000        # core/asr_engine.py:2117

ProbFn = Callable[[np.ndarray], np.ndarray]


class GpuVad:
    """The voice-activity network on the GPU (csrc/vad.cu through B200AsrVad*): all windows of a recording - or of a batch of
    recordings - in one call, the LSTM state carried on the device, where the reference makes 31 sequential onnxruntime calls
    per audio-second (core/vad_utils.py:97-104)."""

    def __init__(self, weights_path: str, device_id: int = 0):
        from . import _capi
        self._capi = _capi
        self._h = _capi.lib().B200AsrVadCreate(os.fsencode(weights_path), int(device_id))
        if not self._h:
            raise RuntimeError("B200AsrVadCreate failed: " + _capi.last_error())

    def probs(self, audio: np.ndarray) -> np.ndarray:
        """Speech probability of every 512-sample window of one recording (zero state at its start)."""
        c = self._capi
        x = np.ascontiguousarray(audio, dtype=np.float32).reshape(-1)
        out = np.empty(len(x) // WINDOW, dtype=np.float32)
        if out.size == 0:
            return out
        if c.lib().B200AsrVadProbs(self._h, c.fptr(x), len(x), c.fptr(out)) < 0:
            raise RuntimeError(c.last_error())
        return out

    def probs_batch(self, recordings: Sequence[np.ndarray]) -> List[np.ndarray]:
        c = self._capi
        offs = np.zeros(len(recordings) + 1, dtype=np.int64)
        offs[1:] = np.cumsum([len(r) for r in recordings])
        x = np.ascontiguousarray(np.concatenate([np.asarray(r, dtype=np.float32).reshape(-1) for r in recordings])
                                 if len(recordings) else np.zeros(0, np.float32))
        poffs = np.zeros(len(recordings) + 1, dtype=np.int64)
        n = c.lib().B200AsrVadProbsBatch(self._h, c.fptr(x), c.i64ptr(offs), len(recordings), None, c.i64ptr(poffs))
        if n < 0:
            raise RuntimeError(c.last_error())
        out = np.empty(max(n, 0), dtype=np.float32)
        if n and c.lib().B200AsrVadProbsBatch(self._h, c.fptr(x), c.i64ptr(offs), len(recordings), c.fptr(out), c.i64ptr(poffs)) < 0:
            raise RuntimeError(c.last_error())
        return [out[poffs[i]:poffs[i + 1]] for i in range(len(recordings))]

    def prob_fn(self) -> "ProbFn":
        """The `get_vad_segments(prob_fn=...)` seam: rows[n, 576] (context + window, one recording in order) -> probs[n]. The rows
        overlap by construction (window_matrix), so the recording is rebuilt from their window parts."""
        def fn(rows: np.ndarray) -> np.ndarray:
            rows = np.asarray(rows, dtype=np.float32)
            return self.probs(np.ascontiguousarray(rows[:, CONTEXT:]).reshape(-1))
        return fn

    def last_timings(self) -> dict:
        import ctypes as C
        a, b = C.c_float(0), C.c_float(0)
        self._capi.lib().B200AsrVadLastTimings(self._h, C.byref(a), C.byref(b))
        return {"frontend_ms": a.value, "recurrence_ms": b.value}

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._capi.lib().B200AsrVadDestroy(self._h)
                self._h = None
        except Exception:
            pass


def window_matrix(audio: np.ndarray) -> np.ndarray:
    """[n_windows, 576] float32: row i = last 64 samples of window i-1 (zeros for i = 0) followed by window i."""
    n = len(audio) // WINDOW
    x = np.zeros((n, CONTEXT + WINDOW), dtype=np.float32)
    if n:
        w = np.asarray(audio[: n * WINDOW], dtype=np.float32).reshape(n, WINDOW)
        x[:, CONTEXT:] = w
        x[1:, :CONTEXT] = w[:-1, -CONTEXT:]
    return x


def segments_from_probs(probs: Sequence[float], threshold: float = 0.5, min_silence_ms: float = 300, min_speech_ms: float = 250,
                        sample_rate: int = 16000) -> List[Tuple[int, int]]:
    """Window-index segments: speech starts at the first window at or over the threshold, ends where a run of
    min_silence windows under it begins; segments shorter than min_speech windows are dropped."""
    p = np.asarray(probs, dtype=np.float64)
    min_sil = int(min_silence_ms * sample_rate / 1000 / WINDOW)
    min_sp = int(min_speech_ms * sample_rate / 1000 / WINDOW)
    out: List[Tuple[int, int]] = []
    start, quiet = None, 0
    for i, speech in enumerate((p >= threshold).tolist()):
        if speech:
            if start is None:
                start = i
            quiet = 0
        elif start is not None:
            quiet += 1
            if quiet >= min_sil:
                end = i - quiet + 1
                if end - start >= min_sp:
                    out.append((start, end))
                start, quiet = None, 0
    if start is not None and len(p) - start >= min_sp:
        out.append((start, len(p)))
    return out


def merge_close_segments(segments: Sequence[Tuple[int, int]], max_gap: int, inclusive: bool) -> List[Tuple[int, int]]:
    """Joins neighbours whose gap is < max_gap (or <= when inclusive)."""
    out: List[Tuple[int, int]] = []
    for s, e in segments:
        if out and (s - out[-1][1] <= max_gap if inclusive else s - out[-1][1] < max_gap):
            out[-1] = (out[-1][0], e)
        else:
            out.append((s, e))
    return out


def get_vad_segments(audio: np.ndarray, prob_fn: ProbFn, sample_rate: int = 16000, threshold: float = 0.2, min_silence_ms: float = 100,
                     min_speech_ms: float = 250, padding_ms: float = 1000, merge_gap_ms: float = 250, auto_boost: bool = True,
                     fallback_full: bool = True) -> Tuple[List[Tuple[int, int]], Optional[np.ndarray]]:
    """-> (speech segments in samples, window probabilities of the last network pass or None). The probabilities are what
    the reference caches for suspect_detect (core/vad_utils.py:118-119)."""
    total = len(audio)
    if total < WINDOW:
        return ([(0, total)] if fallback_full else []), None
    x = audio
    if auto_boost:
        peak = np.max(np.abs(audio))
        if 1e-6 < peak < VAD_BOOST_TARGET:
            x = (audio * (VAD_BOOST_TARGET / peak)).astype(np.float32)
    probs = np.asarray(prob_fn(window_matrix(x)), dtype=np.float32)
    found = segments_from_probs(probs.tolist(), threshold, min_silence_ms, min_speech_ms, sample_rate)
    if not found:                                  # second look with a lower bar; same windows, so the network is not rerun
        found = segments_from_probs(probs.tolist(), 0.3, 100, 150, sample_rate)
    if not found:
        return ([(0, total)] if fallback_full else []), probs
    pad = int(padding_ms * sample_rate / 1000)
    padded = [(max(0, s * WINDOW - pad), min(total, e * WINDOW + pad)) for s, e in found]
    if merge_gap_ms > 0 and len(padded) > 1:
        padded = merge_close_segments(padded, int(merge_gap_ms * sample_rate / 1000), inclusive=False)
    return padded, probs


def speech_plan(audio: np.ndarray, prob_fn: ProbFn) -> Tuple[List[Tuple[int, int]], Optional[np.ndarray]]:
    """The VAD phase of the transcription pipeline (core/asr_engine.py:2090-2128): segments, then the 5 s gap merge."""
    segments, probs = get_vad_segments(audio, prob_fn)
    return merge_close_segments(segments, MAX_VAD_GAP, inclusive=True), probs
