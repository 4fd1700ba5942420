"""What the reference does with the stitched word list right after the recognizer: suspect-word
flags from the search's own per-token statistics (tsallis_max / margin_min come out of the CUDA joiner epilogue, no second
pass over logits), acoustic gap checks, filler removal, and the asr_json wire format.

Host-side restatement (own code, same behaviour) of /root/reference:
  compute_disagree_indices  core/asr_engine.py:1677-1708   words where a second model's text differs (difflib opcodes)
  count_energy_peaks        :1619-1647                     syllable peaks of the 10 ms / 5 ms-hop RMS envelope
  _compute_gap_features     :1651-1674                     (energy range, 300-3000 Hz band ratio) of a gap
  suspect_detect            :1711-1865                     disagree OR (tsallis_max > 0.04 AND margin_min < 0.6) OR gap evidence
  remove_filler_words       :1587-1608
  post-ASR sequence         :2556-2580                     `finish_transcript`
  serialize_segments / deserialize_segments  core/asr_json.py:9-148, :151-223
The reference reads the Silero window probabilities from a module-level cache (core/vad_utils.py:51-55); here they are an
argument (`vad_probs`, one value per 512 samples, or None).
Parity: tests/test_postprocess.py (golden vectors from, and live runs against, the reference's functions).
"""
from __future__ import annotations

from datetime import datetime
from difflib import SequenceMatcher
from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple

import numpy as np

from .chunking import normalize_word_for_overlap

FILLER_WORDS = frozenset({"à", "ờ", "ừ", "ơ", "uh", "um"})       # core/asr_engine.py:1584

TSALLIS_TH, MARGIN_TH, ENTROPY_TH, TSALLIS_ALONE_TH = 0.04, 0.6, 0.10, 0.12      # :1777-1779, :1800
GAP_MIN_MS, GAP_VAD_TH, GAP_ERANGE_TH, GAP_LONG_MS, GAP_PEAKS_TH = 200, 0.90, 0.04, 500, 3   # :1809-1813


def compute_disagree_indices(words_main: Sequence[dict], words_other_text: Sequence[str]) -> Set[int]:
    main = [normalize_word_for_overlap(w["text"]) for w in words_main]
    other = [normalize_word_for_overlap(w) for w in words_other_text]
    out: Set[int] = set()
    for tag, i1, i2, _j1, _j2 in SequenceMatcher(None, main, other).get_opcodes():
        if tag == "equal":
            continue
        out.update(range(i1, i2))
        if tag == "insert":                      # only the other model has words here: flag both neighbours
            if i1 > 0:
                out.add(i1 - 1)
            if i1 < len(main):
                out.add(i1)
    return out


def _frame_rms(x: np.ndarray, sr: int) -> np.ndarray:
    """RMS of 10 ms frames at a 5 ms hop, in the dtype NumPy gives `np.sqrt(np.mean(frame ** 2))` for x's dtype."""
    frame, hop = int(sr * 0.010), int(sr * 0.005)
    if len(x) < frame:                           # one short frame: the whole segment
        return np.array([np.sqrt(np.mean(x ** 2))]) if len(x) else np.zeros(1, dtype=x.dtype if x.dtype.kind == "f" else np.float64)
    n = (len(x) - frame) // hop + 1
    w = np.lib.stride_tricks.sliding_window_view(x, frame)[::hop][:n]
    return np.sqrt(np.mean(w ** 2, axis=1))


def _peaks_from_envelope(energy: np.ndarray, sr: int, threshold_factor: float) -> List[float]:
    from scipy.signal import find_peaks
    hop = int(sr * 0.005)
    kernel = np.hanning(7)
    kernel /= kernel.sum()
    smooth = np.convolve(np.asarray(energy, dtype=np.float64), kernel, mode="same")
    voiced = smooth[smooth > np.max(smooth) * 0.05]
    if len(voiced) == 0:
        return []
    th = np.mean(voiced) * threshold_factor
    peaks, _ = find_peaks(smooth, distance=int(90 / (hop / sr * 1000)), height=th, prominence=th * 0.3)
    return (peaks * hop / sr).tolist()


def count_energy_peaks(audio_segment: np.ndarray, sr: int = 16000, threshold_factor: float = 1.0) -> List[float]:
    return _peaks_from_envelope(_frame_rms(np.asarray(audio_segment), sr), sr, threshold_factor)


def _band_ratio(x: np.ndarray, sr: int) -> float:
    n_fft = min(512, len(x))
    mag2 = np.abs(np.fft.rfft(x[:n_fft] * np.hanning(n_fft))) ** 2
    freqs = np.fft.rfftfreq(n_fft, 1.0 / sr)
    return float(np.sum(mag2[(freqs >= 300) & (freqs <= 3000)]) / (np.sum(mag2) + 1e-10))


def compute_gap_features(audio_segment: np.ndarray, sr: int = 16000) -> Tuple[float, float]:
    x = np.asarray(audio_segment)
    if len(x) < 50:
        return 0.0, 0.0
    e = _frame_rms(x, sr)
    return float(np.max(e) - np.min(e)), _band_ratio(x, sr)


def suspect_detect(all_words: List[dict], audio: np.ndarray, disagree_indices: Optional[Iterable[int]] = None,
                   vad_probs: Optional[np.ndarray] = None, sr: int = 16000) -> List[dict]:
    """Tags `_suspect_level = "warning"` (and gap_after_ms / gap_before_ms) in place; returns all_words."""
    n = len(all_words)
    if n < 2:
        return all_words
    disagree = set(disagree_indices) if disagree_indices else set()
    has_tsallis = any(w.get("tsallis_max") is not None for w in all_words)
    has_entropy = any(w.get("entropy_norm") is not None for w in all_words)
    has_margin = any(w.get("margin_min") is not None for w in all_words)

    def uncertain(w: dict) -> bool:
        if has_tsallis:
            ts, mg = w.get("tsallis_max"), w.get("margin_min")
            if ts is None or not ts > TSALLIS_TH:
                return False
            if has_margin and mg is not None:
                return mg < MARGIN_TH
            return ts > TSALLIS_ALONE_TH
        if has_entropy:
            ent = w.get("entropy_norm")
            return ent is not None and ent > ENTROPY_TH
        return False

    flagged = [i in disagree or uncertain(w) for i, w in enumerate(all_words)]

    gap_after: Set[int] = set()
    for i in range(n - 1):
        cur, nxt = all_words[i], all_words[i + 1]
        gap_ms = (nxt["start"] - cur["end"]) * 1000
        if gap_ms < GAP_MIN_MS:
            continue
        gs, ge = int(cur["end"] * sr), int(nxt["start"] * sr)
        if gs >= ge or gs < 0 or ge > len(audio) or ge - gs < 80:
            continue
        vad_max = 0.0
        if vad_probs is not None:
            w0 = max(0, min(gs // 512, len(vad_probs) - 1))
            w1 = max(w0 + 1, min(ge // 512, len(vad_probs)))
            window = vad_probs[w0:w1]
            if len(window):
                vad_max = float(np.max(window))
        if vad_max < GAP_VAD_TH:                 # without speech evidence neither the peaks nor the range can flag the gap
            continue
        envelope = _frame_rms(np.asarray(audio[gs:ge]), sr)      # one envelope serves both the peak count and the range
        if float(np.max(envelope) - np.min(envelope)) < GAP_ERANGE_TH:
            continue
        if gap_ms >= GAP_LONG_MS or len(_peaks_from_envelope(envelope, sr, 1.0)) >= GAP_PEAKS_TH:
            gap_after.add(i)
            cur["gap_after_ms"] = int(gap_ms)
            nxt["gap_before_ms"] = int(gap_ms)

    for i, w in enumerate(all_words):
        if flagged[i] or i in gap_after or (i - 1) in gap_after:
            w["_suspect_level"] = "warning"
    return all_words


def remove_filler_words(words: List[dict]) -> List[dict]:
    if not words:
        return words
    return [w for w in words if w["text"].lower() not in FILLER_WORDS]


def finish_transcript(all_words: List[dict], audio: np.ndarray, is_rover: bool = False,
                      vad_probs: Optional[np.ndarray] = None) -> Tuple[List[dict], str]:
    """The post-ASR sequence of core/asr_engine.py:2556-2580: disagreement set rebuilt from the `_disagree` flags ROVER left
    on the words, suspect detection, filler removal, capitalised full text."""
    disagree = None
    if is_rover:
        found = {i for i, w in enumerate(all_words) if w.get("_disagree")}
        for i in found:
            all_words[i].pop("_disagree", None)
        disagree = found or None
    words = remove_filler_words(suspect_detect(all_words, audio, disagree, vad_probs))
    text = " ".join(w["text"] for w in words)
    return words, text.capitalize() if text else text


# ----------------------------------------------------------------------------- asr_json (core/asr_json.py)
def _speaker_id_value(sid):
    if isinstance(sid, (int, float)) or (isinstance(sid, str) and sid.isdigit()):
        return int(sid)
    return sid


def _raw_word(w: dict) -> dict:
    out = {"text": w.get("text", "")}
    for key in ("start", "end"):
        if key in w:
            try:
                out[key] = round(float(w.get(key, 0)), 3)
            except (TypeError, ValueError):
                pass
    for key in ("gap_after_ms", "gap_before_ms"):
        if w.get(key):
            out[key] = w[key]
    if w.get("_suspect_level"):
        out["suspect"] = w["_suspect_level"]
    return out


def serialize_segments(segments: Sequence[dict], speaker_name_mapping: Optional[Dict[str, str]] = None,
                       speaker_colors: Optional[Dict[str, str]] = None, model_name: str = "unknown", model_type: str = "file",
                       duration_sec: float = 0.0, timing: Optional[dict] = None,
                       overlap_segments: Optional[Sequence[dict]] = None) -> dict:
    names = speaker_name_mapping or {}
    entries: List[dict] = []
    shown = None
    for i, seg in enumerate(segments):
        sid = seg.get("speaker_id", 0)
        display = names.get(str(sid), seg.get("speaker", ""))
        start = seg.get("start", seg.get("start_time", 0))
        if display and display != shown:
            entries.append({"type": "speaker", "speaker": display, "speaker_id": _speaker_id_value(sid), "start_time": start})
            shown = display
        partials = [{"text": p.get("text", ""), "timestamp": p.get("timestamp", 0)} for p in seg.get("partials", [])]
        if not partials:
            partials = [{"text": seg.get("text", ""), "timestamp": seg.get("end", seg.get("start", 0) + 1.0)}]
        entry = {"type": "text", "text": seg.get("text", ""), "start_time": start, "segment_id": i, "partials": partials}
        if seg.get("raw_words"):
            entry["raw_words"] = [_raw_word(w) for w in seg["raw_words"]]
        entries.append(entry)
    data = {"version": 1, "model": model_name, "model_type": model_type, "created_at": datetime.now().isoformat(),
            "duration_sec": round(duration_sec, 2), "timing": timing or {}, "speaker_names": dict(names),
            "speaker_colors": dict(speaker_colors) if speaker_colors else {}, "segments": entries}
    if overlap_segments:
        out = []
        for ov in overlap_segments:
            sid = ov.get("speaker_id", 0)
            display = ov.get("speaker", f"Người nói {sid + 1}") if "speaker" in ov else f"Người nói {sid + 1}"
            display = names.get(str(sid), display)
            e = {"speaker": display, "speaker_id": int(sid) if isinstance(sid, (int, float)) else sid,
                 "start_time": round(float(ov.get("start", 0)), 3), "end_time": round(float(ov.get("end", 0)), 3),
                 "text": ov.get("text", "")}
            if ov.get("raw_words"):
                e["raw_words"] = [{"text": w.get("word") or w.get("text") or "", "start": round(float(w.get("start", 0)), 3),
                                   "end": round(float(w.get("end", 0)), 3)} for w in ov["raw_words"]]
            out.append(e)
        data["overlap_segments"] = out
    return data


def deserialize_segments(data: dict):
    """-> (segments, speaker_mapping, speaker_colors, has_speakers)"""
    if "segments" not in data:
        raise ValueError("Invalid JSON: no 'segments' key")
    segments: List[dict] = []
    speaker, speaker_id, has_speakers = "", 0, False
    for seg in data["segments"]:
        kind = seg.get("type", "text")
        if kind == "speaker":
            speaker = seg.get("speaker", "")
            raw = seg.get("speaker_id", 0)
            try:
                speaker_id = int(raw)
            except (ValueError, TypeError):
                speaker_id = raw
            has_speakers = True
        elif kind == "text":
            text = seg.get("text", "")
            start = seg.get("start_time", 0)
            partials = [p for p in seg.get("partials", []) if p.get("text", "").strip()]
            if not partials and text:
                partials = [{"text": text}]
            end = partials[-1].get("timestamp", start + 1.0) if partials else start + 1.0
            item = {"text": text, "start": start, "start_time": start, "index": len(segments), "speaker": speaker,
                    "speaker_id": speaker_id, "partials": partials or [{"text": text, "timestamp": end}], "end": end}
            if seg.get("raw_words"):
                item["raw_words"] = list(seg["raw_words"])
            segments.append(item)
    return segments, data.get("speaker_names", {}), data.get("speaker_colors", {}), has_speakers
