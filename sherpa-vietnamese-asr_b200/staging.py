"""Audio staging between the decoder and the recognizer, host side: PCM ingest, the low-volume
boost applied at load time, per-segment RMS normalisation and the peak limiter.

Host-side restatement (own code, same behaviour) of /root/reference:
  load_audio tail            core/asr_engine.py:493-517      mono mix-down, peak < 0.5 -> scaled to 0.95
  compute_segment_rms        core/audio_preprocessing.py:39-43
  per_segment_rms_normalize  :46-155    gain = median segment RMS / segment RMS, clamped to +-20 dB, 5 ms linear fades
  adaptive_peak_limit        :226-244   peak > 0.95 -> linear scale to 0.95
  preprocess_audio           :251-292   RMS normalise (optional) then peak limit
File decoding (soundfile / ffmpeg, :457-510) is outside the path: `ingest_pcm` starts from decoded samples.
Parity: tests/test_staging.py (live against the reference's functions, and golden vectors from them).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def ingest_pcm(samples: np.ndarray) -> np.ndarray:
    """Decoded PCM -> mono float32 in [-1, 1): int16 is scaled by 1/32768 (what soundfile's float32 read and ffmpeg's f32le
    output give), channels (frames, ch) are averaged (:498-499)."""
    x = np.asarray(samples)
    if x.dtype == np.int16:
        x = x.astype(np.float32) / np.float32(32768.0)
    elif x.dtype != np.float32:
        x = x.astype(np.float32)
    if x.ndim == 2:
        x = x.mean(axis=1)
    return x


def boost_low_volume(audio: np.ndarray) -> np.ndarray:
    peak = np.max(np.abs(audio)) if len(audio) else 0.0
    if 0 < peak < 0.5:
        return audio / peak * 0.95
    return audio


def compute_segment_rms(segment: np.ndarray) -> float:
    return float(np.sqrt(np.mean(segment ** 2))) if len(segment) else 0.0


def segment_gains(audio: np.ndarray, vad_segments: Sequence[Tuple[int, int]], sample_rate: int = 16000, min_segment_ms: float = 100,
                  max_gain_db: float = 20.0) -> List[Tuple[int, int, float]]:
    """[(start, end, gain)] for the segments that take part: at least min_segment_ms long and not silent. Empty when there is
    nothing to normalise."""
    min_samples = int(min_segment_ms * sample_rate / 1000)
    measured = []
    for s, e in vad_segments:
        if e - s >= min_samples:
            rms = compute_segment_rms(audio[s:e])
            if rms > 1e-8:
                measured.append((s, e, rms))
    if not measured:
        return []
    target = float(np.median(np.array([r for _, _, r in measured])))
    if target < 1e-8:
        return []
    hi = 10 ** (max_gain_db / 20.0)
    return [(s, e, max(min(target / r, hi), 1.0 / hi)) for s, e, r in measured]


def gain_curve(n: int, gains: Sequence[Tuple[int, int, float]], sample_rate: int = 16000, crossfade_ms: float = 5) -> np.ndarray:
    """Per-sample gain: 1 outside the segments, the segment gain inside, linear ramps of up to crossfade_ms (a quarter of the
    segment at most) to the neighbouring value at both segment edges, applied segment by segment in order."""
    curve = np.ones(n, dtype=np.float32)
    for s, e, g in gains:
        curve[s:e] = g
    fade = int(crossfade_ms * sample_rate / 1000)
    if fade > 0:
        for s, e, _ in gains:
            k = min(fade, (e - s) // 4)
            if k <= 0:
                continue
            if s > 0:
                curve[s:s + k] = np.linspace(curve[s - 1], curve[s], k, dtype=np.float32)
            if e < n:
                curve[e - k:e] = np.linspace(curve[e - 1], curve[min(n - 1, e)], k, dtype=np.float32)
    return curve


def per_segment_rms_normalize(audio: np.ndarray, vad_segments: Sequence[Tuple[int, int]], sample_rate: int = 16000,
                              min_segment_ms: float = 100, max_gain_db: float = 20.0, crossfade_ms: float = 5) -> np.ndarray:
    if len(vad_segments) == 0:
        return audio
    gains = segment_gains(audio, vad_segments, sample_rate, min_segment_ms, max_gain_db)
    if not gains:
        return audio
    return audio * gain_curve(len(audio), gains, sample_rate, crossfade_ms)


def adaptive_peak_limit(audio: np.ndarray, target_peak: float = 0.95) -> np.ndarray:
    peak = np.max(np.abs(audio))
    return audio * (target_peak / peak) if peak > target_peak else audio


def preprocess_audio(audio: np.ndarray, vad_segments: Sequence[Tuple[int, int]], sample_rate: int = 16000,
                     enable_rms_normalize: bool = True) -> np.ndarray:
    out = audio.copy()
    if enable_rms_normalize and len(vad_segments) > 0:
        out = per_segment_rms_normalize(out, vad_segments, sample_rate)
    return adaptive_peak_limit(out)


def preprocess_audio_gpu(audio: np.ndarray, vad_segments: Sequence[Tuple[int, int]], sample_rate: int = 16000,
                         enable_rms_normalize: bool = True, boost_low: bool = False, device_id: int = 0) -> np.ndarray:
    """preprocess_audio on the GPU (csrc/staging.cu through B200AsrPreprocessAudio): the per-sample passes run over the
    uploaded PCM, the host only sees a few numbers per VAD segment. `boost_low` also applies boost_low_volume first."""
    from . import _capi
    x = np.ascontiguousarray(audio, dtype=np.float32).reshape(-1)
    out = np.empty_like(x)
    seg = np.asarray(list(vad_segments), dtype=np.int64).reshape(-1, 2)
    s = np.ascontiguousarray(seg[:, 0]) if len(seg) else np.zeros(1, np.int64)
    e = np.ascontiguousarray(seg[:, 1]) if len(seg) else np.zeros(1, np.int64)
    rc = _capi.lib().B200AsrPreprocessAudio(_capi.fptr(x), len(x), _capi.i64ptr(s), _capi.i64ptr(e), len(seg), int(bool(enable_rms_normalize)),
                                            int(bool(boost_low)), int(sample_rate), _capi.fptr(out), int(device_id))
    if rc != 0:
        raise RuntimeError("B200AsrPreprocessAudio failed: " + _capi.last_error())
    return out
