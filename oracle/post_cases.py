"""TEST INFRASTRUCTURE ONLY. Seeded inputs for tests/test_postprocess.py and the golden vectors oracle/make_golden.py
writes from the reference's own functions (/root/reference core/asr_engine.py:1587-1865, core/asr_json.py:9-223)."""
import numpy as np

from oracle.chunk_cases import SYLLABLES, silence_audio
from sherpa_vietnamese_asr_b200.synth import speech_like

FILLERS = ["à", "ờ", "ừ", "ơ", "uh", "um", "À", "Um"]


def suspect_case(seed, seconds=40.0, stats="tsallis+margin", with_disagree=False, with_vad=True, dtype=np.float32):
    """A word list over a bursty recording: words sit on the bursts, some gaps between them hold unclaimed sound."""
    rng = np.random.default_rng(seed)
    if seed % 2:
        audio = silence_audio(seed, seconds).astype(dtype)
    else:                                                   # syllabic amplitude modulation: several energy peaks per gap
        audio = speech_like(int(seconds * 16000), seed).astype(dtype)
    n = len(audio)
    words, t = [], float(rng.uniform(0, 0.5))
    while t < seconds - 1.0:
        dur = float(rng.uniform(0.08, 0.5))
        w = {"text": (SYLLABLES + FILLERS)[int(rng.integers(len(SYLLABLES) + len(FILLERS)))], "start": round(t, 4),
             "end": round(min(t + dur, seconds), 4), "prob": round(float(rng.uniform(0.2, 1.0)), 4)}
        if stats in ("tsallis+margin", "tsallis"):
            w["tsallis_max"] = None if rng.uniform() < 0.05 else round(float(rng.exponential(0.05)), 4)
        if stats == "tsallis+margin":
            w["margin_min"] = None if rng.uniform() < 0.05 else round(float(rng.uniform(0, 1)), 4)
        if stats == "entropy":
            w["entropy_norm"] = round(float(rng.exponential(0.08)), 4)
        words.append(w)
        r = rng.uniform()
        gap = 0.0 if r < 0.3 else (float(rng.uniform(0.0, 0.3)) if r < 0.7 else float(rng.uniform(0.3, 2.5)))
        t += dur + gap
    # a few degenerate gaps: overlapping words, a gap shorter than 80 samples is impossible above 200 ms, so add one by hand
    if len(words) > 6:
        words[3]["start"] = round(words[2]["end"] - 0.05, 4)
    vad = None
    if with_vad:
        vad = np.clip(rng.normal(0.7, 0.35, n // 512 + 1), 0, 1).astype(np.float32)
    disagree = None
    if with_disagree:
        disagree = set(int(i) for i in rng.choice(len(words), max(1, len(words) // 10), replace=False))
    return words, audio, disagree, vad


def suspect_cases():
    return [suspect_case(1), suspect_case(2, 25.0, "tsallis"), suspect_case(3, 25.0, "entropy"), suspect_case(4, 30.0, "none"),
            suspect_case(5, 60.0, with_disagree=True), suspect_case(6, 20.0, with_vad=False),
            suspect_case(7, 30.0, dtype=np.float64), (suspect_case(8, 5.0)[0][:1],) + suspect_case(8, 5.0)[1:]]


def gap_segments(seed=21, n=40):
    rng = np.random.default_rng(seed)
    audio = np.concatenate([silence_audio(seed, 60.0), speech_like(60 * 16000, seed)])
    out = []
    for k in range(n):
        ln = int(rng.choice([30, 49, 50, 79, 80, 120, 159, 160, 161, 239, 240, 400, 511, 512, 513, 1600, 8000, 24000, 40000]))
        s = int(rng.integers(0, len(audio) - ln))
        seg = audio[s:s + ln]
        out.append(seg.astype(np.float64) if k % 5 == 4 else seg)
    out.append(np.zeros(3200, np.float32))
    return out


def disagree_cases(seed=31, n=30):
    rng = np.random.default_rng(seed)
    out = [([], []), ([{"text": "a"}], []), ([], ["a"])]
    for _ in range(n):
        main = [SYLLABLES[int(rng.integers(len(SYLLABLES)))] for _ in range(int(rng.integers(1, 40)))]
        other = []
        for w in main:
            r = rng.uniform()
            if r < 0.1:
                continue
            other.append(SYLLABLES[int(rng.integers(len(SYLLABLES)))] if r < 0.2 else (w.upper() if r < 0.3 else w))
            if r > 0.92:
                other.append(SYLLABLES[int(rng.integers(len(SYLLABLES)))])
        out.append(([{"text": w} for w in main], other))
    return out


def segment_cases(seed=41):
    """Internal segments as the pipeline hands them to serialize_segments, with and without speakers / partials / raw words."""
    rng = np.random.default_rng(seed)
    cases = []
    for c in range(8):
        segs, t = [], 0.0
        for i in range(int(rng.integers(0, 12))):
            words, _, _, _ = suspect_case(seed * 100 + c * 20 + i, 6.0)
            for w in words[::3]:
                w["_suspect_level"] = "warning"
            for w in words[1::4]:
                w["gap_after_ms"] = int(rng.integers(0, 900))
                w["gap_before_ms"] = int(rng.integers(0, 900))
            seg = {"text": " ".join(w["text"] for w in words), "start": round(t, 3), "end": round(t + 6.0, 3)}
            if c % 2:
                sid = int(rng.integers(0, 3))
                seg["speaker_id"] = sid if c != 3 else str(sid)
                seg["speaker"] = f"Người nói {sid + 1}"
            if c == 5:
                seg["speaker_id"], seg["speaker"] = "guest", "Khách"
            if c % 3 == 0:
                seg["partials"] = [{"text": w["text"], "timestamp": w["end"], "extra": 1} for w in words[:5]]
            if c % 4 != 1:
                seg["raw_words"] = words if c != 6 else words + [{"text": "x", "start": "bad", "end": None}]
            if c == 7:
                seg = {"text": seg["text"], "start_time": seg["start"]}
            segs.append(seg)
            t += 6.0
        mapping = {"0": "Anh A", "2": "Chị C"} if c in (1, 3) else None
        colors = {"0": "#ff0000"} if c == 1 else None
        overlaps = None
        if c in (1, 2):
            overlaps = [{"speaker_id": 0, "start": 1.23456, "end": 2.5, "text": "xin chào",
                         "raw_words": [{"word": "xin", "start": 1.2, "end": 1.5}, {"text": "chào", "start": 1.5, "end": 2.0}, {}]},
                        {"speaker_id": 2, "speaker": "B", "start": 3, "end": 4}]
        cases.append({"segments": segs, "speaker_name_mapping": mapping, "speaker_colors": colors, "model_name": f"m{c}",
                      "model_type": "file" if c % 2 else "online", "duration_sec": round(t, 5) + 0.004,
                      "timing": {"asr": 1.5} if c % 2 else None, "overlap_segments": overlaps})
    return cases
