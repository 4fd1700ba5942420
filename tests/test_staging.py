"""CPU suite for the audio staging steps against the reference's own functions: live when
/root/reference is present, and through golden checksums written from them (oracle/make_golden.py)."""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import pytest

from oracle import chunk_cases as cc
from sherpa_vietnamese_asr_b200 import staging as st

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference"
have_ref = os.path.isdir(os.path.join(REF, "core"))


def staging_cases():
    """(audio, vad_segments): bursts at very different levels, segments incl. too short / silent / touching / at the edges."""
    rng = np.random.default_rng(17)
    cases = []
    for c in range(8):
        n = int(rng.uniform(3, 40) * 16000)
        audio = cc.silence_audio(100 + c, n / 16000.0)
        segs, p = [], 0 if c % 2 else int(rng.integers(0, 8000))
        while p < n:
            ln = int(rng.choice([800, 1599, 1600, 2000, 16000, 48000, 100000]))
            e = min(n, p + ln)
            audio[p:e] *= np.float32(rng.choice([1e-10, 0.02, 0.3, 1.0, 4.0, 60.0]))
            segs.append((p, e))
            p = e + int(rng.choice([0, 0, 1, 40, 400, 8000]))
        if c == 5:
            audio *= np.float32(0.01)
        if c == 6:
            segs = []
        cases.append((audio.astype(np.float32), segs))
    cases.append((np.zeros(16000, np.float32), [(0, 16000)]))
    return cases


@pytest.fixture(scope="module")
def ref():
    if not have_ref:
        pytest.skip("/root/reference not present (GPU box)")
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with redirect_stdout(io.StringIO()):
        import core.audio_preprocessing as ap
    return ap


def _digest(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    return [int(len(x)), float(np.sum(x.astype(np.float64))), float(np.sum(np.abs(x).astype(np.float64))),
            float(np.max(np.abs(x))) if len(x) else 0.0]


def test_staging_golden():
    with open(os.path.join(GOLD, "staging.json"), encoding="utf-8") as f:
        gold = json.load(f)
    for (audio, segs), want in zip(staging_cases(), gold):
        assert _digest(st.per_segment_rms_normalize(audio.copy(), segs)) == want["rms"]
        assert _digest(st.preprocess_audio(audio, segs)) == want["pre"]
        assert _digest(st.preprocess_audio(audio, segs, enable_rms_normalize=False)) == want["limit"]


def test_staging_live(ref):
    for audio, segs in staging_cases():
        for kw in ({}, {"min_segment_ms": 50, "max_gain_db": 6.0, "crossfade_ms": 0}, {"crossfade_ms": 20}):
            got = st.per_segment_rms_normalize(audio.copy(), segs, **kw)
            want = ref.per_segment_rms_normalize(audio.copy(), segs, **kw)
            assert got.dtype == want.dtype and np.array_equal(got, want)
        assert np.array_equal(st.preprocess_audio(audio, segs), ref.preprocess_audio(audio, segs))
        assert np.array_equal(st.adaptive_peak_limit(audio * 3), ref.adaptive_peak_limit(audio * 3))
        assert st.compute_segment_rms(audio[:5000]) == ref.compute_segment_rms(audio[:5000])
    assert st.compute_segment_rms(np.zeros(0, np.float32)) == ref.compute_segment_rms(np.zeros(0, np.float32)) == 0.0


def test_ingest_and_low_volume_boost():
    rng = np.random.default_rng(3)
    pcm = rng.integers(-32768, 32767, 16000, dtype=np.int16)
    x = st.ingest_pcm(pcm)
    assert x.dtype == np.float32 and np.array_equal(x, pcm.astype(np.float32) / 32768.0) and np.abs(x).max() <= 1.0
    stereo = np.stack([pcm, pcm[::-1]], axis=1)
    m = st.ingest_pcm(stereo)
    assert m.shape == (16000,) and np.allclose(m, (x + x[::-1]) / 2, atol=1e-7)
    assert st.ingest_pcm(x) is x
    # peak < 0.5 -> 0.95 (core/asr_engine.py:512-516); louder or silent audio is left alone
    quiet = (x * np.float32(0.1)).astype(np.float32)
    b = st.boost_low_volume(quiet)
    peak = np.max(np.abs(quiet))
    assert np.array_equal(b, quiet / peak * 0.95) and abs(float(np.max(np.abs(b))) - 0.95) < 1e-6
    assert st.boost_low_volume(x) is x
    z = np.zeros(10, np.float32)
    assert st.boost_low_volume(z) is z


def test_gain_curve_properties():
    audio, segs = staging_cases()[0]
    gains = st.segment_gains(audio, segs)
    assert gains and all(0.1 - 1e-9 <= g <= 10.0 + 1e-9 for _, _, g in gains)
    curve = st.gain_curve(len(audio), gains)
    inside = np.zeros(len(audio), bool)
    for s, e, _ in gains:
        inside[s:e] = True
    assert np.all(curve[~inside] == 1.0)
    for s, e, g in gains:
        k = min(80, (e - s) // 4)
        assert np.all(curve[s + k:e - k] == np.float32(g))
