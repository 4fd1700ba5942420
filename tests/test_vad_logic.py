"""CPU suite for the host logic around the VAD network. The Silero model is not available
offline; the reference's functions are driven with a stand-in ONNX session (energy -> probability, with a carried state) and
ours with the same function in batch form, so everything around the network is compared: windows, context carry,
thresholding, padding, merging, retry and fallback."""
import io
import sys
from contextlib import redirect_stdout

import numpy as np
import pytest

from oracle import chunk_cases as cc
from sherpa_vietnamese_asr_b200 import vad

REF = "/root/reference"
import os
have_ref = os.path.isdir(os.path.join(REF, "core"))


def standin_prob(rows, gain):
    """rows [n, 576] -> probs [n]: a saturating function of the RMS of the 576 inputs (context included, so a wrong context
    carry changes the answer)."""
    rms = np.sqrt(np.mean(rows.astype(np.float64) ** 2, axis=1))
    return (1.0 - np.exp(-gain * rms)).astype(np.float32)


class StandinSession:
    """Looks like the onnxruntime session the reference calls once per window (core/vad_utils.py:98-100)."""

    def __init__(self, gain):
        self.gain, self.calls = gain, 0

    def run(self, _names, feeds):
        x, state = feeds["input"], feeds["state"]
        assert x.shape == (1, 576) and x.dtype == np.float32 and state.shape == (2, 1, 128) and int(feeds["sr"]) == 16000
        self.calls += 1
        return [standin_prob(x, self.gain).reshape(1, 1), state]


@pytest.fixture(scope="module")
def vu():
    if not have_ref:
        pytest.skip("/root/reference not present (GPU box)")
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    with redirect_stdout(io.StringIO()):
        import core.vad_utils as m
    return m


def _audio(seed, seconds, level=1.0):
    return (cc.silence_audio(seed, seconds) * np.float32(level)).astype(np.float32)


@pytest.mark.parametrize("seed,seconds,level,gain", [(1, 30.0, 1.0, 12.0), (2, 75.0, 1.0, 6.0), (3, 20.0, 0.01, 12.0), (4, 12.0, 1.0, 0.5),
                                                     (5, 10.0, 1.0, 0.05), (6, 0.02, 1.0, 12.0), (7, 40.0, 0.2, 40.0)])
def test_get_vad_segments_live(vu, monkeypatch, seed, seconds, level, gain):
    audio = _audio(seed, seconds, level)
    sess = StandinSession(gain)
    monkeypatch.setattr(vu, "_get_vad_session", lambda: sess)
    for kw in ({}, {"fallback_full": False}, {"threshold": 0.5, "min_silence_ms": 300, "padding_ms": 200, "merge_gap_ms": 0}):
        with redirect_stdout(io.StringIO()):
            want = vu.get_vad_segments(audio, **kw)
        want_probs = vu.get_cached_vad_probs() if len(audio) >= 512 else None
        got, probs = vad.get_vad_segments(audio, lambda rows: standin_prob(rows, gain), **kw)
        assert got == want
        if want_probs is not None:
            assert np.array_equal(probs, want_probs)
    vu._last_vad_probs = None


def test_segments_from_probs_live(vu, monkeypatch):
    rng = np.random.default_rng(0)
    for trial in range(40):
        n = int(rng.integers(1, 400))
        probs = np.clip(np.repeat(rng.uniform(0, 1, n // 5 + 1), 5)[:n] + rng.normal(0, 0.1, n), 0, 1).astype(np.float32)
        it = iter(probs.tolist())

        class S:
            def run(self, _n, feeds):
                return [np.array([[next(it)]], np.float32), feeds["state"]]
        monkeypatch.setattr(vu, "_get_vad_session", lambda: S())
        th, sil, sp = float(rng.choice([0.2, 0.5, 0.8])), float(rng.choice([32, 100, 300])), float(rng.choice([0, 150, 250]))
        want = vu._run_vad_inference(np.zeros(n * 512, np.float32), 16000, th, sil, sp)
        assert vad.segments_from_probs(probs.tolist(), th, sil, sp) == want
    vu._last_vad_probs = None


def test_window_matrix_layout():
    audio = np.arange(512 * 3 + 100, dtype=np.float32)
    x = vad.window_matrix(audio)
    assert x.shape == (3, 576) and np.all(x[0, :64] == 0)
    assert np.array_equal(x[0, 64:], audio[:512]) and np.array_equal(x[1, :64], audio[512 - 64:512])
    assert np.array_equal(x[2, 64:], audio[1024:1536]) and np.array_equal(x[2, :64], audio[1024 - 64:1024])
    assert vad.window_matrix(audio[:100]).shape == (0, 576)


def test_speech_plan_merges_gaps_up_to_five_seconds():
    assert vad.merge_close_segments([(0, 10), (10 + 80000, 20 + 80000), (200000 + 80001, 300000)], vad.MAX_VAD_GAP, True) == \
        [(0, 20 + 80000), (200000 + 80001, 300000)]
    assert vad.merge_close_segments([(0, 10), (10 + 4000, 50)], 4000, False) == [(0, 10), (4010, 50)]
    assert vad.merge_close_segments([], 5, True) == []
    audio = _audio(11, 60.0)
    segs, probs = vad.speech_plan(audio, lambda rows: standin_prob(rows, 12.0))
    base, _ = vad.get_vad_segments(audio, lambda rows: standin_prob(rows, 12.0))
    assert segs == vad.merge_close_segments(base, vad.MAX_VAD_GAP, True) and len(probs) == len(audio) // 512
    assert all(b[0] - a[1] > vad.MAX_VAD_GAP for a, b in zip(segs, segs[1:]))
